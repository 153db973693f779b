"""Waves waited for one at a time (yc_render_wave) against waves left in flight (yc_render_wave_async), same process,
same box: step time of `--steps` waves of `--spp` samples of the 1080p frame per workload, and a hash of the frames.

  python tools/async_ab.py [--workloads soup,sponza,mclaren] [--spp 4] [--steps 10] [--reps 3]
"""
import argparse
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="soup,sponza,mclaren")
    ap.add_argument("--spp", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import yart_b200 as Y
    import bench
    W, H = 1920, 1080
    for wl in a.workloads.split(","):
        tris = bench.DEFAULT_TRIS[wl]
        bench.select_workload(wl, tris)
        sc = Y.Scene(bench.scene_path(tris, wl))
        cam = Y.make_camera(W, H, bench.CAM["focal"], bench.CAM["fnum"], bench.CAM["pos"], bench.CAM["target"], (0, 0, 0),
                            bench.CAM["exposure"])
        ctx = Y.Context(max_depth=bench.MAX_DEPTH)
        ctx.upload_scene(sc)
        ctx.set_camera(cam)
        line = []
        for rep in range(a.reps):
            for mode in ("sync", "async"):
                f = ctx.render_wave_async if mode == "async" else ctx.render_wave
                ctx.begin_frame(W, H, a.spp * (a.steps + 3), 64, (0, 0, 0), Y.TONEMAP_AGX)
                for k in range(3):
                    f(k * a.spp, a.spp, k * a.spp)
                s0 = ctx.stats()
                t0 = time.time()
                for k in range(3, 3 + a.steps):
                    f(k * a.spp, a.spp, k * a.spp)
                s1 = ctx.stats()
                wall = (time.time() - t0) * 1e3 / a.steps
                hdr, _, _ = ctx.resolve()
                line.append(f"{mode} {(s1.gpuMs - s0.gpuMs) / a.steps:.3f} ms (wall {wall:.3f}) [{hashlib.sha1(hdr.tobytes()).hexdigest()[:8]}]")
        print(f"{wl}: " + " | ".join(line), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
