"""Times build variants of libyart_b200.so against each other (same scenes, one process per variant).

  python tools/variant_bench.py [--workloads soup,sponza,mclaren] [--opts refill:inner] lib1.so lib2.so ...

Per variant: the device-resident primary-ray trace of the workload's camera (ms, hash of the hit
records so that variants can be checked for identical results) and the step time of 4-spp waves.
"""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(lib_path, workloads, steps):
    import numpy as np
    import yart_b200 as Y
    from yart_b200 import capi
    import bench
    Y.use_library(capi.load(lib_path))
    W, H, spp = 1920, 1080, 4
    out = []
    for wl in workloads:
        tris = bench.DEFAULT_TRIS[wl]
        bench.select_workload(wl, tris)
        sc = Y.Scene(bench.scene_path(tris, wl))
        cam = Y.make_camera(W, H, bench.CAM["focal"], bench.CAM["fnum"], bench.CAM["pos"], bench.CAM["target"], (0, 0, 0),
                            bench.CAM["exposure"])
        ctx = Y.Context(max_depth=bench.MAX_DEPTH)
        ctx.upload_scene(sc)
        ctx.set_camera(cam)
        ctx.begin_frame(W, H, spp * (steps + 3), 64, (0, 0, 0), Y.TONEMAP_AGX)
        n = W * H * spp
        rays, hits = ctx.device_alloc(n * 32), ctx.device_alloc(n * 20)
        ctx.generate_primary_rays(0, spp, rays)
        tr = ctx.trace_device(rays, n, hits, Y.TRACE_CLOSEST, repeat=5)
        got = np.empty(n, Y.COMPACT_HIT_DTYPE)
        ctx.d2h(got, hits)
        hh = hashlib.sha1(got.tobytes()).hexdigest()[:10]
        ctx.device_free(rays)
        ctx.device_free(hits)
        for k in range(3):
            ctx.render_wave(k * spp, spp, k * spp)
        s0 = ctx.stats()
        for k in range(3, 3 + steps):
            ctx.render_wave(k * spp, spp, k * spp)
        s1 = ctx.stats()
        step = (s1.gpuMs - s0.gpuMs) / steps
        hdr, _, _ = ctx.resolve()
        fh = hashlib.sha1(np.ascontiguousarray(hdr).tobytes()).hexdigest()[:10]
        out.append(f"{wl}: trace {tr:.3f} ms [{hh}]  step {step:.2f} ms [{fh}]")
        ctx.close()
    print(f"{os.path.basename(lib_path):28s} " + " | ".join(out), flush=True)


def main():
    args = sys.argv[1:]
    workloads, steps = ["soup"], 8
    libs = []
    i = 0
    while i < len(args):
        if args[i] == "--workloads":
            workloads = args[i + 1].split(",")
            i += 2
        elif args[i] == "--steps":
            steps = int(args[i + 1])
            i += 2
        elif args[i] == "--child":
            child(args[i + 1], workloads, steps)
            return
        else:
            libs.append(args[i])
            i += 1
    for lib in libs:
        subprocess.run([sys.executable, os.path.abspath(__file__), "--workloads", ",".join(workloads), "--steps", str(steps),
                        "--child", lib], check=False)


if __name__ == "__main__":
    main()
