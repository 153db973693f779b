"""Step time of 4-spp 1080p waves left in flight against the number of paths in flight (YcOptions::maxPathsInFlight; two
lanes share it): python tools/capacity_sweep.py soup 0,4194304,2097152"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yart_b200 as Y
import bench

workload = sys.argv[1] if len(sys.argv) > 1 else "soup"
caps = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0,4194304,2097152,16777216").split(",")]
tris = bench.DEFAULT_TRIS[workload]
bench.select_workload(workload, tris)
sc = Y.Scene(bench.scene_path(tris, workload))
cam = Y.make_camera(bench.W, bench.H, bench.CAM["focal"], bench.CAM["fnum"], bench.CAM["pos"], bench.CAM["target"], (0, 0, 0),
                    bench.CAM["exposure"])
for cap in caps + caps[:1]:
    ctx = Y.Context(max_depth=bench.MAX_DEPTH, max_paths=cap)
    ctx.upload_scene(sc)
    ctx.set_camera(cam)
    ctx.begin_frame(bench.W, bench.H, 4 * 13, 64, (0, 0, 0), Y.TONEMAP_AGX)
    for k in range(3):
        ctx.render_wave_async(4 * k, 4, 4 * k)
    s0 = ctx.stats()
    for k in range(3, 13):
        ctx.render_wave_async(4 * k, 4, 4 * k)
    s1 = ctx.stats()
    hdr, _, _ = ctx.resolve()
    print(f"{workload} paths in flight {cap or 'default'}: {(s1.gpuMs - s0.gpuMs) / 10:.3f} ms/step, launches/step "
          f"{(s1.kernelLaunches - s0.kernelLaunches) / 10:.0f} [{hashlib.sha1(hdr.tobytes()).hexdigest()[:8]}]", flush=True)
    ctx.close()
