"""Sweep the tail-kernel hand-over threshold on a full-path workload (device-resident render_wave)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yart_b200 as Y
import bench

workload = sys.argv[1] if len(sys.argv) > 1 else "mclaren"
tris = bench.DEFAULT_TRIS[workload]
bench.select_workload(workload, tris)
sc = Y.Scene(bench.scene_path(tris, workload))
cam = Y.make_camera(bench.W, bench.H, bench.CAM["focal"], bench.CAM["fnum"], bench.CAM["pos"], bench.CAM["target"], (0, 0, 0), bench.CAM["exposure"])
for tail in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "-1,4096,16384,65536,262144,1048576").split(",")]:
    ctx = Y.Context(max_depth=bench.MAX_DEPTH, tail_threshold=tail)
    ctx.upload_scene(sc)
    ctx.set_camera(cam)
    ctx.begin_frame(bench.W, bench.H, 4, 64, (0, 0, 0), Y.TONEMAP_AGX)
    for _ in range(3):
        ctx.render_wave_async(0, 4, 0)
    s0 = ctx.stats()
    for _ in range(6):
        ctx.render_wave_async(0, 4, 0)  # waves left in flight, as bench.py times them
    s1 = ctx.stats()
    print(f"{workload} tail={tail:8d}: {(s1.gpuMs - s0.gpuMs) / 6:.2f} ms/step, launches/step {(s1.kernelLaunches - s0.kernelLaunches) / 6:.0f}, rays {s1.raysReference - s0.raysReference}")
    ctx.close()
