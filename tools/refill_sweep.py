"""Sweep the warp-scheduling knobs of the traversal kernels (YcOptions::traceRefillMin / traceInnerMin) on a workload:
step time of 4-spp 1080p waves left in flight, frame hash (the knobs never change results).

  python tools/refill_sweep.py sponza 0:0,4:16,8:16,8:20,8:24,12:20,16:24
"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yart_b200 as Y
import bench

workload = sys.argv[1] if len(sys.argv) > 1 else "sponza"
pairs = [tuple(int(v) for v in p.split(":")) for p in (sys.argv[2] if len(sys.argv) > 2 else "0:0,4:16,8:16,8:20,8:24,12:20,16:24").split(",")]
tris = bench.DEFAULT_TRIS[workload]
bench.select_workload(workload, tris)
sc = Y.Scene(bench.scene_path(tris, workload))
cam = Y.make_camera(bench.W, bench.H, bench.CAM["focal"], bench.CAM["fnum"], bench.CAM["pos"], bench.CAM["target"], (0, 0, 0),
                    bench.CAM["exposure"])
for (refill, inner) in pairs + pairs[:1]:
    ctx = Y.Context(max_depth=bench.MAX_DEPTH, refill_min=refill, inner_min=inner)
    ctx.upload_scene(sc)
    ctx.set_camera(cam)
    ctx.begin_frame(bench.W, bench.H, 4 * 11, 64, (0, 0, 0), Y.TONEMAP_AGX)
    for k in range(3):
        ctx.render_wave_async(4 * k, 4, 4 * k)
    s0 = ctx.stats()
    for k in range(3, 11):
        ctx.render_wave_async(4 * k, 4, 4 * k)
    s1 = ctx.stats()
    hdr, _, _ = ctx.resolve()
    print(f"{workload} refill {refill:2d} inner {inner:2d}: {(s1.gpuMs - s0.gpuMs) / 8:.2f} ms/step [{hashlib.sha1(hdr.tobytes()).hexdigest()[:8]}]", flush=True)
    ctx.close()
