"""One closest-hit launch per walk over 1080p primary rays of the 1M-triangle soup (for ncu):
python tools/trace_prof.py [spp] [walks: wide,reference] [repeat]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import yart_b200 as Y
import bench

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 2
walks = sys.argv[2].split(",") if len(sys.argv) > 2 else ["wide", "reference"]
repeat = int(sys.argv[3]) if len(sys.argv) > 3 else 1
sc = Y.Scene(bench.scene_path(1_000_000))
ctx = Y.Context(max_depth=1)
ctx.upload_scene(sc)
W, H = 1920, 1080
ctx.set_camera(Y.make_camera(W, H, 35.0, 0.0, (0, 0, 40), (0, 0, 0)))
ctx.begin_frame(W, H, 64, 64, (0, 0, 0), Y.TONEMAP_NONE)
n = W * H * spp
rays, hits = ctx.device_alloc(n * 32), ctx.device_alloc(n * 20)
ctx.generate_primary_rays(12, spp, rays)
flag = {"wide": Y.TRACE_WIDE, "reference": Y.TRACE_REFERENCE_ORDER}
for w in walks:
    ms = ctx.trace_device(rays, n, hits, Y.TRACE_CLOSEST | flag[w], repeat=repeat)
    _, st = None, None
    print(f"{w}: {n} rays {ms:.3f} ms -> {n / ms / 1e3:.1f} Mrays/s")
one = np.array([[0, 0, 40, 0.001, 0.01, 0.02, -0.99975, np.inf]], np.float32)
for w in walks:
    ctx.trace(one, Y.TRACE_CLOSEST | Y.TRACE_COUNT | flag[w])
    s0 = ctx.stats()
    import ctypes as C
    ms_c = C.c_float()
    Y.lib().yc_trace_device(ctx._h, rays, n, Y.TRACE_CLOSEST | Y.TRACE_COUNT | flag[w], hits, 1, C.byref(ms_c))
    s1 = ctx.stats()
    print(f"{w}: box tests/ray {(s1.boxTests - s0.boxTests) / n:.2f} tri tests/ray {(s1.triTests - s0.triTests) / n:.2f}")
