"""Traversal microbench (C2): 1080p primary rays of the 1M-triangle soup, device resident."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import yart_b200 as Y
import bench

tris = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
repeat = int(sys.argv[2]) if len(sys.argv) > 2 else 5
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 4
sc = Y.Scene(bench.scene_path(tris))
import itertools
combos = [(0, 0)] + ([tuple(map(int, c.split(":"))) for c in sys.argv[4].split(",")] if len(sys.argv) > 4 else [])
ctx = Y.Context(max_depth=1)
ctx.upload_scene(sc)
W, H = 1920, 1080
ctx.set_camera(Y.make_camera(W, H, 35.0, 0.0, (0, 0, 40), (0, 0, 0)))
ctx.begin_frame(W, H, spp, 64, (0, 0, 0), Y.TONEMAP_NONE)
n = W * H * spp
rays, hits = ctx.device_alloc(n * 32), ctx.device_alloc(n * 20)
ctx.generate_primary_rays(0, spp, rays)
ms = ctx.trace_device(rays, n, hits, Y.TRACE_CLOSEST, repeat=repeat)
print(f"closest: {n} rays {ms:.3f} ms → {n / ms / 1e3:.1f} Mrays/s")
ref = np.empty(n, Y.COMPACT_HIT_DTYPE)
ctx.d2h(ref, hits)
for rf, im in combos[1:]:
    c2 = Y.Context(max_depth=1, refill_min=rf, inner_min=im)
    c2.upload_scene(sc)
    h2 = c2.device_alloc(n * 20)
    ms = c2.trace_device(rays, n, h2, Y.TRACE_CLOSEST, repeat=repeat)
    got = np.empty(n, Y.COMPACT_HIT_DTYPE)
    c2.d2h(got, h2)
    print(f"refill {rf:2d} inner {im:2d}: {ms:.3f} ms  same={got.tobytes() == ref.tobytes()}")
    c2.device_free(h2)
    c2.close()
