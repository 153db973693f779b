"""Isolated closest-hit launch over every sample offset 0..63 of the C2 camera (4 spp of 1080p primary rays per launch,
totalSamples 256): the wide walk against the reference-order walk.  One ray per ~5 offsets has d.x == 0 and o.x == 0 —
NaN slabs in the reference's box test, 14 K boxes walked alone; the wide walk clamps the reciprocal direction."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import yart_b200 as Y
import bench
sc = Y.Scene(bench.scene_path(1_000_000))
ctx = Y.Context(max_depth=1)
ctx.upload_scene(sc)
W, H = 1920, 1080
ctx.set_camera(Y.make_camera(W, H, 35.0, 0.0, (0, 0, 40), (0, 0, 0)))
n = W * H * 4
rays, hits = ctx.device_alloc(n * 32), ctx.device_alloc(n * 20)
ctx.begin_frame(W, H, 256, 64, (0, 0, 0), Y.TONEMAP_NONE)
out = {"wide": [], "reference": []}
for off in range(0, 64):
    ctx.generate_primary_rays(off * 4, 4, rays)
    out["wide"].append(ctx.trace_device(rays, n, hits, Y.TRACE_CLOSEST | Y.TRACE_WIDE, repeat=2))
    out["reference"].append(ctx.trace_device(rays, n, hits, Y.TRACE_CLOSEST | Y.TRACE_REFERENCE_ORDER, repeat=2))
for k, v in out.items():
    v = np.array(v)
    print(f"{k:10s} median {np.median(v):.3f} ms  min {v.min():.3f}  max {v.max():.3f}  max/median {v.max() / np.median(v):.3f}  launches above 1.1 x median: {(v > 1.1 * np.median(v)).sum()} of {len(v)}")
