"""Renders the same waves repeatedly and compares the frames bitwise (a scheduling race shows up as a differing pixel).
python tools/stress_determinism.py [workload] [repeats] [spp] [tail_threshold] [waves]
waves > 1: that many progressive waves of spp samples each, left in flight (yc_render_wave_async)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import yart_b200 as Y
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "sponza"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 1
tail = int(sys.argv[4]) if len(sys.argv) > 4 else 0
waves = int(sys.argv[5]) if len(sys.argv) > 5 else 1
tris = bench.DEFAULT_TRIS[wl]
bench.select_workload(wl, tris)
sc = Y.Scene(bench.scene_path(tris, wl))
W, H = 1920, 1080
cam = Y.make_camera(W, H, bench.CAM["focal"], bench.CAM["fnum"], bench.CAM["pos"], bench.CAM["target"], (0, 0, 0), bench.CAM["exposure"])
first, bad = None, 0
for r in range(reps):
    ctx = Y.Context(max_depth=bench.MAX_DEPTH, tail_threshold=tail)
    ctx.upload_scene(sc)
    ctx.set_camera(cam)
    ctx.begin_frame(W, H, spp * waves, 64, (0, 0, 0), Y.TONEMAP_AGX)
    if waves == 1:
        ctx.render_wave(0, spp, 0)
    else:
        for k in range(waves):
            ctx.render_wave_async(k * spp, spp, k * spp)
    hdr, _, st = ctx.resolve()
    ctx.close()
    if first is None:
        first, rays0 = hdr, st.raysReference
        continue
    diff = (hdr.view(np.uint32) != first.view(np.uint32)).any(-1)
    if diff.any() or st.raysReference != rays0:
        bad += 1
        ys, xs = np.nonzero(diff)
        print(f"run {r}: {diff.sum()} pixels differ, rays {st.raysReference} vs {rays0}, first at {list(zip(xs[:5].tolist(), ys[:5].tolist()))}", flush=True)
print(f"{wl} spp {spp} tail {tail} waves {waves}: {bad} of {reps - 1} repeats differ from the first")
