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
npx = W * H
rays, hits = ctx.device_alloc(npx * 4 * 32), ctx.device_alloc(npx * 4 * 20)
ctx.begin_frame(W, H, 16, 64, (0, 0, 0), Y.TONEMAP_NONE)
ctx.generate_primary_rays(0, 4, rays)
host = np.empty((npx * 4, 8), np.float32)
ctx.d2h(host, rays)
print("nonfinite rays:", (~np.isfinite(host[:, :7])).any(1).sum(), " |d| range", np.linalg.norm(host[:, 4:7], axis=1).min(), np.linalg.norm(host[:, 4:7], axis=1).max())
print("zero dir comps:", (host[:, 4:7] == 0).sum(0), "min |d comp|", np.abs(host[:, 4:7]).min(0))
for s in range(4):
    ms = ctx.trace_device(rays + s * npx * 32, npx, hits, Y.TRACE_CLOSEST, repeat=3)
    print(f"sample {s}: {ms:.3f} ms")
# bisect the slow sample by pixel-list ranges
def t(off, cnt):
    return ctx.trace_device(rays + off * 32, cnt, hits, Y.TRACE_CLOSEST, repeat=2)
times = [t(s * npx, npx) for s in range(4)]
s = int(np.argmax(times))
lo, cnt = s * npx, npx
while cnt > 4096:
    h = cnt // 2
    a, b = t(lo, h), t(lo + h, cnt - h)
    print(f"  range {lo}+{cnt}: halves {a:.3f} {b:.3f}")
    if a > b: cnt = h
    else: lo, cnt = lo + h, cnt - h
sub = host[lo:lo + cnt]
print("suspect rays", lo, cnt)
hs, st = ctx.trace(sub, Y.TRACE_CLOSEST | Y.TRACE_COUNT)
print("box tests in suspect block:", st.boxTests, "per ray", st.boxTests / cnt)
# per-ray counts
worst = []
for i in range(0, cnt, 64):
    _, st = ctx.trace(sub[i:i + 64], Y.TRACE_CLOSEST | Y.TRACE_COUNT)
    worst.append((st.boxTests, i))
worst.sort(reverse=True)
print("worst 64-ray groups:", worst[:5])
b, i = worst[0]
for k in range(i, min(i + 64, cnt)):
    _, st = ctx.trace(sub[k:k + 1], Y.TRACE_CLOSEST | Y.TRACE_COUNT)
    if st.boxTests > 2000:
        print("ray", lo + k, sub[k], "box tests", st.boxTests, "tri", st.triTests)
