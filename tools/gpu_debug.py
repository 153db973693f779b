"""Scratch diagnostics for GPU-vs-golden differences (run under gpurun)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import harness as H, parity_common as PC, yart_b200 as Y
from yart_b200 import capi
Y.use_library(capi.load())

g = PC.load(os.path.join(H.GOLDEN, "kat_bsdf.npz"))
blob, ref = g["blob"].tobytes(), g["out"].reshape(-1, 27)
ctx = Y.Context(); sc = Y.Scene(H.scene_file("material_zoo")); ctx.upload_scene(sc)
out = ctx.kat("bsdf", blob, ref.size).reshape(-1, 27)
eq = H.bits_equal(out, ref).reshape(-1, 27)
rec = np.frombuffer(blob, np.uint32, offset=4).reshape(-1, 22)
print("bsdf cols mismatching:", (~eq).sum(0))
bad = np.flatnonzero(~eq[:, 21:24].all(1))
print("rows with normal mismatch", len(bad), "materials", np.unique(rec[bad, 0], return_counts=True))
for i in bad[:6]:
    print(" row", i, "mat", rec[i, 0], "out", out[i, 21:24], "ref", ref[i, 21:24], "ulps",
          out[i, 21:24].view(np.int32) - ref[i, 21:24].view(np.int32))
for c in range(27):
    if c == 4: continue
    d = np.abs(out[:, c] - ref[:, c]); fin = np.isfinite(d)
    if (~eq[:, c]).any():
        print(f" col {c}: mismatches {(~eq[:, c]).sum()} max abs {d[fin].max():.3g} max rel {(d[fin] / np.maximum(np.abs(ref[fin, c]), 1e-6)).max():.3g}")
print("flags differ:", (~eq[:, 4]).sum())

for tag in ("zoo", "zoo_waves", "cornell"):
    path = os.path.join(H.GOLDEN, f"render_{tag}.npz")
    g, data, hdr, ldr, st = PC.render_golden(path)
    ref = g["hdr"]
    err = np.abs(hdr[..., :3] - ref[..., :3]).max(-1)
    rel = err / (np.abs(ref[..., :3]).max(-1) + 1e-3)
    print(tag, "rays", data["total_rays"], int(g["rays"]), "relMSE", H.rel_mse(hdr, ref), "pixels rel>1e-3:", (rel > 1e-3).sum(),
          "exact pixels", H.bits_equal(hdr, ref).all(-1).sum(), "of", rel.size, "nan", np.isnan(hdr).sum(), np.isnan(ref).sum())
    ys, xs = np.nonzero(rel > 1e-2)
    if len(ys):
        print("  bbox y", ys.min(), ys.max(), "x", xs.min(), xs.max())
        for y, x in list(zip(ys, xs))[:8]:
            print("   ", y, x, hdr[y, x, :3], ref[y, x, :3])
