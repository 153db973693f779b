"""Parity checks shared by the hostsim (CPU, `-m "not gpu"`) and CUDA (`-m gpu`) test modules:
each replays a committed golden input through whichever build of the C ABI is active
(yart_b200.use_library) and compares with the reference's recorded output."""
from __future__ import annotations

import ast
import glob
import os

import numpy as np

import harness as H
import yart_b200 as Y

# Every KAT is +,-,*,/,sqrt in IEEE fp32 without FMA contraction plus the glibc-exact
# sinf/cosf/logf/expf/log2f/powf of csrc/libm_exact.cuh: bit-exact on the GPU too.
TRANSCENDENTAL_COLUMNS = {}


def golden_files(prefix: str):
    return sorted(glob.glob(os.path.join(H.GOLDEN, prefix + "_*.npz")))


def load(path):
    return np.load(path, allow_pickle=False)


def scene_from_golden(g):
    name = str(g["scene"])
    kw = ast.literal_eval(str(g["kwargs"])) if "kwargs" in g.files else {}
    return name, kw


def check_kat(ctx_factory, path, exact: bool, rtol=2e-5, atol=1e-6):
    g = load(path)
    kind, scene, blob, ref = str(g["kind"]), str(g["scene"]), g["blob"].tobytes(), g["out"]
    ctx = ctx_factory()
    keep = []
    if scene:
        sc = Y.Scene(H.scene_file(scene))
        ctx.upload_scene(sc)
        keep.append(sc)
    if kind == "camera":
        import struct
        w, h, focal, fnum, sides = struct.unpack_from("<IIffI", blob, 0)
        v = np.frombuffer(blob, np.float32, 9, 20)
        ctx.set_camera(Y.make_camera(w, h, focal, fnum, v[0:3], v[3:6], v[6:9], 0.0, sides))
    W = H.KAT_OUT_WORDS[kind]
    out = ctx.kat(kind, blob, H.kat_count(kind, blob) * W)
    assert out.shape == ref.shape
    eq = H.bits_equal(out, ref).reshape(-1, W)
    if exact:
        assert eq.all(), f"{os.path.basename(path)}: {(~eq).sum()} words differ; columns {np.flatnonzero(~eq.all(0))}"
        return
    loose = TRANSCENDENTAL_COLUMNS.get(kind, [])
    strict = [c for c in range(W) if c not in loose]
    assert eq[:, strict].all(), (f"{os.path.basename(path)}: exact columns differ: "
                                 f"{[c for c in strict if not eq[:, c].all()]}")
    if loose:
        o, r = out.reshape(-1, W)[:, loose], ref.reshape(-1, W)[:, loose]
        assert np.allclose(o, r, rtol=rtol, atol=atol, equal_nan=True), f"{os.path.basename(path)}: beyond tolerance"


def check_trace(ctx_factory, path):
    """Hit triangle IDs bit-exact, t bit-exact (the bar is 1e-5 relative), all Hit fields exact."""
    g = load(path)
    name, kw = scene_from_golden(g)
    sc = Y.Scene(H.scene_file(name, **kw))
    ctx = ctx_factory()
    ctx.upload_scene(sc)
    rays = g["rays"]
    for mode, key in ((Y.TRACE_CLOSEST, "closest"), (Y.TRACE_ANY, "anyhit")):
        ref = g[key]
        hits, _ = ctx.trace(rays, mode)
        assert np.array_equal(hits["didHit"], ref["didHit"]), f"{name}/{key}: didHit differs"
        m = ref["didHit"] == 1
        assert np.array_equal(hits["attenuation"].view(np.uint32), ref["attenuation"].view(np.uint32))
        if key == "closest":
            assert np.array_equal(hits["prim"][m], ref["prim"][m]), f"{name}: hit triangle IDs differ"
            rel = np.abs(hits["t"][m] - ref["t"][m]) / np.abs(ref["t"][m])
            assert rel.max(initial=0.0) <= 1e-5
            for f in ("t", "material", "lightIdx", "backSide", "p", "n", "tg", "uv"):
                assert np.array_equal(hits[f][m].view(np.uint32), ref[f][m].view(np.uint32)), f"{name}: Hit::{f} differs"
    # counting trace: same hits, and the reference-traversal work counters are populated
    hits2, st = ctx.trace(rays, Y.TRACE_CLOSEST | Y.TRACE_COUNT)
    assert np.array_equal(hits2["prim"], ctx.trace(rays, Y.TRACE_CLOSEST)[0]["prim"])
    assert st.boxTests >= len(rays) and st.triTests > 0
    return sc, ctx


TONEMAPS = {"none": Y.TONEMAP_NONE, "agx": Y.TONEMAP_AGX, "golden": Y.TONEMAP_AGX_GOLDEN, "punchy": Y.TONEMAP_AGX_PUNCHY}


def render_golden(path, **renderer_kw):
    g = load(path)
    name, kw = scene_from_golden(g)
    w, h, spp, first, mx, depth = (int(v) for v in g["settings"])
    cam = H.scene_camera(name, **kw)
    sc = Y.Scene(H.scene_file(name, **kw))
    c = Y.make_camera(w, h, cam["focal"], cam["fnum"], cam["pos"], cam["target"], (0, 0, 0), cam["exposure"],
                      cam.get("sides", 0))
    if "integrator" in g.files and str(g["integrator"]) == "naive":
        renderer_kw = dict(renderer_kw, integrator=Y.INTEGRATOR_NAIVE)
    if "sampler" in g.files:
        renderer_kw = dict(renderer_kw, sampler={"naive": Y.SAMPLER_NAIVE, "stratified": Y.SAMPLER_STRATIFIED}[str(g["sampler"])])
    if "scrambler" in g.files:
        renderer_kw = dict(renderer_kw, scrambler={"owen": Y.SCRAMBLER_OWEN, "binary": Y.SCRAMBLER_BINARY_PERMUTE}[str(g["scrambler"])])
    r = Y.Renderer(w, h, c, sc, samples=spp, first_wave_samples=first, max_wave_samples=mx, max_depth=depth,
                   tonemap=TONEMAPS[str(g["tonemap"])], **renderer_kw)
    data = r.render_sync()
    hdr, ldr, st = r.read()
    r.close()
    return g, data, hdr, ldr, st


def check_render(path, exact: bool):
    g, data, hdr, ldr, st = render_golden(path)
    tag = os.path.basename(path)
    assert data["samples_taken"] == data["total_samples"]
    if exact:
        assert data["total_rays"] == int(g["rays"]), f"{tag}: ray count {data['total_rays']} vs {int(g['rays'])}"
        assert H.bits_equal(hdr, g["hdr"]).all(), f"{tag}: HDR differs"
        assert H.bits_equal(ldr, g["ldr"]).all(), f"{tag}: LDR differs"
    else:
        # GPU.  north_star tolerance: per-pixel relative MSE < 1e-3 at equal spp with the reference's
        # sampler seeds, HDR and after AgX.  The path arithmetic is bit-compatible (see above), so the
        # HDR frame, the LDR frame and the ray count are in fact identical.
        assert H.rel_mse(hdr, g["hdr"]) < 1e-3, f"{tag}: HDR relMSE {H.rel_mse(hdr, g['hdr'])}"
        assert H.rel_mse(ldr, g["ldr"]) < 1e-3, f"{tag}: LDR relMSE {H.rel_mse(ldr, g['ldr'])}"
        assert data["total_rays"] == int(g["rays"]), f"{tag}: ray count {data['total_rays']} vs {int(g['rays'])}"
        assert H.bits_equal(hdr, g["hdr"]).all(), f"{tag}: HDR differs in {(~H.bits_equal(hdr, g['hdr'])).sum()} words"
        assert H.bits_equal(ldr, g["ldr"]).all(), f"{tag}: LDR differs in {(~H.bits_equal(ldr, g['ldr'])).sum()} words"
    assert st.raysExtend >= data["total_rays"] - st.raysShadow
    if "integrator" in g.files and str(g["integrator"]) == "naive":
        assert st.raysShadow == 0 and st.raysExtend == data["total_rays"]  # no NEE: one closest-hit ray per segment
    return hdr, ldr
