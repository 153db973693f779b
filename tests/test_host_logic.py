"""Host-side logic (scene loading / validation, camera, wave schedule, renderer API semantics),
exercised through the hostsim build of the same sources."""
import os
import struct

import numpy as np
import pytest

import harness as H
import parity_common as PC
import yart_b200 as Y
from yart_b200 import scenes

pytestmark = pytest.mark.usefixtures("hostsim_lib")


def reference_wave_schedule(samples, first, mx):
    """TileRenderer::renderImpl + finishTile (tile-renderer.hpp:120-124, 264-288) restated."""
    waves, wave, remaining, k = [], min(first, samples), samples, 0
    while wave > 0:
        waves.append(wave)
        remaining -= wave
        nxt = min(wave * 2, mx) if (k > 0 or wave > 1) else 1
        wave = min(nxt, remaining)
        k += 1
    return waves


@pytest.mark.parametrize("samples,first,mx", [(16, 16, 16), (8, 1, 4), (64, 4, 32), (100, 64, 128), (7, 2, 3), (1, 64, 128)])
def test_wave_schedule_matches_tile_renderer(samples, first, mx):
    cam = H.scene_camera("two_quads")
    sc = Y.Scene(H.scene_file("two_quads"))
    c = Y.make_camera(16, 16, cam["focal"], cam["fnum"], cam["pos"], cam["target"])
    r = Y.Renderer(16, 16, c, sc, samples=samples, first_wave_samples=first, max_wave_samples=mx)
    seen = []
    r.on_wave_complete(lambda rd, wd: seen.append((wd["wave"], wd["wave_samples"], rd["samples_taken"])))
    d = r.render_sync()
    want = reference_wave_schedule(samples, first, mx)
    assert [s[1] for s in seen] == want and [s[0] for s in seen] == list(range(len(want)))
    assert seen[-1][2] == samples == d["samples_taken"] and sum(want) == samples


def test_async_render_abort_wait():
    cam = H.scene_camera("cornell")
    sc = Y.Scene(H.scene_file("cornell"))
    c = Y.make_camera(32, 32, cam["focal"], cam["fnum"], cam["pos"], cam["target"])
    r = Y.Renderer(32, 32, c, sc, samples=64, first_wave_samples=1, max_wave_samples=2)
    r.render()
    r.wait()
    hdr_async, _, _ = r.read()
    d = r.render_sync()
    hdr_sync, _, _ = r.read()
    assert H.bits_equal(hdr_async, hdr_sync).all() and d["samples_taken"] == 64
    r.render()
    r.abort()  # stops between waves; must not hang or crash
    r.wait()
    hdr_part, _, _ = r.read()
    assert np.isfinite(hdr_part).all()


def test_renderer_without_scene_reports_no_scene():
    c = Y.make_camera(8, 8)
    r = Y.Renderer(8, 8, c, None, samples=1)
    with pytest.raises(Y.YartError, match="NO_SCENE"):
        r.render_sync()


def test_call_order_is_enforced():
    ctx = Y.Context()
    with pytest.raises(Y.YartError, match="NO_SCENE"):
        ctx.begin_frame(8, 8, 1)
    sc = Y.Scene(H.scene_file("two_quads"))
    ctx.upload_scene(sc)
    with pytest.raises(Y.YartError, match="STATE"):
        ctx.begin_frame(8, 8, 1)  # camera missing
    ctx.set_camera(Y.make_camera(8, 8))
    with pytest.raises(Y.YartError):
        ctx.frame = Y.capi.YcFrameDesc(8, 8, 1, 64)
        ctx.render_wave(0, 1, 0)  # no begin_frame yet
    ctx.begin_frame(8, 8, 1)
    with pytest.raises(Y.YartError, match="INVALID"):
        ctx.render_wave(0, 1, 0, rect=(4, 4, 8, 8))  # rectangle leaves the frame
    ctx.render_wave(0, 0, 0)  # zero samples: a no-op, like an empty wave
    with pytest.raises(Y.YartError, match="INVALID"):
        ctx.begin_frame(0, 8, 1)
    with pytest.raises(Y.YartError, match="INVALID"):
        ctx.begin_frame(8, 8, 1, shard_index=2, shard_count=2)


def test_scene_file_errors(tmp_path):
    with pytest.raises(Y.YartError, match="IO"):
        Y.Scene(str(tmp_path / "missing.ysc"))
    bad = tmp_path / "bad.ysc"
    bad.write_bytes(b"NOPE" + b"\0" * 64)
    with pytest.raises(Y.YartError, match="bad magic"):
        Y.Scene(str(bad))
    good = open(H.scene_file("two_quads"), "rb").read()
    trunc = tmp_path / "trunc.ysc"
    trunc.write_bytes(good[: len(good) // 2])
    with pytest.raises(Y.YartError, match="truncated|exceeds the file"):
        Y.Scene(str(trunc))
    # a face that indexes past the vertex array
    s = scenes.two_quads()
    s.meshes[0].faces[0, 1] = 10_000
    p = tmp_path / "oob.ysc"
    s.write(str(p))
    with pytest.raises(Y.YartError, match="out of range"):
        Y.Scene(str(p))
    # material index out of range
    s = scenes.two_quads()
    s.meshes[0].faces[0, 3] = 99
    s.write(str(p))
    with pytest.raises(Y.YartError, match="material index"):
        Y.Scene(str(p))
    # an empty mesh
    s = scenes.two_quads()
    s.meshes[1].faces = s.meshes[1].faces[:0]
    s.meshes[1].light_idx = s.meshes[1].light_idx[:0]
    s.lights = []
    s.write(str(p))
    with pytest.raises(Y.YartError, match="empty mesh"):
        Y.Scene(str(p))


def test_scene_without_lights_renders_background_only():
    s = scenes.two_quads()
    s.lights = []
    for m in s.meshes:
        m.light_idx[:] = -1
    p = os.path.join(H.CACHE, "two_quads_nolights.ysc")
    s.write(p)
    cam = s.camera
    ref = H.oracle_render(p, 24, 24, 4, cam, first=4, max=4, bg="0.2,0.3,0.4", tonemap="agx") if H.have_oracle() else None
    sc = Y.Scene(p)
    c = Y.make_camera(24, 24, cam["focal"], cam["fnum"], cam["pos"], cam["target"])
    r = Y.Renderer(24, 24, c, sc, samples=4, first_wave_samples=4, max_wave_samples=4, background=(0.2, 0.3, 0.4))
    r.render_sync()
    hdr, ldr, st = r.read()
    assert st.raysShadow == 0
    if ref is not None:
        assert H.bits_equal(hdr, ref["hdr"]).all() and H.bits_equal(ldr, ref["ldr"]).all()


def test_flattened_layout_invariants():
    sc = Y.Scene(H.scene_file("material_zoo"))
    f = sc.flat
    assert f.nNodes == 4 and f.nMeshes == 2 and f.hasAlpha == 1
    assert f.nInfinite == 2 and f.nArea == 4 and f.nLights == 6
    nodes, idx = sc.bvh(0)
    assert sorted(idx.tolist()) == list(range(f.meshes[0].nTris))  # a permutation
    inner = nodes[nodes["span"] == 0]
    assert f.meshes[0].nInner == len(inner)
    leaves = nodes[nodes["span"] > 0]
    assert leaves["span"].sum() == f.meshes[0].nTris and leaves["span"].max() <= 20
    # children adjacent, parent box contains children (padded boxes, exact min/max folds)
    for n in inner:
        a, b = nodes[n["left"]], nodes[n["left"] + 1]
        assert (np.minimum(a["min"], b["min"]) == n["min"]).all() and (np.maximum(a["max"], b["max"]) == n["max"]).all()


def test_camera_make_defaults_and_up_vector():
    a = Y.make_camera(640, 480, 50.0, 2.0, (1, 2, 3), (0, 0, 0))
    b = Y.make_camera(640, 480, 50.0, 2.0, (1, 2, 3), (0, 0, 0), up=(0, 1, 0))
    assert bytes(a) == bytes(b)
    assert abs(a.apertureRadius - (50.0 / 2000.0) / 2.0) < 1e-9
    pin = Y.make_camera(640, 480, 50.0, 0.0, (1, 2, 3), (0, 0, 0))
    assert pin.apertureRadius == 0.0


def test_u8_unit_conversion_is_the_division():
    """csrc/texture.cuh u8ToUnit: q = b*r, rem = fma(-q, 255, b), q' = fma(rem, r, q) with r = fl(1/255) equals the
    correctly rounded float(b) / 255.0f for every byte (the reference's conversion, texture.hpp:108-116)."""
    from fractions import Fraction
    f32 = np.float32

    def rn(x: Fraction) -> np.float32:  # exact round-to-nearest-even of a rational to float32
        c = f32(float(x))
        cands = [c, np.nextafter(c, f32(np.inf)), np.nextafter(c, f32(-np.inf))]
        return min(cands, key=lambda v: (abs(Fraction(float(v)) - x), int(v.view(np.uint32)) & 1))

    r = f32(1.0) / f32(255.0)
    for b in range(256):
        fb = f32(b)
        q = rn(Fraction(float(fb)) * Fraction(float(r)))
        rem = rn(-Fraction(float(q)) * 255 + Fraction(float(fb)))
        q2 = rn(Fraction(float(rem)) * Fraction(float(r)) + Fraction(float(q)))
        assert q2 == fb / f32(255.0), b
