"""`-m gpu`: the yr_* renderer over several contexts / devices on the real library — one driver thread per GPU inside
libyart_b200.so, NCCL for distinct GPUs, the in-process group transport when contexts share a GPU (a one-GPU box)."""
import time

import numpy as np
import pytest

import harness as H
import yart_b200 as Y

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("cuda_lib")]


def renderer(scene="material_zoo", w=160, h=90, **kw):
    cam = H.scene_camera(scene)
    sc = Y.Scene(H.scene_file(scene))
    c = Y.make_camera(w, h, cam["focal"], cam["fnum"], cam["pos"], cam["target"], (0, 0, 0), cam["exposure"])
    kw.setdefault("samples", 48)
    return Y.Renderer(w, h, c, sc, tile_size=16, tonemap=Y.TONEMAP_AGX, first_wave_samples=16, max_wave_samples=32, max_depth=6, **kw)


def devices(n):
    import torch
    have = torch.cuda.device_count()
    return [i % have for i in range(n)]


@pytest.mark.parametrize("sharding", [Y.SHARD_TILES, Y.SHARD_BUCKETS])
@pytest.mark.parametrize("n", [2, 3])
def test_yr_create_multi_equals_one_gpu_bitwise(sharding, n):
    r1 = renderer()
    d1 = r1.render_sync()
    hdr1, ldr1, _ = r1.read()
    r1.close()
    rn = renderer(devices=devices(n), sharding=sharding)
    waves, done = [], []
    rn.on_wave_complete(lambda rd, wd: waves.append(wd["rays"]))
    rn.on_done(lambda rd, aborted: done.append(aborted))
    dn = rn.render_sync()
    hdr, ldr, st = rn.read()
    assert H.bits_equal(hdr, hdr1).all() and H.bits_equal(ldr, ldr1).all()
    assert dn["total_rays"] == d1["total_rays"] == st.raysReference == sum(waves)
    assert len(waves) == 2 and done == [False]
    rn.render()
    rn.abort()
    rn.wait()
    assert len(done) == 2
    d2 = rn.render_sync()
    assert d2["total_rays"] == d1["total_rays"] and H.bits_equal(rn.read()[0], hdr1).all()
    rn.close()


def test_async_render_from_a_worker_thread_on_the_last_device():
    """yr_render runs on a fresh host thread whose current CUDA device defaults to 0; every entry point now
    re-selects the context's device.  Render on the highest-numbered GPU (device 0 on a one-GPU box) asynchronously,
    with a second context alive on device 0, and compare with the synchronous frame."""
    import torch
    dev = torch.cuda.device_count() - 1
    other = Y.Context(device=0)
    r = renderer(device=dev)
    d1 = r.render_sync()
    hdr1, _, _ = r.read()
    r.render()
    assert r.wait() is True
    hdr2, _, st = r.read()
    assert H.bits_equal(hdr1, hdr2).all() and st.raysReference == d1["total_rays"]
    r.close()
    other.close()


def test_abort_returns_quickly_mid_wave():
    cam = H.scene_camera("cornell")
    sc = Y.Scene(H.scene_file("cornell"))
    c = Y.make_camera(1920, 1080, cam["focal"], cam["fnum"], cam["pos"], cam["target"])
    r = Y.Renderer(1920, 1080, c, sc, samples=1024, first_wave_samples=1024, max_wave_samples=1024)  # ~10 s of GPU work in one wave
    done = []
    r.on_done(lambda rd, aborted: done.append((aborted, rd["samples_taken"])))
    r.render()
    time.sleep(0.5)
    t0 = time.time()
    r.abort()
    assert r.wait() is False
    assert time.time() - t0 < 1.0 and done == [(True, 0)]  # polled per chunk and per bounce, not per wave
    # and the renderer is reusable: a short render afterwards matches a fresh renderer's
    r.close()
