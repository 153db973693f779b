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


@pytest.mark.parametrize("mode", ["nccl_tiles", "nccl_tiles_reduce"])
def test_one_process_per_gpu_direct_frame_delivery_equals_one_gpu_bitwise(tmp_path, mode):
    """Two processes, one GPU each, NCCL inside the library (yr_create_dist).  nccl_tiles: rank 0's combined frames are
    mapped into rank 1 and both finalize kernels store their tiles there (peer memory; the per-wave collective is a
    barrier); nccl_tiles_reduce: the ncclReduce path.  Every wave's frame as the callback sees it, the final frames of
    two consecutive renders and the ray count equal one GPU's."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL refuses two ranks on one device)")
    import test_multi_gpu_cpu as M
    port = M.free_port()
    H.scene_file("material_zoo")
    procs = [subprocess.Popen([sys.executable, M.WORKER, str(r), "2", str(port), mode, str(tmp_path)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    for p in procs:
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, out[-3000:]
    cam = H.scene_camera("material_zoo")
    sc = Y.Scene(H.scene_file("material_zoo"))
    c = Y.make_camera(160, 90, cam["focal"], cam["fnum"], cam["pos"], cam["target"], (0, 0, 0), cam["exposure"])
    r1 = Y.Renderer(160, 90, c, sc, tile_size=16, tonemap=Y.TONEMAP_AGX, samples=56, first_wave_samples=8, max_wave_samples=16,
                    max_depth=6, traversal=Y.TRAVERSAL_REFERENCE_ORDER)
    waves = []
    target = np.zeros((90, 160, 4), np.float32)
    r1.set_frame_target(target)
    r1.on_wave_complete(lambda rd, wd: waves.append(target.copy()))
    d1 = r1.render_sync()
    hdr1, ldr1, _ = r1.read()
    r1.close()
    rays, direct = (int(v) for v in np.load(tmp_path / f"{mode}_info.npy"))
    assert rays == d1["total_rays"] and direct == (1 if mode == "nccl_tiles" else 0)
    for k in range(2):
        assert H.bits_equal(np.load(tmp_path / f"{mode}_hdr{k}.npy"), hdr1).all()
        assert H.bits_equal(np.load(tmp_path / f"{mode}_ldr{k}.npy"), ldr1).all()
    got = np.load(tmp_path / f"{mode}_waves.npy")
    assert len(waves) == 4 and got.shape[0] == 8  # waves of 8, 16, 16, 16 samples, two renders
    for k in range(8):
        assert H.bits_equal(got[k], waves[k % 4]).all(), k
