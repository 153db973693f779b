"""N > 1 host logic on the CPU: two processes, gloo backend, each standing in for one GPU."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import harness as H
import yart_b200 as Y

pytestmark = pytest.mark.usefixtures("hostsim_lib")
WORKER = os.path.join(H.ROOT, "tests", "mp_worker.py")


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def launch(mode, out_dir, world=2):
    H.hostsim()  # build once before the ranks race for it
    H.scene_file("cornell")
    port = free_port()
    procs = [subprocess.Popen([sys.executable, WORKER, str(r), str(world), str(port), mode, str(out_dir)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    for p in procs:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, out[-3000:]


def single(spp_total, waves, tile=16):
    cam = H.scene_camera("cornell")
    sc = Y.Scene(H.scene_file("cornell"))
    c = Y.make_camera(48, 48, cam["focal"], cam["fnum"], cam["pos"], cam["target"])
    ctx = Y.Context()
    ctx.upload_scene(sc)
    ctx.set_camera(c)
    ctx.begin_frame(48, 48, spp_total, tile, (0, 0, 0), Y.TONEMAP_AGX)
    taken = 0
    for wv in waves:
        ctx.render_wave(taken, wv, taken)
        taken += wv
    return ctx.resolve()


def test_two_ranks_tile_sharding_equals_one_rank_bitwise(tmp_path):
    launch("tiles", tmp_path)
    hdr1, ldr1, st = single(8, [8])
    assert H.bits_equal(np.load(tmp_path / "tiles_hdr.npy"), hdr1).all()
    assert H.bits_equal(np.load(tmp_path / "tiles_ldr.npy"), ldr1).all()
    assert int(np.load(tmp_path / "tiles_rays.npy")[0]) == st.raysReference


def test_two_ranks_wave_sharding_equals_progressive_render(tmp_path):
    launch("waves", tmp_path)
    hdr1, _, st = single(16, [8, 8])  # one rank, two progressive waves of 8 (finishTile's blend)
    got = np.load(tmp_path / "waves_hdr.npy")
    # 0.5 * a + 0.5 * b on both sides: identical rounding for two equal waves
    assert H.bits_equal(got, hdr1).all()
    assert int(np.load(tmp_path / "waves_rays.npy")[0]) == st.raysReference


def test_two_ranks_bucket_sharding_equals_one_rank_bitwise(tmp_path):
    """GMoN accumulation buffers combined across ranks (all-reduce of the bucket planes as int32), two
    progressive waves of 16 samples (m = 3 buckets x 2 pixel classes: every rank owns one class of each bucket)."""
    launch("buckets", tmp_path)
    hdr1, ldr1, st = single(32, [16, 16])
    assert H.bits_equal(np.load(tmp_path / "buckets_hdr.npy"), hdr1).all()
    assert H.bits_equal(np.load(tmp_path / "buckets_ldr.npy"), ldr1).all()
    assert int(np.load(tmp_path / "buckets_rays.npy")[0]) == st.raysReference


# ------------------------------------------------------------------------------------------
# the same through yr_* only: the library owns the wave loop, the sharding and the collectives
# ------------------------------------------------------------------------------------------
def single_renderer(**kw):
    cam = H.scene_camera("cornell")
    sc = Y.Scene(H.scene_file("cornell"))
    c = Y.make_camera(48, 48, cam["focal"], cam["fnum"], cam["pos"], cam["target"])
    r = Y.Renderer(48, 48, c, sc, tile_size=16, tonemap=Y.TONEMAP_AGX, samples=32, first_wave_samples=16, max_wave_samples=16, **kw)
    return r


@pytest.mark.parametrize("mode", ["yr_tiles", "yr_buckets"])
def test_two_processes_through_yr_create_dist_equal_one_renderer_bitwise(tmp_path, mode):
    """yr_create_dist_custom with gloo as the sum collective (NCCL's place on the GPU box): rank 0's frame, the
    whole-job ray count and the wave callbacks equal the single renderer's."""
    launch(mode, tmp_path)
    r = single_renderer()
    data = r.render_sync()
    hdr1, ldr1, _ = r.read()
    r.close()
    assert H.bits_equal(np.load(tmp_path / f"{mode}_hdr.npy"), hdr1).all()
    assert H.bits_equal(np.load(tmp_path / f"{mode}_ldr.npy"), ldr1).all()
    rays, n_waves, wave_ray_sum = (int(v) for v in np.load(tmp_path / f"{mode}_rays.npy"))
    assert rays == data["total_rays"] and n_waves == 2 and wave_ray_sum == rays
    if mode == "yr_buckets":
        assert H.bits_equal(np.load(tmp_path / f"{mode}_hdr_rank1.npy"), hdr1).all()


@pytest.mark.parametrize("sharding", [Y.SHARD_TILES, Y.SHARD_BUCKETS])
@pytest.mark.parametrize("n", [2, 3])
def test_one_process_yr_create_multi_equals_one_renderer_bitwise(sharding, n):
    """yr_create_multi: n contexts, one driver thread each, the in-process group transport (what two contexts on one
    GPU use); frames, ray counts and callbacks equal the single renderer's."""
    r1 = single_renderer()
    d1 = r1.render_sync()
    hdr1, ldr1, _ = r1.read()
    r1.close()
    rn = single_renderer(devices=list(range(n)), sharding=sharding)
    waves, tiles, done = [], [], []
    rn.on_wave_complete(lambda rd, wd: waves.append(wd["rays"]))
    rn.on_tile_complete(lambda rd, td: tiles.append(td["index"]))
    rn.on_done(lambda rd, aborted: done.append(aborted))
    dn = rn.render_sync()
    hdr, ldr, st = rn.read()
    assert H.bits_equal(hdr, hdr1).all() and H.bits_equal(ldr, ldr1).all()
    assert dn["total_rays"] == d1["total_rays"] == st.raysReference and sum(waves) == d1["total_rays"]
    assert len(waves) == 2 and len(tiles) == 2 * 9 and done == [False]
    # asynchronous render + abort: exactly one completion callback, no hang, the renderer stays usable
    rn.render()
    rn.abort()
    rn.wait()
    assert len(done) == 2
    d2 = rn.render_sync()
    hdr2, _, _ = rn.read()
    assert d2["total_rays"] == d1["total_rays"] and H.bits_equal(hdr2, hdr1).all()
    rn.close()


def test_abort_stops_a_long_render_between_chunks():
    """yr_abort returns at once and the wave in flight stops at its next chunk / bounce boundary (the reference
    stops between tiles): a render that would take many seconds ends early with the aborted callback."""
    import time
    cam = H.scene_camera("cornell")
    sc = Y.Scene(H.scene_file("cornell"))
    c = Y.make_camera(64, 64, cam["focal"], cam["fnum"], cam["pos"], cam["target"])
    r = Y.Renderer(64, 64, c, sc, samples=4096, first_wave_samples=4096, max_wave_samples=4096)  # 16.8 M paths: minutes on the CPU build
    done = []
    r.on_done(lambda rd, aborted: done.append((aborted, rd["samples_taken"])))
    t0 = time.time()
    r.render()
    time.sleep(0.3)
    r.abort()
    assert r.wait() is False
    assert time.time() - t0 < 90 and done == [(True, 0)]  # one bounce of the chunks in flight at most
    r.close()


@pytest.mark.parametrize("force_reduce", [False, True])
def test_direct_frame_delivery_in_process_group_with_partial_and_repeated_waves(force_reduce, monkeypatch):
    """Tile sharding where every participant reaches the root's memory (here: one process, the in-process group): the
    finalize kernel stores finished pixels into the root's combined frame and yc_comm_reduce_frames is a barrier.  The
    root's frame equals one context's after every wave — also after a wave that finalized part of the frame only (the
    stale path pushes the shard), after two waves without a reduce in between, and over a second frame (the two
    alternating copies of the combined frame are reused)."""
    import threading
    if force_reduce:  # the summing path behind the same calls (what participants without peer access fall back to)
        monkeypatch.setenv("YART_B200_FRAMES_REDUCE", "1")
    n, size, tile = 3, 48, 16
    cam = H.scene_camera("cornell")
    sc = Y.Scene(H.scene_file("cornell"))
    c = Y.make_camera(size, size, cam["focal"], cam["fnum"], cam["pos"], cam["target"])
    # (sample offset, samples, rectangle or None, reduce after it?)
    plan = [(0, 2, None, True), (2, 1, (0, 0, 20, size), True), (2, 1, (20, 0, size - 20, size), True), (3, 2, None, False),
            (5, 1, None, True), (6, 2, None, True)]

    def run(ctx, rank, world, out):
        for frame in range(2):
            ctx.begin_frame(size, size, 8, tile, (0, 0, 0), Y.TONEMAP_AGX, shard_index=rank, shard_count=world)
            for (s0, k, rect, reduce) in plan:
                ctx.render_wave(s0, k, s0, rect=rect)
                if world > 1 and reduce:
                    ctx.comm_reduce_frames(0)
                if reduce and rank == 0:
                    out.append(ctx.resolve_combined() if world > 1 else ctx.resolve()[:2])

    one = Y.Context()
    one.upload_scene(sc)
    one.set_camera(c)
    want = []
    run(one, 0, 1, want)
    ctxs = [Y.Context() for _ in range(n)]
    for x in ctxs:
        x.upload_scene(sc)
        x.set_camera(c)
    Y.comm_init_all(ctxs)
    got, errs = [], []

    def worker(r):
        try:
            run(ctxs[r], r, n, got)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(120)
    assert not errs, errs
    assert all(x.comm_frames_direct() != force_reduce for x in ctxs)
    assert len(got) == len(want) == 10
    for k, ((hdr, ldr), (hdr1, ldr1)) in enumerate(zip(got, want)):
        assert H.bits_equal(hdr, hdr1).all() and H.bits_equal(ldr, ldr1).all(), k


def test_waves_left_in_flight_state_machine_on_the_cpu_build():
    """yc_render_wave_async / yc_wave_sync in the CPU build (nothing overlaps there, but the bookkeeping is the same):
    waves issued without waiting, a synchronous call in the middle, statistics read while waves are pending."""
    size = 48
    cam = H.scene_camera("cornell")
    sc = Y.Scene(H.scene_file("cornell"))
    c = Y.make_camera(size, size, cam["focal"], cam["fnum"], cam["pos"], cam["target"])
    res = []
    for mode in ("sync", "async"):
        ctx = Y.Context()
        ctx.upload_scene(sc)
        ctx.set_camera(c)
        ctx.begin_frame(size, size, 8, 16, (0, 0, 0), Y.TONEMAP_AGX)
        f = ctx.render_wave_async if mode == "async" else ctx.render_wave
        f(0, 2, 0)
        f(2, 2, 2)
        mid = ctx.stats().raysReference  # settles what is pending
        f(4, 1, 4, rect=(0, 0, 20, size))
        f(4, 1, 4, rect=(20, 0, size - 20, size))
        f(5, 3, 5)
        hdr, ldr, st = ctx.resolve()
        res.append((hdr, ldr, mid, st.raysReference, st.gpuMs))
        ctx.close()
    assert H.bits_equal(res[0][0], res[1][0]).all() and H.bits_equal(res[0][1], res[1][1]).all()
    assert res[0][2:4] == res[1][2:4] and res[1][4] > 0
