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
