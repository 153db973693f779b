"""oracle/port.py (independent numpy restatement of the integer-exact pieces) against the reference's
recorded outputs — a third leg next to oracle/_ref and the product."""
import importlib.util
import os
import struct

import numpy as np
import pytest

import harness as H
import parity_common as PC

spec = importlib.util.spec_from_file_location("oracle_port", os.path.join(H.ROOT, "oracle", "port.py"))
port = importlib.util.module_from_spec(spec)
spec.loader.exec_module(port)


@pytest.mark.parametrize("tag", ["sampler16", "sampler128", "sampler1024", "sampler4096", "sampler3", "sampler12", "sampler100"])
def test_port_sampler_matches_reference(tag):
    g = PC.load(os.path.join(H.GOLDEN, f"kat_{tag}.npz"))
    blob, ref = g["blob"].tobytes(), g["out"].reshape(-1, 8)
    spp, n = struct.unpack_from("<II", blob, 0)
    rec = np.frombuffer(blob, np.uint32, n * 3, 8).reshape(n, 3)
    s = port.SobolSampler(spp)
    for i in range(0, n, 7):  # every 7th record keeps the pure-Python loop short
        x, y, smp = (int(v) for v in rec[i])
        s.start_pixel_sample(x, y, smp)
        a, b = s.get2d(), s.get2d()
        c, d = s.get1d(), s.get1d()
        e = s.get2d()
        got = np.array([*a, *b, c, d, *e], np.float32)
        assert np.array_equal(got.view(np.uint32), ref[i].view(np.uint32)), (tag, i)


def test_port_survey_constants():
    # SURVEY.md Appendix C
    assert port.murmur64a_u32(1) == 0xf52ab5e6fe56c909 and port.murmur64a_u32(2) == 0x038495654e7850ac
    assert port.mix_bits(1) == 0xccde22c1faa4d20f and port.encode_morton2(3, 5) == 39
    assert port.fast_owen(0x80000000, 0x12345678) == 0x87e32b06
    assert [port.log2_int(v) for v in (16, 1024, 4096, 100)] == [4, 10, 12, 7] and port.round_up_pow2(64) == 64


@pytest.mark.parametrize("tag", ["gmon3", "gmon16", "gmon64", "gmon128"])
def test_port_estimators_match_reference(tag):
    g = PC.load(os.path.join(H.GOLDEN, f"kat_{tag}.npz"))
    blob, ref = g["blob"].tobytes(), g["out"].reshape(-1, 9)
    n, npix = struct.unpack_from("<II", blob, 0)
    samples = np.frombuffer(blob, np.float32, npix * n * 3, 8).reshape(npix, n, 3)
    for i in range(0, npix, 5):
        for k, kind in enumerate(("gmon", "mon", "mean")):
            got = port.estimate(samples[i], kind)
            assert H.bits_equal(got, ref[i, 3 * k:3 * k + 3]).all(), (tag, i, kind, got, ref[i, 3 * k:3 * k + 3])


def test_port_wave_schedule():
    assert port.wave_schedule(4096, 64, 128) == [64] + [128] * 31 + [64]  # 33 waves
    assert port.wave_schedule(16, 4, 8) == [4, 8, 4] and port.wave_schedule(8, 1, 4) == [1, 1, 2, 4]
