"""oracle/port.py — TEST INFRASTRUCTURE ONLY (never imported by yart_b200/).

An independent numpy restatement of the integer-exact pieces of the render path, written from the
reference's sources (not from the CUDA code), so that tests can cross-check three ways: the unmodified
reference (oracle/_ref/oracle_ref), this port, and the product.  It is pinned against the reference's
recorded outputs in tests/golden (tests/test_oracle_port.py).  The floating-point bulk of the path
(traversal, BSDF, lights, integrator) has no port: the reference itself compiles here and is the oracle.

  SobolSampler<FastOwenScrambler>   src/core/sampler.hpp:71-174, src/core/scrambler.hpp:53-69,
                                    src/core/rng.hpp:25-100, src/math/math.hpp:102-134, math_base.hpp:156-170
  Mean / MoN / GMoN estimators      src/core/estimator.hpp:29-198
  wave schedule                     src/cpu/tile-renderer.hpp:120-124, 264-288
"""
from __future__ import annotations

import os
import struct

import numpy as np

_M64 = (1 << 64) - 1
_M32 = (1 << 32) - 1

# permutations[24][4], sampler.hpp:116-141
PERMUTATIONS = [
    (0, 1, 2, 3), (0, 1, 3, 2), (0, 2, 1, 3), (0, 2, 3, 1), (0, 3, 2, 1), (0, 3, 1, 2),
    (1, 0, 2, 3), (1, 0, 3, 2), (1, 2, 0, 3), (1, 2, 3, 0), (1, 3, 2, 0), (1, 3, 0, 2),
    (2, 1, 0, 3), (2, 1, 3, 0), (2, 0, 1, 3), (2, 0, 3, 1), (2, 3, 0, 1), (2, 3, 1, 0),
    (3, 1, 2, 0), (3, 1, 0, 2), (3, 2, 1, 0), (3, 2, 0, 1), (3, 0, 2, 1), (3, 0, 1, 2),
]


def sobol_dim1_matrix() -> list[int]:
    """sobol::matrices[52..103] (dimension 1), from the table dump of the reference build."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "yart_b200", "data", "tables.bin")
    raw = open(path, "rb").read()
    return list(struct.unpack_from("<52I", raw, 14112 * 4 + 52 * 4))


def log2_int(v: float) -> int:
    """math_base.hpp:156-160: exponent, +1 when the significand is >= sqrt(2)'s."""
    v = np.float32(v)
    if v < 1:
        return -log2_int(np.float32(1) / v)
    bits = int(np.float32(v).view(np.uint32))
    return ((bits >> 23) - 127) + (1 if (bits & ((1 << 23) - 1)) >= 0b00000000001101010000010011110011 else 0)


def round_up_pow2(v: int) -> int:
    v -= 1
    for s in (1, 2, 4, 8, 16):
        v |= v >> s
    return v + 1


def murmur64a_u32(key: int) -> int:
    """hash(uint32) = MurmurHash64A over the 4 key bytes, seed 0 (rng.hpp:25-91)."""
    m, r = 0xc6a4a7935bd1e995, 47
    h = (0 ^ (4 * m)) & _M64
    # len 4: no 8-byte blocks; tail switch: case 4..1 fold the bytes, then h *= m
    for i in (3, 2, 1, 0):
        h ^= ((key >> (8 * i)) & 0xff) << (8 * i)
    h = (h * m) & _M64
    h ^= h >> r
    h = (h * m) & _M64
    h ^= h >> r
    return h


def mix_bits(v: int) -> int:
    v ^= v >> 31
    v = (v * 0x7fb5d329728ea185) & _M64
    v ^= v >> 27
    v = (v * 0x81dadef4bc2dd44d) & _M64
    v ^= v >> 33
    return v


def left_shift2(x: int) -> int:
    x &= 0xffffffff
    x = (x ^ (x << 16)) & 0x0000ffff0000ffff
    x = (x ^ (x << 8)) & 0x00ff00ff00ff00ff
    x = (x ^ (x << 4)) & 0x0f0f0f0f0f0f0f0f
    x = (x ^ (x << 2)) & 0x3333333333333333
    x = (x ^ (x << 1)) & 0x5555555555555555
    return x


def encode_morton2(x: int, y: int) -> int:
    return (left_shift2(y) << 1) | left_shift2(x)


def reverse_bits32(n: int) -> int:
    return int(f"{n & _M32:032b}"[::-1], 2)


def fast_owen(v: int, seed: int) -> int:
    """FastOwenScrambler, scrambler.hpp:53-69."""
    v = reverse_bits32(v)
    v ^= (v * 0x3d20adea) & _M32
    v = (v + seed) & _M32
    v = (v * ((seed >> 16) | 1)) & _M32
    v ^= (v * 0x05526c56) & _M32
    v ^= (v * 0x53a22864) & _M32
    return reverse_bits32(v)


class SobolSampler:
    def __init__(self, spp: int, render_size=(64, 64)):
        self.log2spp = log2_int(float(spp))
        res = round_up_pow2(int(max(render_size)))
        self.n_base4 = log2_int(float(res)) + (self.log2spp + 1) // 2
        self.dim = 0
        self.morton = 0
        self.m1 = sobol_dim1_matrix()

    def start_pixel_sample(self, x: int, y: int, sample: int):
        self.dim = 0
        self.morton = ((encode_morton2(x, y) << self.log2spp) | sample) & _M64

    def sample_index(self) -> int:
        index = 0
        pow2 = self.log2spp & 1
        last = 1 if pow2 else 0
        dim_mix = (0x55555555 * self.dim) & _M32
        for i in range(self.n_base4 - 1, last - 1, -1):
            shift = 2 * i - last
            digit = (self.morton >> shift) & 3
            higher = self.morton >> (shift + 2)
            p = (mix_bits(higher ^ dim_mix) >> 24) % 24
            index |= PERMUTATIONS[p][digit] << shift
        if pow2:
            digit = self.morton & 1
            index |= digit ^ (mix_bits((self.morton >> 1) ^ dim_mix) & 1)
        return index

    @staticmethod
    def _to_float(v: int) -> np.float32:
        return min(np.float32(v) * np.float32(2.0 ** -32), np.float32(float.fromhex("0x1.fffffep-1")))

    def _sobol(self, index: int, dim: int) -> int:
        if dim == 0:
            return reverse_bits32(index & _M32)
        v, i = 0, 0
        while index:
            if index & 1:
                v ^= self.m1[i]
            index >>= 1
            i += 1
        return v

    def get1d(self) -> np.float32:
        idx = self.sample_index()
        self.dim += 1
        h = murmur64a_u32(self.dim) & _M32
        return self._to_float(fast_owen(self._sobol(idx, 0), h))

    def get2d(self):
        idx = self.sample_index()
        self.dim += 2
        h = murmur64a_u32(self.dim)
        return (self._to_float(fast_owen(self._sobol(idx, 0), h & _M32)),
                self._to_float(fast_owen(self._sobol(idx, 1), h >> 32)))


# ---- estimators (float32 arithmetic in the reference's order) -----------------------------------------
f32 = np.float32
LW = (f32(0.2126), f32(0.7152), f32(0.0722))


def luma(v) -> np.float32:
    return f32(f32(f32(v[0] * LW[0]) + f32(v[1] * LW[1])) + f32(v[2] * LW[2]))


def n_buckets(n: int, m_max: int = 15) -> int:
    return min(m_max, max(1, 1 + 2 * int((n - 5) / 10)))  # C++ integer division truncates toward zero


def estimate(samples: np.ndarray, kind: str) -> np.ndarray:
    """samples: (n, 3) float32 in the order they are added.  kind in {"gmon", "mon", "mean"}."""
    n = len(samples)
    with np.errstate(all="ignore"):
        if kind == "mean":
            acc = np.zeros(3, f32)
            for s in samples:
                if not np.isnan(s).any():
                    acc = (acc + s).astype(f32)
            return (acc / f32(n)).astype(f32)
        m = n_buckets(n)
        acc = np.zeros((m, 3), f32)
        cnt = np.zeros(m, np.int64)
        for i, s in enumerate(samples):
            ok = not np.isnan(s).any()
            if kind == "gmon":
                ok = ok and bool((s >= 0).all())
            if ok:
                acc[i % m] = (acc[i % m] + s).astype(f32)
                cnt[i % m] += 1
        if m == 1:
            return (acc[0] / f32(cnt[0])).astype(f32)
        acc = np.stack([(acc[i] / f32(cnt[i])).astype(f32) for i in range(m)])
        # std::sort on <= 16 elements == libstdc++ insertion sort; replay it so NaN lumas land identically
        a = [acc[i].copy() for i in range(m)]
        for i in range(1, m):
            val = a[i]
            if luma(val) < luma(a[0]):
                a[1:i + 1] = a[0:i]
                a[0] = val
            else:
                j = i
                while luma(val) < luma(a[j - 1]):
                    a[j] = a[j - 1]
                    j -= 1
                a[j] = val
        if kind == "mon":
            return a[m // 2]
        total = np.zeros(3, f32)
        weighted = np.zeros(3, f32)
        for i in range(m):
            total = (total + a[i]).astype(f32)
            weighted = (weighted + (a[i] * f32(i + 1)).astype(f32)).astype(f32)
        g = f32(f32(f32(2.0) * luma(weighted)) / f32(f32(m) * luma(total))) - f32(f32(m + 1) / f32(m))
        g = f32(g)
        if g > 1:
            g = f32(1)
        gc = f32(g * f32(m // 2))
        c = 0 if (np.isnan(gc) or gc < 0) else int(gc)  # size_t(NaN) wraps to "all buckets", see csrc/estimator.cuh
        s2 = np.zeros(3, f32)
        for i in range(c, m - c):
            s2 = (s2 + a[i]).astype(f32)
        return (s2 / f32(m - 2 * c)).astype(f32)


def wave_schedule(samples: int, first: int, max_wave: int) -> list[int]:
    """TileRenderer::renderImpl + finishTile: the sequence of wave sizes."""
    waves, wave, remaining, k = [], min(first, samples), samples, 0
    while wave > 0:
        waves.append(wave)
        remaining -= wave
        nxt = min(wave * 2, max_wave) if (k > 0 or wave > 1) else 1
        wave = min(nxt, remaining)
        k += 1
    return waves
