// sampler.cuh — device SobolSampler<FastOwenScrambler> (pbrt-v4 ZSobol), integer exact.
//
// Follows reference src/core/sampler.hpp:71-174 (SobolSampler), src/core/scrambler.hpp:53-69
// (FastOwenScrambler), src/core/rng.hpp:25-100 (murmurHash64A / hash / mixBits) and
// src/math/math.hpp:102-134 (reverseBits32, encodeMorton2).  State per path is
// (mortonIndex: u64, dim: u32); the stream depends only on (pixel, sample index, dim, totalSpp,
// tileSize) — there is no seed.
#pragma once
#include "dmath.cuh"

namespace yb {

struct SamplerConfig {
  uint32_t log2spp;       // log2Int(float(totalSamples))
  uint32_t nBase4Digits;  // log2Int(roundUpPow2(tileSize)) + (log2spp + 1) / 2
};

// permutations[24][4], sampler.hpp:116-141 (the 24 permutations of {0,1,2,3} in the order the
// reference lists them), packed 2 bits per digit: entry p, digit d → (kPerm[p] >> (2*d)) & 3.
YB_CONST uint8_t kPerm[24] = {
  0xE4, 0xB4, 0xD8, 0x78, 0x6C, 0x9C, 0xE1, 0xB1, 0xC9, 0x39, 0x2D, 0x8D,
  0xC6, 0x36, 0xD2, 0x72, 0x4E, 0x1E, 0x27, 0x87, 0x1B, 0x4B, 0x63, 0x93};

// rng.hpp:93-100
YB_DEV uint64_t mixBits(uint64_t v) {
  v ^= (v >> 31);
  v *= 0x7fb5d329728ea185ull;
  v ^= (v >> 27);
  v *= 0x81dadef4bc2dd44dull;
  v ^= (v >> 33);
  return v;
}

// hash(uint32 dim): MurmurHash64A over the 4 bytes of `dim`, seed 0 (rng.hpp:25-91).
// len = 4 → no 8-byte blocks; the tail switch folds bytes 3..0, i.e. h ^= dim; h *= m.
YB_DEV uint64_t hashDim(uint32_t dim) {
  const uint64_t m = 0xc6a4a7935bd1e995ull;
  uint64_t h = 0ull ^ (4ull * m);
  h ^= uint64_t(dim);
  h *= m;
  h ^= h >> 47;
  h *= m;
  h ^= h >> 47;
  return h;
}

// math.hpp:102-109 (== __brev)
YB_DEV uint32_t reverseBits32(uint32_t n) { return __brev(n); }

// math.hpp:122-134
YB_DEV uint64_t leftShift2(uint64_t x) {
  x &= 0xffffffffull;
  x = (x ^ (x << 16)) & 0x0000ffff0000ffffull;
  x = (x ^ (x << 8)) & 0x00ff00ff00ff00ffull;
  x = (x ^ (x << 4)) & 0x0f0f0f0f0f0f0f0full;
  x = (x ^ (x << 2)) & 0x3333333333333333ull;
  x = (x ^ (x << 1)) & 0x5555555555555555ull;
  return x;
}
YB_DEV uint64_t encodeMorton2(uint32_t x, uint32_t y) { return (leftShift2(y) << 1) | leftShift2(x); }

// scrambler.hpp:53-69
YB_DEV uint32_t fastOwen(uint32_t v, uint32_t seed) {
  v = reverseBits32(v);
  v ^= v * 0x3d20adeau;
  v += seed;
  v *= (seed >> 16) | 1u;
  v ^= v * 0x05526c56u;
  v ^= v * 0x53a22864u;
  return reverseBits32(v);
}

// Sobol dimension 1 generator matrix column i (sobol.tables entries 52..103): the Pascal-mod-2
// columns v_0 = 2^31, v_i = v_{i-1} ^ (v_{i-1} >> 1), repeating with period 32 in the table.
YB_DEV uint32_t sobolDim1Column(uint32_t i) {
  // closed form of the recurrence: bit-reversed row (i & 31) of Pascal's triangle mod 2
  uint32_t k = i & 31u;
  uint32_t v = 0x80000000u;
  // v_k = Π (1 + S)^k applied to v_0 where S is a right shift; use binary decomposition of k
  if (k & 1u) v ^= v >> 1;
  if (k & 2u) v ^= v >> 2;
  if (k & 4u) v ^= v >> 4;
  if (k & 8u) v ^= v >> 8;
  if (k & 16u) v ^= v >> 16;
  return v;
}

struct Sampler {
  uint64_t morton;
  uint32_t dim;
  uint32_t log2spp, nBase4Digits;

  YB_DEV void start(const SamplerConfig& c, uint32_t px, uint32_t py, uint32_t sample) {
    log2spp = c.log2spp;
    nBase4Digits = c.nBase4Digits;
    dim = 0;
    morton = (encodeMorton2(px, py) << log2spp) | uint64_t(sample);  // sampler.hpp:84-87
  }

  // sampler.hpp:155-173
  YB_DEV uint64_t sampleIndex() const {
    uint64_t index = 0;
    const bool pow2Samples = log2spp & 1u;
    const int lastDigit = pow2Samples ? 1 : 0;
    const uint64_t dimMix = uint64_t(0x55555555u * dim);
    for (int i = int(nBase4Digits) - 1; i >= lastDigit; i--) {
      uint32_t digitShift = 2 * i - lastDigit;
      uint32_t digit = uint32_t(morton >> digitShift) & 3u;
      uint64_t higherDigits = morton >> (digitShift + 2);
      // (mixBits(..) >> 24) % 24 on the 40-bit quotient, in 32-bit pieces: 2^32 mod 24 = 16
      uint64_t mb = mixBits(higherDigits ^ dimMix) >> 24;
      uint32_t p = (((uint32_t(mb >> 32) % 24u) * 16u) + (uint32_t(mb) % 24u)) % 24u;
      digit = (kPerm[p] >> (2 * digit)) & 3u;
      index |= uint64_t(digit) << digitShift;
    }
    if (pow2Samples) {
      uint32_t digit = uint32_t(morton) & 1u;
      index |= uint64_t(digit ^ uint32_t(mixBits((morton >> 1) ^ dimMix) & 1ull));
    }
    return index;
  }

  // sampler.hpp:143-153, dimension 0: v = reverseBits32(uint32(d))
  static YB_DEV float finish(uint32_t v, uint32_t seed) {
    v = fastOwen(v, seed);
    return fminf(float(v) * 0x1p-32f, 0x1.fffffep-1f);
  }
  static YB_DEV uint32_t sobolDim1(uint64_t d) {
    uint32_t v = 0;
    for (uint32_t i = 0; d != 0; d >>= 1, i++)
      if (d & 1ull) v ^= sobolDim1Column(i);
    return v;
  }

  // sampler.hpp:89-94: index uses the CURRENT dim, the hash the incremented one
  YB_DEV float get1D() {
    uint64_t idx = sampleIndex();
    dim++;
    uint32_t h = uint32_t(hashDim(dim));
    return finish(reverseBits32(uint32_t(idx)), h);
  }

  // sampler.hpp:96-107
  YB_DEV V2 get2D() {
    uint64_t idx = sampleIndex();
    dim += 2;
    uint64_t hb = hashDim(dim);
    return V2(finish(reverseBits32(uint32_t(idx)), uint32_t(hb)), finish(sobolDim1(idx), uint32_t(hb >> 32)));
  }
};

}  // namespace yb
