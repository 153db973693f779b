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
  uint32_t scrambler;     // the R of SobolSampler<R>: 0 FastOwenScrambler (the measured path), 1 OwenScrambler, 2 BinaryPermuteScrambler
};
enum : uint32_t { kScrambleFastOwen = 0, kScrambleOwen = 1, kScrambleBinaryPermute = 2 };

// permutations[24][4], sampler.hpp:116-141 (the 24 permutations of {0,1,2,3} in the order the
// reference lists them), packed 2 bits per digit: entry p, digit d → (kPermPacked[p] >> (2*d)) & 3.
// On the device the table is one byte per (p, digit) in GLOBAL memory, read through the read-only
// path: the index differs per lane, and a __constant__ table would serialise a warp's lookup into one
// pass per distinct address (up to 24), while these 96 bytes are three L1 sectors.
#define YB_PERM_ROW(b) (b) & 3, ((b) >> 2) & 3, ((b) >> 4) & 3, ((b) >> 6) & 3
YB_TABLE uint8_t kPerm[96] = {
  YB_PERM_ROW(0xE4), YB_PERM_ROW(0xB4), YB_PERM_ROW(0xD8), YB_PERM_ROW(0x78), YB_PERM_ROW(0x6C), YB_PERM_ROW(0x9C),
  YB_PERM_ROW(0xE1), YB_PERM_ROW(0xB1), YB_PERM_ROW(0xC9), YB_PERM_ROW(0x39), YB_PERM_ROW(0x2D), YB_PERM_ROW(0x8D),
  YB_PERM_ROW(0xC6), YB_PERM_ROW(0x36), YB_PERM_ROW(0xD2), YB_PERM_ROW(0x72), YB_PERM_ROW(0x4E), YB_PERM_ROW(0x1E),
  YB_PERM_ROW(0x27), YB_PERM_ROW(0x87), YB_PERM_ROW(0x1B), YB_PERM_ROW(0x4B), YB_PERM_ROW(0x63), YB_PERM_ROW(0x93)};
#undef YB_PERM_ROW

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

// scrambler.hpp:71-85 (OwenScrambler): one hashed flip decision per bit, each depending on the bits above it
YB_DEV uint32_t owenScramble(uint32_t v, uint32_t seed) {
  if (seed & 1u) v ^= 1u << 31;
  for (uint32_t b = 1; b < 32; b++) {
    const uint32_t mask = (~0u) << (32 - b);
    if (uint32_t(mixBits(uint64_t(v & mask)) ^ uint64_t(seed)) & (1u << b)) v ^= 1u << (31 - b);
  }
  return v;
}

// The scramblers other than FastOwen, out of line: they are off the measured path and must not grow every draw site.
YB_DEV_NI uint32_t scrambleOther(uint32_t v, uint32_t seed, uint32_t scrambler) {
  if (scrambler == kScrambleOwen) return owenScramble(v, seed);
  return seed ^ v;  // BinaryPermuteScrambler, scrambler.hpp:35-46
}

// Sobol dimension 1 (sobol.tables entries 52..103): generator-matrix column i is the Pascal-mod-2
// column v_0 = 2^31, v_i = v_{i-1} ^ (v_{i-1} >> 1) = (1 + S)^i v_0 (S = shift right by one), repeating
// with period 32 in the table.  Bit (31 - j) of column i is C(i, j) mod 2 = [j ⊆ i] (Lucas), so the
// XOR of the columns selected by the bits of `d` (sampler.hpp:143-153) is the superset-sum transform
// over GF(2) of those bits — five butterfly steps — followed by a bit reversal.
YB_DEV uint32_t sobolDim1Closed(uint64_t d) {
  uint32_t x = uint32_t(d) ^ uint32_t(d >> 32);  // columns repeat with period 32
  x ^= (x >> 1) & 0x55555555u;
  x ^= (x >> 2) & 0x33333333u;
  x ^= (x >> 4) & 0x0f0f0f0fu;
  x ^= (x >> 8) & 0x00ff00ffu;
  x ^= (x >> 16) & 0x0000ffffu;
  return reverseBits32(x);
}

struct Sampler {
  uint64_t morton;
  uint32_t dim;
  uint32_t log2spp, nBase4Digits, scrambler;

  YB_DEV void start(const SamplerConfig& c, uint32_t px, uint32_t py, uint32_t sample) {
    log2spp = c.log2spp;
    nBase4Digits = c.nBase4Digits;
    scrambler = c.scrambler;
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
      // (mixBits(..) >> 24) % 24 on the 40-bit quotient, in 32-bit pieces: 2^32 mod 24 = 16, hi < 256
      uint64_t mb = mixBits(higherDigits ^ dimMix) >> 24;
      uint32_t p = ((uint32_t(mb >> 32) * 16u) + (uint32_t(mb) % 24u)) % 24u;
      digit = __ldg(&kPerm[4u * p + digit]);
      index |= uint64_t(digit) << digitShift;
    }
    if (pow2Samples) {
      uint32_t digit = uint32_t(morton) & 1u;
      index |= uint64_t(digit ^ uint32_t(mixBits((morton >> 1) ^ dimMix) & 1ull));
    }
    return index;
  }

  // sampler.hpp:143-153, dimension 0: v = reverseBits32(uint32(d))
  YB_DEV float finish(uint32_t v, uint32_t seed) const {
    v = scrambler == kScrambleFastOwen ? fastOwen(v, seed) : scrambleOther(v, seed, scrambler);
    return fminf(float(v) * 0x1p-32f, 0x1.fffffep-1f);
  }
  static YB_DEV uint32_t sobolDim1(uint64_t d) { return sobolDim1Closed(d); }

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
