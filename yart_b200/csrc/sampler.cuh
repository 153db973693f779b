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
  uint32_t kind;          // the `Sampler` template argument: 0 SobolSampler<R>, 1 NaiveSampler, 2 StratifiedSampler
  uint32_t strata;        // StratifiedSampler: m_xSamples = m_ySamples = ceil(sqrt(samplesPerPixel)), sampler.hpp:49-51
  // One word for the per-lane copy (the persistent traversal kernels of alpha-tested scenes keep a Sampler per lane):
  // log2spp [0,6) | nBase4Digits [6,12) | scrambler [12,14) | kind [14,16) | strata [16,32)
  YB_DEV uint32_t packed() const {
    return log2spp | (nBase4Digits << 6) | (scrambler << 12) | (kind << 14) | (strata << 16);
  }
};
enum : uint32_t { kSamplerSobol = 0, kSamplerNaive = 1, kSamplerStratified = 2 };
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

// ---- NaiveSampler / StratifiedSampler (src/core/sampler.cpp:5-50) -------------------------------------------------
// Both draw from an xoshiro256++ generator seeded per pixel sample with hash(pixel, sample) and advance it once per
// 1-D value.  The path state carries no generator: the stream position equals the sampler dimension (1 per get1D,
// 2 per get2D, exactly the number of uniform() calls so far), so a draw re-seeds and skips `dim` outputs.  Off the
// measured path, kept out of line.

// hash(uint2 p, uint32_t v): MurmurHash64A over 12 bytes, seed 0 (rng.hpp:25-91): one 8-byte block, a 4-byte tail
YB_DEV uint64_t hashPixel(uint32_t px, uint32_t py, uint32_t v) {
  const uint64_t m = 0xc6a4a7935bd1e995ull;
  uint64_t h = 12ull * m;
  uint64_t k = uint64_t(px) | (uint64_t(py) << 32);
  k *= m;
  k ^= k >> 47;
  k *= m;
  h ^= k;
  h *= m;
  h ^= uint64_t(v);
  h *= m;
  h ^= h >> 47;
  h *= m;
  h ^= h >> 47;
  return h;
}

// xoshiro-rng/xoshiro.hpp:119-124
YB_DEV uint64_t splitmix64(uint64_t seed) {
  uint64_t z = seed + 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// Xoshiro::Xoshiro256PP (xoshiro.hpp:136-214) + RNG::uniform (rng.cpp:7-9): std::uniform_real_distribution<float>
// over a 64-bit generator is libstdc++'s generate_canonical<float, 24>: ONE draw, float(u64) / 2^64 (the conversion
// rounds to nearest, so it can reach 1), and a result >= 1 is replaced by nextafter(1, 0).
struct Xoshiro256pp {
  uint64_t s0, s1, s2, s3;
  YB_DEV void seed(uint64_t v) {
    s0 = splitmix64(splitmix64(v));
    s1 = splitmix64(s0);
    s2 = splitmix64(s1);
    s3 = splitmix64(s2);
  }
  static YB_DEV uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  YB_DEV uint64_t next() {
    const uint64_t result = rotl(s0 + s3, 23) + s0;
    const uint64_t t = s1 << 17;
    s2 ^= s0;
    s3 ^= s1;
    s1 ^= s2;
    s0 ^= s3;
    s2 ^= t;
    s3 = rotl(s3, 45);
    return result;
  }
  YB_DEV float uniform() {
    const float r = float(next()) * 0x1p-64f;
    return r >= 1.0f ? 0x1.fffffep-1f : r;
  }
};

// rng.hpp:103-133
YB_DEV uint32_t permel(uint32_t i, uint32_t l, uint32_t p) {
  uint32_t w = l - 1;
  w |= w >> 1;
  w |= w >> 2;
  w |= w >> 4;
  w |= w >> 8;
  w |= w >> 16;
  do {
    i ^= p;
    i *= 0xe170893du;
    i ^= p >> 16;
    i ^= (i & w) >> 4;
    i ^= p >> 8;
    i *= 0x0929eb3fu;
    i ^= p >> 23;
    i ^= (i & w) >> 1;
    i *= 1u | p >> 27;
    i *= 0x6935fa69u;
    i ^= (i & w) >> 11;
    i *= 0x74dcb303u;
    i ^= (i & w) >> 2;
    i *= 0x9e501cc3u;
    i ^= (i & w) >> 2;
    i *= 0xc860a3dfu;
    i &= w;
    i ^= i >> 5;
  } while (i >= l);
  return (i + p) % l;
}

// One get1D (two = false) or get2D of the RNG-based samplers at stream position `dim` of pixel sample (px, py, sample).
YB_DEV_NI V2 rngSamplerDraw(uint32_t px, uint32_t py, uint32_t sample, uint32_t dim, uint32_t kind, uint32_t strata, bool two) {
  Xoshiro256pp rng;
  rng.seed(hashPixel(px, py, sample));  // startPixelSample, sampler.cpp:5-7 / 21-26
  for (uint32_t k = 0; k < dim; k++) rng.next();
  const float d0 = rng.uniform();
  const float d1 = two ? rng.uniform() : 0.0f;
  if (kind == kSamplerNaive) return V2(d0, d1);  // sampler.cpp:9-15
  // StratifiedSampler, sampler.cpp:28-46
  const uint32_t n = strata * strata;
  const uint32_t stratum = permel(sample, n, uint32_t(hashPixel(px, py, dim)));
  if (!two) return V2((float(stratum) + d0) / float(n), 0.0f);
  const uint32_t x = stratum % strata, y = stratum / strata;
  return V2((float(x) + d0) / float(strata), (float(y) + d1) / float(strata));
}

struct Sampler {
  uint64_t morton;
  uint32_t dim;
  uint32_t cfg;  // SamplerConfig::packed()
  YB_DEV uint32_t log2spp() const { return cfg & 63u; }
  YB_DEV uint32_t nBase4Digits() const { return (cfg >> 6) & 63u; }
  // Every `Sampler` template argument other than SobolSampler<FastOwenScrambler> (the Owen / BinaryPermute scramblers,
  // NaiveSampler, StratifiedSampler) lives in the YB_RNG_SAMPLERS build of the library (libyart_b200_samplers.so,
  // same ABI): in the default build kind() and scrambler() are constants and every branch on them folds away —
  // measured, even a never-taken call or an inlined cold path in the persistent traversal kernels of alpha-tested
  // scenes costs 4-11 % of a Sponza-shaped step (the scrambler switch alone 1-2 % of a McLaren-shaped one).
#ifdef YB_RNG_SAMPLERS
  YB_DEV uint32_t kind() const { return (cfg >> 14) & 3u; }
  YB_DEV uint32_t scrambler() const { return (cfg >> 12) & 3u; }
#else
  YB_DEV uint32_t kind() const { return kSamplerSobol; }
  YB_DEV uint32_t scrambler() const { return kScrambleFastOwen; }
#endif
  YB_DEV uint32_t strata() const { return cfg >> 16; }

  YB_DEV void start(const SamplerConfig& c, uint32_t px, uint32_t py, uint32_t sample) {
    cfg = c.packed();
    dim = 0;
    if (kind() == kSamplerSobol) morton = (encodeMorton2(px, py) << c.log2spp) | uint64_t(sample);  // sampler.hpp:84-87
    else morton = (uint64_t(px | (py << 16)) << 32) | uint64_t(sample);  // RNG samplers: the pixel sample itself
  }
  YB_DEV V2 drawOther(bool two) const {
    const uint32_t pix = uint32_t(morton >> 32);
    return rngSamplerDraw(pix & 0xffffu, pix >> 16, uint32_t(morton), dim, kind(), strata(), two);
  }

  // sampler.hpp:155-173
  YB_DEV uint64_t sampleIndex() const {
    uint64_t index = 0;
    const bool pow2Samples = log2spp() & 1u;
    const int lastDigit = pow2Samples ? 1 : 0;
    const uint64_t dimMix = uint64_t(0x55555555u * dim);
    for (int i = int(nBase4Digits()) - 1; i >= lastDigit; i--) {
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
    v = scrambler() == kScrambleFastOwen ? fastOwen(v, seed) : scrambleOther(v, seed, scrambler());
    return fminf(float(v) * 0x1p-32f, 0x1.fffffep-1f);
  }
  static YB_DEV uint32_t sobolDim1(uint64_t d) { return sobolDim1Closed(d); }

  // sampler.hpp:89-94: index uses the CURRENT dim, the hash the incremented one
  YB_DEV float get1D() {
    if (kind() != kSamplerSobol) {
      const V2 r = drawOther(false);
      dim++;
      return r.x;
    }
    uint64_t idx = sampleIndex();
    dim++;
    uint32_t h = uint32_t(hashDim(dim));
    return finish(reverseBits32(uint32_t(idx)), h);
  }

  // sampler.hpp:96-107
  YB_DEV V2 get2D() {
    if (kind() != kSamplerSobol) {
      const V2 r = drawOther(true);
      dim += 2;
      return r;
    }
    uint64_t idx = sampleIndex();
    dim += 2;
    uint64_t hb = hashDim(dim);
    return V2(finish(reverseBits32(uint32_t(idx)), uint32_t(hb)), finish(sobolDim1(idx), uint32_t(hb >> 32)));
  }
};

}  // namespace yb
