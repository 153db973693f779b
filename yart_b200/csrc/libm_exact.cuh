// libm_exact.cuh — sinf / cosf / logf / expf that round like the oracle's libm.
//
// The reference calls std::sin/cos/log/exp on floats (src/math/sampling.hpp:24-44, 83-84;
// src/bsdf/parametric.cpp:837).  The oracle is the reference built against glibc 2.39, whose
// float routines (sysdeps/ieee754/flt-32/{s_sinf,s_cosf,e_logf,e_expf}.c, from ARM's
// optimized-routines) evaluate short polynomials in DOUBLE precision and round once to float.
// CUDA's libm uses other algorithms (1-2 ulp), and a last-bit change of a sampled direction sends
// a path through different discrete choices (lobe pick, roulette, alpha test) — single pixels then
// differ by whole samples.  Restating the published algorithms in double precision makes the GPU
// path follow the oracle's paths.  The constants are the published ones; tests/hostsim/libm_check
// compares every function with the host's glibc over every float of the ranges the path uses.
// (glibc's x86-64 ifunc variants compile the same C with FMA contraction; the double intermediates
// then differ by < 1e-16 relative, which changes the rounded float about once in 1e8 calls.)
#pragma once
#include "dmath.cuh"

namespace yb {

#ifdef YB_HOSTSIM
YB_DEV double asDouble(uint64_t u) {
  double d;
  memcpy(&d, &u, 8);
  return d;
}
YB_DEV uint64_t asU64(double d) {
  uint64_t u;
  memcpy(&u, &d, 8);
  return u;
}
#else
YB_DEV double asDouble(uint64_t u) { return __longlong_as_double((long long)u); }
YB_DEV uint64_t asU64(double d) { return (uint64_t)__double_as_longlong(d); }
#endif

// no FMA contraction in these routines either (the non-multiarch C source is the specification)
#ifdef YB_HOSTSIM
YB_DEV double dmul(double a, double b) { return a * b; }
YB_DEV double dadd(double a, double b) { return a + b; }
#else
YB_DEV double dmul(double a, double b) { return __dmul_rn(a, b); }
YB_DEV double dadd(double a, double b) { return __dadd_rn(a, b); }
#endif

namespace libm {

// __sincosf_table[0] (s_sincosf_data.c); table[1] negates c0..c4
constexpr double kHpiInv = 0x1.45F306DC9C883p+23, kHpi = 0x1.921FB54442D18p0;
constexpr double kC0 = 0x1p0, kC1 = -0x1.ffffffd0c621cp-2, kC2 = 0x1.55553e1068f19p-5, kC3 = -0x1.6c087e89a359dp-10,
                 kC4 = 0x1.99343027bf8c3p-16;
constexpr double kS1 = -0x1.555545995a603p-3, kS2 = 0x1.1107605230bc4p-7, kS3 = -0x1.994eb3774cf24p-13;

YB_DEV uint32_t abstop12(float x) { return (__float_as_uint(x) >> 20) & 0x7ff; }

// sinf_poly (s_sincosf.h): n odd → cosine polynomial; `neg` = use table[1] (cosine coefficients negated)
YB_DEV float sincosPoly(double x, double x2, bool neg, int n) {
  if ((n & 1) == 0) {
    const double x3 = dmul(x, x2);
    const double s1 = dadd(kS2, dmul(x2, kS3));
    const double x5 = dmul(x3, x2);
    const double s = dadd(x, dmul(x3, kS1));
    return float(dadd(s, dmul(x5, s1)));
  }
  const double sg = neg ? -1.0 : 1.0;
  const double x4 = dmul(x2, x2);
  const double c2 = dadd(sg * kC3, dmul(x2, sg * kC4));
  const double c1 = dadd(sg * kC0, dmul(x2, sg * kC1));
  const double x6 = dmul(x4, x2);
  const double c = dadd(c1, dmul(x4, sg * kC2));
  return float(dadd(c, dmul(x6, c2)));
}

// reduce_fast: valid for |x| < 120
YB_DEV double reduceFast(double x, int& n) {
  const double r = dmul(x, kHpiInv);
  n = (int(r) + 0x800000) >> 24;
  return dadd(x, -dmul(double(n), kHpi));
}

}  // namespace libm

// s_sinf.c.  |y| >= 120, inf and NaN do not occur on the path (arguments are 2*pi*u, u in [0,1));
// they fall back to the platform sinf.
YB_DEV float sinfExact(float y) {
  using namespace libm;
  double x = y;
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {  // pio4f
    const double s = dmul(x, x);
    if (abstop12(y) < abstop12(0x1p-12f)) return y;
    return sincosPoly(x, s, false, 0);
  }
  if (abstop12(y) < abstop12(120.0f)) {
    int n;
    x = reduceFast(x, n);
    const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;  // sign[n & 3] = {1,-1,-1,1}
    return sincosPoly(dmul(x, s), dmul(x, x), (n & 2) != 0, n);
  }
  return sinf(y);
}

// s_cosf.c
YB_DEV float cosfExact(float y) {
  using namespace libm;
  double x = y;
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
    const double x2 = dmul(x, x);
    if (abstop12(y) < abstop12(0x1p-12f)) return 1.0f;
    return sincosPoly(x, x2, false, 1);
  }
  if (abstop12(y) < abstop12(120.0f)) {
    int n;
    x = reduceFast(x, n);
    const int m = (n + 1) & 3;
    const double s = (m == 1 || m == 2) ? -1.0 : 1.0;
    return sincosPoly(dmul(x, s), dmul(x, x), ((n + 1) & 2) != 0, n ^ 1);
  }
  return cosf(y);
}

namespace libm {
// __logf_data (e_logf_data.c): 16 x {invc, logc}
YB_TABLE double kLogTab[32] = {
  0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2,
  0x1.49539f0f010bp+0,  -0x1.01eae7f513a67p-2, 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3,
  0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8eap+0,  -0x1.1aa2bc79c81p-3,
  0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4,
  0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5, 0x1p+0,               0x0p+0,
  0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5,  0x1.ca4b31f026aap-1,  0x1.c5e53aa362eb4p-4,
  0x1.b2036576afce6p-1, 0x1.526e57720db08p-3,  0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3,
  0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2,  0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2};
constexpr double kLn2 = 0x1.62e42fefa39efp-1;
constexpr double kLogA0 = -0x1.00ea348b88334p-2, kLogA1 = 0x1.5575b0be00b6ap-2, kLogA2 = -0x1.ffffef20a4123p-2;
}  // namespace libm

// e_logf.c.  x <= 0, inf, NaN: platform logf (u.x == 0 gives -inf there as in the reference).
YB_DEV float logfExact(float x) {
  using namespace libm;
  uint32_t ix = __float_as_uint(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2 == 0 || ix == 0x7f800000u || (ix & 0x80000000u) || ix * 2 >= 0xff000000u) return logf(x);
    ix = __float_as_uint(x * 0x1p23f);  // subnormal: normalise
    ix -= 23u << 23;
  }
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = int((tmp >> (23 - 4)) % 16u);
  const int k = int32_t(tmp) >> 23;
  const uint32_t iz = ix - (tmp & 0xff800000u);
  const double invc = kLogTab[2 * i], logc = kLogTab[2 * i + 1];
  const double z = double(__uint_as_float(iz));
  const double r = dadd(dmul(z, invc), -1.0);
  const double y0 = dadd(logc, dmul(double(k), kLn2));
  const double r2 = dmul(r, r);
  double y = dadd(dmul(kLogA1, r), kLogA2);
  y = dadd(dmul(kLogA0, r2), y);
  y = dadd(dmul(y, r2), dadd(y0, r));
  return float(y);
}

namespace libm {
// __exp2f_data.tab (e_exp2f_data.c): asuint64(2^(i/32)) - (i << 47), i = 0..31
YB_TABLE uint64_t kExp2Tab[32] = {
  0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, 0x3fef72b83c7d517bull,
  0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, 0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull,
  0x3feedea64c123422ull, 0x3feece086061892dull, 0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull,
  0x3feea47eb03a5585ull, 0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
  0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, 0x3feee89f995ad3adull,
  0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, 0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full,
  0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};
constexpr double kExpShift = 0x1.8p+52;
constexpr double kInvLn2N = 0x1.71547652b82fep+0 * 32.0;
constexpr double kExpC0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0, kExpC1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0,
                 kExpC2 = 0x1.62e42ff0c52d6p-1 / 32.0;
}  // namespace libm

// e_expf.c.  |x| >= 88 (overflow / underflow handling), NaN: platform expf.
YB_DEV float expfExact(float x) {
  using namespace libm;
  const uint32_t abstop = (__float_as_uint(x) >> 20) & 0x7ff;
  if (abstop >= ((__float_as_uint(88.0f) >> 20) & 0x7ff)) return expf(x);
  const double xd = double(x);
  const double z = dmul(kInvLn2N, xd);
  double kd = dadd(z, kExpShift);
  const uint64_t ki = asU64(kd);
  kd = dadd(kd, -kExpShift);
  const double r = dadd(z, -kd);
  uint64_t t = kExp2Tab[ki % 32u];
  t += ki << (52 - 5);
  const double s = asDouble(t);
  const double zz = dadd(dmul(kExpC0, r), kExpC1);
  const double r2 = dmul(r, r);
  double y = dadd(dmul(kExpC2, r), 1.0);
  y = dadd(dmul(zz, r2), y);
  y = dmul(y, s);
  return float(y);
}

namespace libm {
// __log2f_data (e_log2f_data.c): 16 x {invc, log2(c)}; __powf_log2_data.tab holds the same pairs
YB_TABLE double kLog2Tab[32] = {
  0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2, 0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2,
  0x1.49539f0f010bp+0,  -0x1.7418b0a1fb77bp-2, 0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2,
  0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2, 0x1.25e227b0b8eap+0,  -0x1.97c1d1b3b7afp-3,
  0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3, 0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4,
  0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5, 0x1p+0,               0x0p+0,
  0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4,  0x1.ca4b31f026aap-1,  0x1.476a9543891bap-3,
  0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3,  0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2,
  0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2,  0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2};
constexpr double kLog2A0 = -0x1.712b6f70a7e4dp-2, kLog2A1 = 0x1.ecabf496832ep-2, kLog2A2 = -0x1.715479ffae3dep-1,
                 kLog2A3 = 0x1.715475f35c8b8p0;
// __powf_log2_data.poly (POWF_SCALE = 1)
constexpr double kPowA0 = 0x1.27616c9496e0bp-2, kPowA1 = -0x1.71969a075c67ap-2, kPowA2 = 0x1.ec70a6ca7baddp-2,
                 kPowA3 = -0x1.7154748bef6c8p-1, kPowA4 = 0x1.71547652ab82bp0;
// __exp2f_data.poly and shift_scaled (= 0x1.8p+52 / 32)
constexpr double kExp2C0 = 0x1.c6af84b912394p-5, kExp2C1 = 0x1.ebfce50fac4f3p-3, kExp2C2 = 0x1.62e42ff0c52d6p-1;
constexpr double kExp2ShiftScaled = 0x1.8p+52 / 32.0;
}  // namespace libm

// e_log2f.c.  x <= 0, inf, NaN: platform log2f (same IEEE special values).
YB_DEV float log2fExact(float x) {
  using namespace libm;
  uint32_t ix = __float_as_uint(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2 == 0 || ix == 0x7f800000u || (ix & 0x80000000u) || ix * 2 >= 0xff000000u) return log2f(x);
    ix = __float_as_uint(x * 0x1p23f);
    ix -= 23u << 23;
  }
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = int((tmp >> (23 - 4)) % 16u);
  const uint32_t top = tmp & 0xff800000u;
  const uint32_t iz = ix - top;
  const int k = int32_t(tmp) >> 23;
  const double invc = kLog2Tab[2 * i], logc = kLog2Tab[2 * i + 1];
  const double z = double(__uint_as_float(iz));
  const double r = dadd(dmul(z, invc), -1.0);
  const double y0 = dadd(logc, double(k));
  const double r2 = dmul(r, r);
  double y = dadd(dmul(kLog2A1, r), kLog2A2);
  y = dadd(dmul(kLog2A0, r2), y);
  const double p = dadd(dmul(kLog2A3, r), y0);
  y = dadd(dmul(y, r2), p);
  return float(y);
}

// e_powf.c for finite non-zero x and finite non-zero y; zeros, infinities and NaNs take the platform
// powf (IEEE 754 special values, identical in both libraries).
YB_DEV float powfExact(float x, float y) {
  using namespace libm;
  uint32_t ix = __float_as_uint(x);
  const uint32_t iy = __float_as_uint(y);
  uint64_t signBias = 0;
  const bool yZeroInfNan = 2u * iy - 1u >= 2u * 0x7f800000u - 1u;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u || yZeroInfNan) {
    const bool xZeroInfNan = 2u * ix - 1u >= 2u * 0x7f800000u - 1u;
    if (yZeroInfNan || xZeroInfNan) return powf(x, y);
    if (ix & 0x80000000u) {
      // finite x < 0: checkint(iy)
      const int e = int(iy >> 23) & 0xff;
      int yint;
      if (e < 0x7f) yint = 0;
      else if (e > 0x7f + 23) yint = 2;
      else if (iy & ((1u << (0x7f + 23 - e)) - 1u)) yint = 0;
      else if (iy & (1u << (0x7f + 23 - e))) yint = 1;
      else yint = 2;
      if (yint == 0) return powf(x, y);  // NaN (invalid)
      if (yint == 1) signBias = 1ull << (5 + 11);
      ix &= 0x7fffffffu;
    }
    if (ix < 0x00800000u) {
      ix = __float_as_uint(x * 0x1p23f);
      ix &= 0x7fffffffu;
      ix -= 23u << 23;
    }
  }
  // log2_inline
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = int((tmp >> (23 - 4)) % 16u);
  const uint32_t top = tmp & 0xff800000u;
  const uint32_t iz = ix - top;
  const int k = int32_t(top) >> 23;
  const double invc = kLog2Tab[2 * i], logc = kLog2Tab[2 * i + 1];
  const double z = double(__uint_as_float(iz));
  const double r = dadd(dmul(z, invc), -1.0);
  const double y0 = dadd(logc, double(k));
  const double r2 = dmul(r, r);
  double yy = dadd(dmul(kPowA0, r), kPowA1);
  const double p = dadd(dmul(kPowA2, r), kPowA3);
  const double r4 = dmul(r2, r2);
  double q = dadd(dmul(kPowA4, r), y0);
  q = dadd(dmul(p, r2), q);
  yy = dadd(dmul(yy, r4), q);
  const double ylogx = dmul(double(y), yy);
  if (((asU64(ylogx) >> 47) & 0xffff) >= (asU64(126.0) >> 47)) {
    if (ylogx > 0x1.fffffffd1d571p+6) return signBias ? -INFINITY : INFINITY;
    if (ylogx <= -150.0) return signBias ? -0.0f : 0.0f;
  }
  // exp2_inline
  double kd = dadd(ylogx, kExp2ShiftScaled);
  const uint64_t ki = asU64(kd);
  kd = dadd(kd, -kExp2ShiftScaled);
  const double rr = dadd(ylogx, -kd);
  uint64_t t = kExp2Tab[ki % 32u];
  const uint64_t ski = ki + signBias;
  t += ski << (52 - 5);
  const double s = asDouble(t);
  const double zz = dadd(dmul(kExp2C0, rr), kExp2C1);
  const double rr2 = dmul(rr, rr);
  double res = dadd(dmul(kExp2C2, rr), 1.0);
  res = dadd(dmul(zz, rr2), res);
  res = dmul(res, s);
  return float(res);
}

}  // namespace yb
