// estimator.cuh — per-pixel estimators (GMoN, median-of-means, mean) and the AgX tonemap.
//
// Restates reference src/core/estimator.hpp:29-46 (MeanEstimator), :53-88 (MoNEstimator),
// :148-198 (GMoNEstimator) and src/core/tonemapping.hpp:14-92 (AgX).  Buckets live in HBM as
// float4 planes [bucket][pixel] = (sum.rgb, count as float bits of a u32 count); samples are added
// in sample order so each bucket sum has the reference's rounding sequence.
#pragma once
#include "../../include/yart_cuda.h"
#include "dmath.cuh"
#include "libm_exact.cuh"

namespace yb {

constexpr int kMaxBuckets = 15;  // integrator.cpp:17: GMoNEstimator(samples, 15)

// estimator.hpp:150-151: m = min(mMax, max(1, 1 + 2 * ((n - 5) / 10)))  (int division, truncating)
#ifdef YB_HOSTSIM
#define YB_HD inline
#else
#define YB_HD __host__ __device__ __forceinline__
#endif
YB_HD int estimatorBuckets(int n, int mMax) {
  int v = 1 + 2 * ((n - 5) / 10);
  if (v < 1) v = 1;
  return v < mMax ? v : mMax;
}

YB_DEV float luma(V3 v) { return dot(v, V3(0.2126f, 0.7152f, 0.0722f)); }  // estimator.hpp:20-23
YB_DEV bool hasnan(V3 v) { return v.x != v.x || v.y != v.y || v.z != v.z; }

// Which samples an estimator accepts (bucket index always advances: estimator.hpp:154-161)
YB_DEV bool estimatorAccepts(int estimator, V3 s) {
  if (hasnan(s)) return false;
  if (estimator == YC_ESTIMATOR_GMON) return s.x >= 0.0f && s.y >= 0.0f && s.z >= 0.0f;
  return true;
}

// std::sort on <= 16 elements is libstdc++'s __insertion_sort (introsort threshold 16); restated
// with the same comparison sequence so NaN lumas (empty buckets) land where the oracle puts them.
YB_DEV void sortByLuma(V3* a, int n) {
  for (int i = 1; i < n; i++) {
    V3 val = a[i];
    float lv = luma(val);
    if (lv < luma(a[0])) {
      for (int j = i; j > 0; j--) a[j] = a[j - 1];
      a[0] = val;
    } else {
      int j = i;
      while (lv < luma(a[j - 1])) {
        a[j] = a[j - 1];
        j--;
      }
      a[j] = val;
    }
  }
}

// getValue() of the three estimators.  acc[i] = bucket sums, cnt[i] = accepted samples, m buckets;
// nSamples = samples handed to the estimator (MeanEstimator divides by it).
YB_DEV V3 estimatorValue(int estimator, V3* acc, const uint32_t* cnt, int m, uint32_t nSamples) {
  if (estimator == YC_ESTIMATOR_MEAN) return acc[0] / float(nSamples);
  if (m == 1) return acc[0] / float(cnt[0]);
  for (int i = 0; i < m; i++) acc[i] /= float(cnt[i]);
  sortByLuma(acc, m);
  if (estimator == YC_ESTIMATOR_MON) return acc[m / 2];
  // Gini function, estimator.hpp:123-130 (GMoNb) = :176-183 (GMoN)
  V3 sum, weightedSum;
  for (int i = 0; i < m; i++) {
    sum += acc[i];
    weightedSum += float(i + 1) * acc[i];  // (i + 1) * vec → vec * float(i + 1)
  }
  float G = (2.0f * luma(weightedSum)) / (float(m) * luma(sum)) - float(m + 1) / float(m);
  if (estimator == YC_ESTIMATOR_GMONB) {
    // estimator.hpp:132-136: below the threshold the mean of the bucket means, else their median (NaN → median)
    if (G <= 0.25f) return sum / float(m);
    return acc[m / 2];
  }
  // GMoN, estimator.hpp:184-191
  if (G > 1.0f) G = 1.0f;
  // c = size_t(G * float(m / 2)) as the x86-64 oracle evaluates it.  G in (-1,0) truncates to 0.
  // G = NaN (black pixel: luma(sum) = 0, or an empty bucket) converts to 2^63; the reference loop
  // `for (i = c; i < m - c; i++) sum += m_acc[i]` then runs i = 2^63 .. 2^63 + m - 1, whose byte
  // offsets 12 * i wrap to 12 * (i - 2^63): it re-reads buckets 0..m-1, and m - 2c wraps to m.
  // That is exactly the c = 0 case, so NaN maps to 0 here.
  float gc = G * float(m / 2);
  int c = (gc != gc || gc < 0.0f) ? 0 : int(gc);
  V3 s2;
  for (int i = c; i < m - c; i++) s2 += acc[i];
  return s2 / float(m - 2 * c);
}

// ---------------------------------------------------------------------------------------
// AgX, tonemapping.hpp:14-92.  Polynomial constants are double literals converted to float per
// term by `double * vec → vec * float(double)` (vec.hpp:276-282); the trailing `- 0.00232` goes
// through vec::operator-(const float&).
// ---------------------------------------------------------------------------------------
struct AgxLook {
  V3 offset, slope, power;
  float sat;
};

YB_DEV AgxLook agxLook(uint32_t tonemap) {
  AgxLook l;
  l.offset = V3(0.0f);
  l.slope = V3(1.0f);
  l.power = V3(1.0f);
  l.sat = 1.0f;
  if (tonemap == YC_TONEMAP_AGX_GOLDEN) {
    l.slope = V3(1.0f, 0.9f, 0.5f);
    l.power = V3(0.8f);
    l.sat = 0.8f;
  } else if (tonemap == YC_TONEMAP_AGX_PUNCHY) {
    l.power = V3(1.35f);
    l.sat = 1.4f;
  }
  return l;
}

YB_DEV V3 agxContrast(V3 x) {
  V3 x2 = x * x;
  V3 x4 = x2 * x2;
  return ((((((x4 * float(15.5)) * x2 - (x4 * float(40.14)) * x) + x4 * float(31.96)) - (x2 * float(6.868)) * x) +
           x2 * float(0.4298)) +
          x * float(0.1191)) -
         float(0.00232);
}

YB_DEV V3 agx(V3 val, const AgxLook& look) {
  const float A[9] = {float(0.842479062253094), float(0.0784335999999992), float(0.0792237451477643),
                      float(0.0423282422610123), float(0.878468636469772), float(0.0791661274605434),
                      float(0.0423756549057051), float(0.0784336), float(0.879142973793104)};
  const float AI[9] = {float(1.19687900512017), float(-0.0980208811401368), float(-0.0990297440797205),
                       float(-0.0528968517574562), float(1.15190312990417), float(-0.0989611768448433),
                       float(-0.0529716355144438), float(-0.0980434501171241), float(1.15107367264116)};
  const float minEv = -12.47393f, maxEv = 4.026069f;
  val = mul3x3(A, val);
  V3 lg(log2fExact(val.x), log2fExact(val.y), log2fExact(val.z));
  // clamp(v, lo, hi) = min(hi, max(lo, v)) with math::min/max semantics (vec.hpp:437-446)
  val = V3(rmin(maxEv, rmax(minEv, lg.x)), rmin(maxEv, rmax(minEv, lg.y)), rmin(maxEv, rmax(minEv, lg.z)));
  val = (val - V3(minEv)) / (V3(maxEv) - V3(minEv));
  val = agxContrast(val);
  // applyLook
  float l = luma(val);
  V3 b = val * look.slope + look.offset;
  val = V3(powfExact(b.x, look.power.x), powfExact(b.y, look.power.y), powfExact(b.z, look.power.z));
  val = V3(l) + look.sat * (val - l);
  // end
  val = mul3x3(AI, val);
  val = V3(rmin(1.0f, rmax(0.0f, val.x)), rmin(1.0f, rmax(0.0f, val.y)), rmin(1.0f, rmax(0.0f, val.z)));
  return V3(powfExact(val.x, 2.2f), powfExact(val.y, 2.2f), powfExact(val.z, 2.2f));
}

}  // namespace yb
