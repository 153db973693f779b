// dmath.cuh — device float math with the reference's operation order (compile with -fmad=false).
//
// Every helper evaluates the same expression, in the same order, as the reference's header-only
// math (src/math/vec.hpp, mat.hpp, transform.hpp, frame.hpp, math.hpp): the oracle is x86-64 g++
// without FMA contraction and its `fma()` is two roundings (vec.hpp:325-334), so the product must
// not fuse either.  IEEE division / sqrt are nvcc defaults (-prec-div/-prec-sqrt).
#pragma once
#ifdef YB_HOSTSIM
#include "hostshim.hpp"
#define YB_DEV inline
#define YB_DEV_NI inline
#define YB_CONST static const
#define YB_TABLE static const
#else
#include <cuda_runtime.h>
#define YB_DEV __device__ __forceinline__
#define YB_DEV_NI __device__ __noinline__
#define YB_CONST __device__ __constant__
#define YB_TABLE __device__ const  // global memory: for tables indexed differently per lane
#endif
#include <math.h>
#include <stdint.h>

namespace yb {

struct V3 {
  float x, y, z;
  YB_DEV V3() : x(0.f), y(0.f), z(0.f) {}
  YB_DEV V3(float a, float b, float c) : x(a), y(b), z(c) {}
  YB_DEV explicit V3(float s) : x(s), y(s), z(s) {}
  YB_DEV explicit V3(const float* p) : x(p[0]), y(p[1]), z(p[2]) {}
  YB_DEV float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct V2 {
  float x, y;
  YB_DEV V2() : x(0.f), y(0.f) {}
  YB_DEV V2(float a, float b) : x(a), y(b) {}
};

YB_DEV V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
YB_DEV V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
YB_DEV V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
YB_DEV V3 operator/(V3 a, V3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
YB_DEV V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
YB_DEV V3 operator*(float s, V3 a) { return V3(a.x * s, a.y * s, a.z * s); }  // vec.hpp:276-282: rhs * T(lhs)
YB_DEV V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
YB_DEV V3 operator+(V3 a, float s) { return V3(a.x + s, a.y + s, a.z + s); }
YB_DEV V3 operator-(V3 a, float s) { return V3(a.x - s, a.y - s, a.z - s); }
YB_DEV V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
YB_DEV V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
YB_DEV V3& operator*=(V3& a, V3 b) { a = a * b; return a; }
YB_DEV V3& operator*=(V3& a, float s) { a = a * s; return a; }
YB_DEV V3& operator/=(V3& a, float s) { a = a / s; return a; }

YB_DEV V2 operator+(V2 a, V2 b) { return V2(a.x + b.x, a.y + b.y); }
YB_DEV V2 operator*(V2 a, float s) { return V2(a.x * s, a.y * s); }
YB_DEV V2 operator*(float s, V2 a) { return V2(a.x * s, a.y * s); }

// vec.hpp:394-396
YB_DEV float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// vec.hpp:399-404: copysign(dot, 1.0f)
YB_DEV float absDot(V3 a, V3 b) { return fabsf(dot(a, b)); }
// vec.hpp:407-416
YB_DEV V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// vec.hpp:337-343: sum starts at 0 (0 + x*x is exact)
YB_DEV float length2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
YB_DEV float length2(V2 a) { return a.x * a.x + a.y * a.y; }
YB_DEV float length(V3 a) { return sqrtf(length2(a)); }
YB_DEV V3 normalized(V3 a) { return a / length(a); }  // vec.hpp:350-353: per-component divide
YB_DEV float sum3(V3 a) { return a.x + a.y + a.z; }   // vec.hpp:316-321 (0 + x + y + z)
// math_base.hpp:85-92: `m < n ? m : n` — a NaN first argument yields the second
YB_DEV float rmin(float m, float n) { return m < n ? m : n; }
YB_DEV float rmax(float m, float n) { return m > n ? m : n; }
// std::clamp(v, lo, hi)
YB_DEV float sclamp(float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); }
// vec.hpp:378-383: running max starts at FLT_MIN
YB_DEV float maxComponent(V3 v) {
  float m = 1.17549435e-38f;
  if (v.x > m) m = v.x;
  if (v.y > m) m = v.y;
  if (v.z > m) m = v.z;
  return m;
}
// math_base.hpp:34-37
YB_DEV float lerpf(float a, float b, float t) { return (1.0f - t) * a + t * b; }

// mat.hpp:562-574 on float4(v, w): accumulate from 0 over the four columns of rows 0..2
YB_DEV V3 xformRows(const float* __restrict__ m, V3 v, float w) {
  V3 r;
  r.x = ((0.0f + m[0] * v.x) + m[1] * v.y + m[2] * v.z) + m[3] * w;
  r.y = ((0.0f + m[4] * v.x) + m[5] * v.y + m[6] * v.z) + m[7] * w;
  r.z = ((0.0f + m[8] * v.x) + m[9] * v.y + m[10] * v.z) + m[11] * w;
  return r;
}
YB_DEV V3 mul3x3(const float* __restrict__ m, V3 v) {
  V3 r;
  r.x = (0.0f + m[0] * v.x) + m[1] * v.y + m[2] * v.z;
  r.y = (0.0f + m[3] * v.x) + m[4] * v.y + m[5] * v.z;
  r.z = (0.0f + m[6] * v.x) + m[7] * v.y + m[8] * v.z;
  return r;
}

// frame.hpp:21-59
struct Frame {
  V3 x, y, z;
  YB_DEV Frame() : x(1, 0, 0), y(0, 1, 0), z(0, 0, 1) {}
  YB_DEV explicit Frame(V3 n) : z(n) { fromNormal(n); }
  YB_DEV Frame(V3 n, V3 t, float hand = 1.0f) : z(n) {
    if (absDot(t, n) > 0.9f) {
      fromNormal(n);
    } else {
      y = normalized(cross(n, t)) * hand;
      x = cross(y, z);
    }
  }
  YB_DEV void fromNormal(V3 n) {
    const V3 a = fabsf(n.x) > 0.5f ? V3(0, 1, 0) : V3(1, 0, 0);
    y = normalized(cross(n, a));
    x = cross(n, y);
  }
  YB_DEV V3 wtl(V3 w) const { return V3(dot(w, x), dot(w, y), dot(w, z)); }
  YB_DEV V3 ltw(V3 l) const { return l.x * x + l.y * y + l.z * z; }
};

// math.hpp:15-20: -wo + normal * 2.0 * dot(wo, normal)
YB_DEV V3 reflect(V3 wo, V3 n) { return -wo + (n * 2.0f) * dot(wo, n); }

// math.hpp:22-41
YB_DEV bool refract(V3 wi, V3 n, float ior, V3& wt) {
  float cosTheta = dot(wi, n);
  if (cosTheta < 0.0f) {
    ior = 1.0f / ior;
    cosTheta *= -1.0f;
    n = n * -1.0f;
  }
  float sin2Theta = (1.0f - cosTheta * cosTheta);
  float sin2Theta_t = sin2Theta / (ior * ior);
  if (sin2Theta_t >= 1.0f) return false;
  float cosTheta_t = sqrtf(1.0f - sin2Theta_t);
  wt = -wi / ior + (cosTheta / ior - cosTheta_t) * n;
  return true;
}

// math.hpp:43-61
YB_DEV float fresnelDielectric(float cosTheta, float ior) {
  cosTheta = sclamp(cosTheta, -1.0f, 1.0f);
  if (cosTheta < 0.0f) {
    ior = 1.0f / ior;
    cosTheta = -cosTheta;
  }
  float sin2Theta = (1.0f - cosTheta * cosTheta);
  float sin2Theta_t = sin2Theta / (ior * ior);
  if (sin2Theta_t >= 1.0f) return 1.0f;
  float cosTheta_t = sqrtf(1.0f - sin2Theta_t);
  float r_prl = (ior * cosTheta - cosTheta_t) / (ior * cosTheta + cosTheta_t);
  float r_per = (cosTheta - ior * cosTheta_t) / (cosTheta + ior * cosTheta_t);
  return (r_prl * r_prl + r_per * r_per) * 0.5f;
}

// math.hpp:81-88
YB_DEV V3 fresnelSchlick(V3 r, float cosTheta) {
  const float k = 1.0f - cosTheta;
  const float k2 = k * k;
  return r + (V3(1.0f) - r) * (k2 * k2 * k);
}

// math.hpp:151-166
YB_DEV V2 octahedralUV(V3 v) {
  V2 res;
  V3 vAbs(fabsf(v.x), fabsf(v.y), fabsf(v.z));
  v /= sum3(vAbs);
  vAbs /= sum3(vAbs);
  if (v.y >= 0) {
    res = V2(v.x, v.z);
  } else {
    res = V2((1.0f - vAbs.z) * copysignf(1.0f, v.x), (1.0f - vAbs.x) * copysignf(1.0f, v.z));
  }
  return V2((res.x + 1.0f) * 0.5f, (res.y + 1.0f) * 0.5f);
}

// math.hpp:168-179
YB_DEV V3 invOctahedralUV(V2 uv) {
  V3 res;
  res.x = 2.0f * uv.x - 1.0f;
  res.z = 2.0f * uv.y - 1.0f;
  res.y = 1.0f - (fabsf(res.x) + fabsf(res.z));
  if (res.y < 0.0f) {
    float xo = res.x;
    res.x = (1.0f - fabsf(res.z)) * copysignf(1.0f, res.x);
    res.z = (1.0f - fabsf(xo)) * copysignf(1.0f, res.z);
  }
  return normalized(res);
}

}  // namespace yb
