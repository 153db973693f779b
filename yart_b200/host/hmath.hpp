// hmath.hpp — host-side float math with the reference's exact operation order.
//
// The scene constants the GPU consumes (inverse transforms, padded bounds, light areas and
// powers, env-map CDFs, camera frame) are computed once on the host.  To be bit-identical to
// what the reference computes for the same scene, each helper evaluates the same expression in
// the same order as src/math/{vec,mat,bounds,transform,frame}.hpp.  Compile with
// -ffp-contract=off and without -march (no FMA contraction), like the oracle.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace yartb {

struct f3 {
  float x = 0, y = 0, z = 0;
  f3() = default;
  f3(float a, float b, float c) : x(a), y(b), z(c) {}
  explicit f3(const float* p) : x(p[0]), y(p[1]), z(p[2]) {}
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};

inline f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline f3 operator*(float s, f3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline f3 operator/(f3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
// vec.hpp:394-396 — left to right, no fma
inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// vec.hpp:407-416
inline f3 cross(f3 a, f3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
// vec.hpp:337-343 — accumulates from 0
inline float length2(f3 a) {
  float s = 0.0f;
  s += a.x * a.x;
  s += a.y * a.y;
  s += a.z * a.z;
  return s;
}
inline float length(f3 a) { return std::sqrt(length2(a)); }
inline f3 normalized(f3 a) { return a / length(a); }  // vec.hpp:350-353: divides per component
inline float absDot(f3 a, f3 b) { return std::fabs(dot(a, b)); }
// math_base.hpp:85-92 — NaN in the first argument returns the second
inline float rmin(float m, float n) { return m < n ? m : n; }
inline float rmax(float m, float n) { return m > n ? m : n; }

struct Bounds3 {
  f3 mn{std::numeric_limits<float>::infinity(), std::numeric_limits<float>::infinity(),
        std::numeric_limits<float>::infinity()};
  f3 mx{-std::numeric_limits<float>::infinity(), -std::numeric_limits<float>::infinity(),
        -std::numeric_limits<float>::infinity()};
  // bounds.hpp:38-41 (half area)
  float area() const {
    f3 s = mx - mn;
    return s.x * s.y + s.y * s.z + s.z * s.x;
  }
  // bounds.hpp:43-46
  void expandToInclude(f3 p) {
    mn = {rmin(mn.x, p.x), rmin(mn.y, p.y), rmin(mn.z, p.z)};
    mx = {rmax(mx.x, p.x), rmax(mx.y, p.y), rmax(mx.z, p.z)};
  }
  // bounds.hpp:48-60: union starts from the empty box and folds each argument in
  static Bounds3 join(const Bounds3& a, const Bounds3& b) {
    Bounds3 u;
    for (const Bounds3* p : {&a, &b}) {
      for (int i = 0; i < 3; i++) {
        u.mn[i] = rmin(u.mn[i], p->mn[i]);
        u.mx[i] = rmax(u.mx[i], p->mx[i]);
      }
    }
    return u;
  }
  // bounds.hpp:92-104: tight box then ±float(0.001)
  template <typename It>
  static Bounds3 fromPoints(It first, It last) {
    Bounds3 b;
    for (It p = first; p != last; ++p) {
      for (int i = 0; i < 3; i++) {
        if ((*p)[i] < b.mn[i]) b.mn[i] = (*p)[i];
        if ((*p)[i] > b.mx[i]) b.mx[i] = (*p)[i];
      }
    }
    const float pad = float(0.001);
    b.mn = b.mn - f3(pad, pad, pad);
    b.mx = b.mx + f3(pad, pad, pad);
    return b;
  }
  static Bounds3 fromTriangle(f3 a, f3 b, f3 c) {
    f3 pts[3] = {a, b, c};
    return fromPoints(pts, pts + 3);
  }
};

// Row-major 4x4 / 3x3 (mat.hpp stores row-major: operator()(i,j) = m_data[i*M + j]).
struct Mat4 {
  float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  float operator()(int i, int j) const { return m[i * 4 + j]; }
  float& operator()(int i, int j) { return m[i * 4 + j]; }
};
struct Mat3 {
  float m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  float operator()(int i, int j) const { return m[i * 3 + j]; }
  float& operator()(int i, int j) { return m[i * 3 + j]; }
};

// mat.hpp:262-273: res(i,j) accumulates from 0 over k
inline Mat4 mul(const Mat4& a, const Mat4& b) {
  Mat4 r;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      float s = 0.0f;
      for (int k = 0; k < 4; k++) s += a(i, k) * b(k, j);
      r(i, j) = s;
    }
  return r;
}

// mat.hpp:400-520 + 528-540: cofactor (adjugate) inverse.  Each adjugate entry is a sum of six
// signed triple products evaluated left to right; the table lists them in the reference's order
// (sign, a, b, c) so the rounding sequence is the same.
inline bool inverse(const Mat4& M, Mat4& out) {
  static const int8_t T[16][6][4] = {
    /* 0*/ {{+1, 5, 10, 15}, {-1, 5, 11, 14}, {-1, 9, 6, 15}, {+1, 9, 7, 14}, {+1, 13, 6, 11}, {-1, 13, 7, 10}},
    /* 1*/ {{-1, 1, 10, 15}, {+1, 1, 11, 14}, {+1, 9, 2, 15}, {-1, 9, 3, 14}, {-1, 13, 2, 11}, {+1, 13, 3, 10}},
    /* 2*/ {{+1, 1, 6, 15}, {-1, 1, 7, 14}, {-1, 5, 2, 15}, {+1, 5, 3, 14}, {+1, 13, 2, 7}, {-1, 13, 3, 6}},
    /* 3*/ {{-1, 1, 6, 11}, {+1, 1, 7, 10}, {+1, 5, 2, 11}, {-1, 5, 3, 10}, {-1, 9, 2, 7}, {+1, 9, 3, 6}},
    /* 4*/ {{-1, 4, 10, 15}, {+1, 4, 11, 14}, {+1, 8, 6, 15}, {-1, 8, 7, 14}, {-1, 12, 6, 11}, {+1, 12, 7, 10}},
    /* 5*/ {{+1, 0, 10, 15}, {-1, 0, 11, 14}, {-1, 8, 2, 15}, {+1, 8, 3, 14}, {+1, 12, 2, 11}, {-1, 12, 3, 10}},
    /* 6*/ {{-1, 0, 6, 15}, {+1, 0, 7, 14}, {+1, 4, 2, 15}, {-1, 4, 3, 14}, {-1, 12, 2, 7}, {+1, 12, 3, 6}},
    /* 7*/ {{+1, 0, 6, 11}, {-1, 0, 7, 10}, {-1, 4, 2, 11}, {+1, 4, 3, 10}, {+1, 8, 2, 7}, {-1, 8, 3, 6}},
    /* 8*/ {{+1, 4, 9, 15}, {-1, 4, 11, 13}, {-1, 8, 5, 15}, {+1, 8, 7, 13}, {+1, 12, 5, 11}, {-1, 12, 7, 9}},
    /* 9*/ {{-1, 0, 9, 15}, {+1, 0, 11, 13}, {+1, 8, 1, 15}, {-1, 8, 3, 13}, {-1, 12, 1, 11}, {+1, 12, 3, 9}},
    /*10*/ {{+1, 0, 5, 15}, {-1, 0, 7, 13}, {-1, 4, 1, 15}, {+1, 4, 3, 13}, {+1, 12, 1, 7}, {-1, 12, 3, 5}},
    /*11*/ {{-1, 0, 5, 11}, {+1, 0, 7, 9}, {+1, 4, 1, 11}, {-1, 4, 3, 9}, {-1, 8, 1, 7}, {+1, 8, 3, 5}},
    /*12*/ {{-1, 4, 9, 14}, {+1, 4, 10, 13}, {+1, 8, 5, 14}, {-1, 8, 6, 13}, {-1, 12, 5, 10}, {+1, 12, 6, 9}},
    /*13*/ {{+1, 0, 9, 14}, {-1, 0, 10, 13}, {-1, 8, 1, 14}, {+1, 8, 2, 13}, {+1, 12, 1, 10}, {-1, 12, 2, 9}},
    /*14*/ {{-1, 0, 5, 14}, {+1, 0, 6, 13}, {+1, 4, 1, 14}, {-1, 4, 2, 13}, {-1, 12, 1, 6}, {+1, 12, 2, 5}},
    /*15*/ {{+1, 0, 5, 10}, {-1, 0, 6, 9}, {-1, 4, 1, 10}, {+1, 4, 2, 9}, {+1, 8, 1, 6}, {-1, 8, 2, 5}},
  };
  const float* m = M.m;
  float inv[16];
  for (int e = 0; e < 16; e++) {
    float s = 0.0f;
    for (int t = 0; t < 6; t++) {
      const int8_t* q = T[e][t];
      float a = q[0] < 0 ? -m[q[1]] : m[q[1]];
      float term = a * m[q[2]] * m[q[3]];
      if (t == 0) s = term;
      else s = s + term;  // a - b == a + (-b) bit for bit
    }
    inv[e] = s;
  }
  float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
  if (det == 0.0f) return false;
  det = 1.0f / det;
  for (int e = 0; e < 16; e++) out.m[e] = inv[e] * det;
  return true;
}

// transform.hpp:11-109
struct Transform {
  Mat4 fwd, inv;
  Mat3 nrm, invNrm;  // nrm = transpose(float3x3(inv)), invNrm = transpose(float3x3(fwd))
  Transform() = default;
  explicit Transform(const Mat4& m) : fwd(m) {
    if (!inverse(m, inv)) inv = Mat4();  // the reference would terminate (optional::value in noexcept)
    derive();
  }
  Transform(const Mat4& m, const Mat4& i) : fwd(m), inv(i) { derive(); }
  void derive() {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        nrm(j, i) = inv(i, j);
        invNrm(j, i) = fwd(i, j);
      }
  }
  // mat.hpp:562-574 with float4(v, float(type)): accumulate from 0 over 4 columns
  static f3 apply(const Mat4& M, f3 v, float w) {
    f3 r;
    for (int i = 0; i < 3; i++) {
      float s = 0.0f;
      s += M(i, 0) * v.x;
      s += M(i, 1) * v.y;
      s += M(i, 2) * v.z;
      s += M(i, 3) * w;
      r[i] = s;
    }
    return r;
  }
  static f3 apply3(const Mat3& M, f3 v) {
    f3 r;
    for (int i = 0; i < 3; i++) {
      float s = 0.0f;
      s += M(i, 0) * v.x;
      s += M(i, 1) * v.y;
      s += M(i, 2) * v.z;
      r[i] = s;
    }
    return r;
  }
  f3 point(f3 v) const { return apply(fwd, v, 1.0f); }
  f3 vector(f3 v) const { return apply(fwd, v, 0.0f); }
  f3 normal(f3 v) const { return normalized(apply3(nrm, v)); }
  f3 invPoint(f3 v) const { return apply(inv, v, 1.0f); }
  f3 invVector(f3 v) const { return apply(inv, v, 0.0f); }
  // transform.hpp:75-90: 8 corners as points, then fromPoints (pads ±0.001 again)
  Bounds3 bounds(const Bounds3& b) const {
    f3 c[8] = {{b.mn.x, b.mn.y, b.mn.z}, {b.mn.x, b.mn.y, b.mx.z}, {b.mn.x, b.mx.y, b.mn.z},
               {b.mn.x, b.mx.y, b.mx.z}, {b.mx.x, b.mn.y, b.mn.z}, {b.mx.x, b.mn.y, b.mx.z},
               {b.mx.x, b.mx.y, b.mn.z}, {b.mx.x, b.mx.y, b.mx.z}};
    for (f3& p : c) p = point(p);
    return Bounds3::fromPoints(c, c + 8);
  }
};

// mat.hpp:52-74: rotation about an axis; rows as laid out by the 16-value constructor
inline Mat4 rotation(float angle, f3 axis) {
  const float a = angle;
  const float c = std::cos(a);
  const float s = std::sin(a);
  const f3 n = normalized(axis);
  const f3 t = float(1.0 - c) * n;  // (1.0 - c) is double in the reference, converted to float by T(lhs)
  Mat4 r;
  r.m[0] = c + t.x * n.x;
  r.m[1] = t.y * n.x - s * n.z;
  r.m[2] = t.z * n.x + s * n.y;
  r.m[3] = 0;
  r.m[4] = t.x * n.y + s * n.z;
  r.m[5] = c + t.y * n.y;
  r.m[6] = t.z * n.y - s * n.x;
  r.m[7] = 0;
  r.m[8] = t.x * n.z - s * n.y;
  r.m[9] = t.y * n.z + s * n.x;
  r.m[10] = c + t.z * n.z;
  r.m[11] = 0;
  r.m[12] = r.m[13] = r.m[14] = 0;
  r.m[15] = 1;
  return r;
}

// frame.hpp:21-59
struct Frame {
  f3 x{1, 0, 0}, y{0, 1, 0}, z{0, 0, 1};
  Frame() = default;
  explicit Frame(f3 n) : z(n) {
    const f3 a = std::fabs(n.x) > 0.5 ? f3(0, 1, 0) : f3(1, 0, 0);
    y = normalized(cross(n, a));
    x = cross(n, y);
  }
  Frame(f3 n, f3 t, float hand = 1.0f) : z(n) {
    if (absDot(t, n) > 0.9f) {
      const f3 a = std::fabs(n.x) > 0.5 ? f3(0, 1, 0) : f3(1, 0, 0);
      y = normalized(cross(n, a));
      x = cross(n, y);
    } else {
      y = normalized(cross(n, t)) * hand;
      x = cross(y, z);
    }
  }
  f3 ltw(f3 l) const { return l.x * x + l.y * y + l.z * z; }
};

}  // namespace yartb
