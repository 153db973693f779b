// bsdf.cuh — device ParametricBSDF: LUT lookups, GGX, the four lobes, eval / pdf / sample.
//
// Restates reference src/bsdf/luts.hpp:33-187 (bi/tri-linear LUT lookups incl. the x86-64
// behaviour of size_t(negative float)), src/core/bsdf.hpp:16-18 (roughen), :175-291 (GGX),
// src/core/bsdf.cpp:5-41 (frame wrappers), src/bsdf/parametric.cpp:84-258 (fImpl / pdfImpl /
// sampleImpl) and :260-838 (lobes, attenuation), and src/math/sampling.hpp:30-45.
// Quirks kept on purpose: sampleImpl returns the chosen lobe's f/pdf alone;
// pdfImpl does not apply the anisotropy rotation; pClearcoat is computed in double; emission is
// only returned by the diffuse branch of sampleGlossy; clearcoat Fresnel uses IOR 1.5 except in
// the smooth-coat branch.
#pragma once
#include "libm_exact.cuh"
#include "scene_dev.cuh"
#include "texture.cuh"

namespace yb {

constexpr float kPi = 3.14159274101257324f;  // float(M_PI)

// ---------------------------------------------------------------------------------------
// LUTs.  Table block layout (floats): ggxE[32*32] ggxEavg[32] baseE[16^3] baseEavg[16^2]
//                                     glassE[16^3] glassEavg[16^2] glassInvE[16^3] glassInvEavg[16^2]
// ---------------------------------------------------------------------------------------
constexpr int kLutE = 0, kLutEavg = 1024, kLutBaseE = 1056, kLutBaseEavg = 5152, kLutGlassE = 5408,
              kLutGlassEavg = 9504, kLutGlassInvE = 9760, kLutGlassInvEavg = 13856;

// min(size_t(x), n) as the x86-64 g++ oracle evaluates it (luts.hpp:35 etc.): the conversion
// truncates toward zero into int64 and reinterprets, so x in (-1,0) → 0, x <= -1 → huge → n,
// NaN → 0x8000000000000000 → n.
YB_DEV int lutIndex(float x, int n) {
  if (x != x) return n;
  if (x <= -1.0f) return n;
  if (x < 0.0f) return 0;
  return x >= float(n) ? n : int(x);
}

YB_DEV float trilerp8(const float* x, float u, float v, float w) {
  float up = 1.0f - u, vp = 1.0f - v, wp = 1.0f - w;
  return x[0] * up * vp * wp + x[1] * up * vp * w + x[2] * up * v * wp + x[3] * up * v * w + x[4] * u * vp * wp +
         x[5] * u * vp * w + x[6] * u * v * wp + x[7] * u * v * w;
}

// luts.hpp:33-44
YB_DEV float ggxE(const float* lut, float cosTheta, float r) {
  float ro = r * 31.0f, co = cosTheta * 31.0f;
  int ri = lutIndex(ro, 30), ci = lutIndex(co, 30);
  ro -= float(ri);
  co -= float(ci);
  const float* t = lut + kLutE;
  float d00 = t[ri * 32 + ci], d01 = t[ri * 32 + ci + 1], d10 = t[(ri + 1) * 32 + ci], d11 = t[(ri + 1) * 32 + ci + 1];
  return bilerp1(d00, d01, d10, d11, ro, co);
}
// luts.hpp:52-57
YB_DEV float ggxEavg(const float* lut, float r) {
  int ri = lutIndex(r * 31.0f, 30);
  float ro = r * 31.0f - float(ri);
  return lerpf(lut[kLutEavg + ri], lut[kLutEavg + ri + 1], ro);
}
// luts.hpp:69-98
YB_DEV float ggxBaseE(const float* lut, float f0, float r, float cosTheta) {
  float f0o = f0 * 15.0f, ro = r * 15.0f, co = cosTheta * 15.0f;
  int f0i = lutIndex(f0o, 14), ri = lutIndex(ro, 14), ci = lutIndex(co, 14);
  f0o -= float(f0i);
  ro -= float(ri);
  co -= float(ci);
  const float* t = lut + kLutBaseE;
  float vals[8] = {t[(f0i * 16 + ri) * 16 + ci],           t[(f0i * 16 + ri) * 16 + ci + 1],
                   t[(f0i * 16 + ri + 1) * 16 + ci],       t[(f0i * 16 + ri + 1) * 16 + ci + 1],
                   t[((f0i + 1) * 16 + ri) * 16 + ci],     t[((f0i + 1) * 16 + ri) * 16 + ci + 1],
                   t[((f0i + 1) * 16 + ri + 1) * 16 + ci], t[((f0i + 1) * 16 + ri + 1) * 16 + ci + 1]};
  return trilerp8(vals, f0o, ro, co);
}
// luts.hpp:107-117
YB_DEV float ggxBaseEavg(const float* lut, float f0, float r) {
  int f0i = lutIndex(f0 * 15.0f, 14), ri = lutIndex(r * 15.0f, 14);
  float f0o = f0 * 15.0f - float(f0i), ro = r * 15.0f - float(ri);
  const float* t = lut + kLutBaseEavg;
  return bilerp1(t[f0i * 16 + ri], t[f0i * 16 + ri + 1], t[(f0i + 1) * 16 + ri], t[(f0i + 1) * 16 + ri + 1], f0o, ro);
}
// luts.hpp:127-159 (note the [f0][cos][r] index order and trilerp(vals, f0o, co, ro))
YB_DEV float ggxGlassE(const float* lut, float ior, float r, float cosTheta) {
  bool inv = ior < 1.0f;
  if (inv) ior = 1.0f / ior;
  float f0 = sqrtf(fabsf((1.0f - ior) / (1.0f + ior)));
  int f0i = lutIndex(f0 * 15.0f, 14), ri = lutIndex(r * 15.0f, 14), ci = lutIndex(cosTheta * 15.0f, 14);
  float f0o = f0 * 15.0f - float(f0i), ro = r * 15.0f - float(ri), co = cosTheta * 15.0f - float(ci);
  const float* t = lut + (inv ? kLutGlassInvE : kLutGlassE);
  float vals[8] = {t[(f0i * 16 + ci) * 16 + ri],           t[(f0i * 16 + ci) * 16 + ri + 1],
                   t[(f0i * 16 + ci + 1) * 16 + ri],       t[(f0i * 16 + ci + 1) * 16 + ri + 1],
                   t[((f0i + 1) * 16 + ci) * 16 + ri],     t[((f0i + 1) * 16 + ci) * 16 + ri + 1],
                   t[((f0i + 1) * 16 + ci + 1) * 16 + ri], t[((f0i + 1) * 16 + ci + 1) * 16 + ri + 1]};
  return trilerp8(vals, f0o, co, ro);
}
// luts.hpp:168-187
YB_DEV float ggxGlassEavg(const float* lut, float ior, float r) {
  bool inv = ior < 1.0f;
  if (inv) ior = 1.0f / ior;
  float f0 = sqrtf(fabsf((1.0f - ior) / (1.0f + ior)));
  int f0i = lutIndex(f0 * 15.0f, 14), ri = lutIndex(r * 15.0f, 14);
  float f0o = f0 * 15.0f - float(f0i), ro = r * 15.0f - float(ri);
  const float* t = lut + (inv ? kLutGlassInvEavg : kLutGlassEavg);
  return bilerp1(t[f0i * 16 + ri], t[f0i * 16 + ri + 1], t[(f0i + 1) * 16 + ri], t[(f0i + 1) * 16 + ri + 1], f0o, ro);
}

// core/bsdf.hpp:16-18
YB_DEV float roughen(float r) { return rmax(r, sclamp(r * 2.0f, 0.1f, 0.3f)); }
// parametric.cpp:7-9
YB_DEV float FavgFit(float ior) { return (ior - 1.0f) / (4.08567f + 1.00071f * ior); }

// sampling.hpp:30-38
YB_DEV V3 sampleCosineHemisphere(V2 u) {
  const float phi = u.x * 2.0f * kPi;
  const float sqrtr2 = sqrtf(u.y);
  const float x = cosfExact(phi) * sqrtr2;
  const float y = sinfExact(phi) * sqrtr2;
  const float z = sqrtf(1.0f - u.y);
  return V3(x, y, z);
}
// sampling.hpp:40-45
YB_DEV V2 sampleDiskUniform(V2 u) {
  const float r = sqrtf(u.x);
  const float theta = 2.0f * kPi * u.y;
  return V2(r * cosfExact(theta), r * sinfExact(theta));
}
// sampling.hpp:54-64
YB_DEV V3 sampleTriUniform(V2 u) {
  float b0, b1;
  if (u.x < u.y) {
    b0 = u.x * 0.5f;
    b1 = u.y - b0;
  } else {
    b1 = u.y * 0.5f;
    b0 = u.x - b1;
  }
  return V3(b0, b1, 1.0f - b0 - b1);
}

// ---------------------------------------------------------------------------------------
// GGX, core/bsdf.hpp:175-291
// ---------------------------------------------------------------------------------------
struct GGX {
  float ax, ay, rough;
  YB_DEV explicit GGX(float roughness) : rough(roughness) { ax = ay = roughness * roughness; }
  YB_DEV GGX(float roughness, float anisotropic) : rough(roughness) {
    float alpha = roughness * roughness;
    float aspect = sqrtf(1.0f - 0.9f * anisotropic);
    ax = alpha / aspect;
    ay = alpha * aspect;
  }
  YB_DEV float r() const { return rough; }
  YB_DEV bool smooth() const { return ax < 1e-3f && ay < 1e-3f; }
  YB_DEV float mdf(V3 w) const {
    const float cos2Theta = w.z * w.z;
    const float sin2Theta = fmaxf(0.0f, 1.0f - cos2Theta);
    const float tan2Theta = sin2Theta / cos2Theta;
    const float cos4Theta = cos2Theta * cos2Theta;
    float k = tan2Theta;
    if (ax != ay) {
      float cos2Phi = sin2Theta == 0.0f ? 1.0f : w.x * w.x / sin2Theta;
      float sin2Phi = sin2Theta == 0.0f ? 1.0f : w.y * w.y / sin2Theta;
      k *= (cos2Phi / (ax * ax) + sin2Phi / (ay * ay));
    } else {
      k /= (ax * ax);
    }
    const float k2 = (1.0f + k) * (1.0f + k);
    return 1.0f / (kPi * ax * ay * cos4Theta * k2);
  }
  YB_DEV float lambda(V3 w) const {
    const float cos2Theta = w.z * w.z;
    const float sin2Theta = (1.0f - cos2Theta);
    const float tan2Theta = sin2Theta / cos2Theta;
    float alpha2 = ax * ax;
    if (ax != ay) {
      float cos2Phi = sin2Theta == 0.0f ? 1.0f : w.x * w.x / sin2Theta;
      float sin2Phi = sin2Theta == 0.0f ? 0.0f : w.y * w.y / sin2Theta;
      alpha2 = alpha2 * cos2Phi + ay * ay * sin2Phi;
    }
    return (sqrtf(1.0f + alpha2 * tan2Theta) - 1.0f) * 0.5f;
  }
  YB_DEV float g1(V3 w) const { return 1.0f / (1.0f + lambda(w)); }
  YB_DEV float g(V3 wo, V3 wi) const { return 1.0f / (1.0f + lambda(wo) + lambda(wi)); }
  YB_DEV float vmdf(V3 w, V3 wm) const { return g1(w) / fabsf(w.z) * mdf(wm) * absDot(w, wm); }
  YB_DEV V3 sampleVisibleMicrofacet(V3 w, V2 u) const {
    V3 wh = normalized(V3(ax * w.x, ay * w.y, w.z));
    if (wh.z < 0) wh = wh * -1.0f;
    const V3 b = (wh.z < 0.9999f) ? normalized(cross(V3(0.0f, 0.0f, 1.0f), wh)) : V3(1.0f, 0.0f, 0.0f);
    const V3 t = cross(wh, b);
    V2 p = sampleDiskUniform(u);
    const float h = sqrtf(1.0f - p.x * p.x);
    p.y = lerpf(h, p.y, 0.5f * wh.z + 0.5f);
    const float pz = sqrtf(fmaxf(0.0f, 1.0f - length2(p)));
    V3 nh = p.x * b + p.y * t + pz * wh;
    return normalized(V3(ax * nh.x, ay * nh.y, fmaxf(1e-6f, nh.z)));
  }
};

// ---------------------------------------------------------------------------------------
// BSDFSample, core/bsdf.hpp:20-41
// ---------------------------------------------------------------------------------------
enum Scatter { Absorbed = 0, Emitted = 1, Reflected = 2, Transmitted = 4, Diffuse = 8, Glossy = 16, Specular = 32 };

struct BSDFSample {
  int scatter;
  V3 f, Le, wi;
  float pdf, roughness;
  YB_DEV BSDFSample() : scatter(0), pdf(0.f), roughness(0.f) {}
  YB_DEV BSDFSample(int s, V3 f_, V3 Le_, V3 wi_, float pdf_, float r_)
    : scatter(s), f(f_), Le(Le_), wi(wi_), pdf(pdf_), roughness(r_) {}
  YB_DEV bool is(int flag) const { return (scatter & flag) != 0; }
};

// Per-hit evaluated material inputs (texture fetches of fImpl/pdfImpl/sampleImpl preambles).
struct MatEval {
  V3 base;
  float r, m, t, c, cr;
};

YB_DEV MatEval evalMaterialTextures(const DScene& sc, const YcMaterial& mat, V2 uv) {
  MatEval e;
  e.base = V3(mat.base);
  if (mat.baseTex >= 0) e.base *= sampleU8RGB(sc, mat.baseTex, uv);
  e.r = mat.roughness, e.m = mat.metallic, e.t = mat.transmission;
  e.c = mat.clearcoat, e.cr = mat.clearcoatRoughness;
  if (mat.mrTex >= 0) {
    V2 mr = sampleU8RG(sc, mat.mrTex, uv);
    e.r *= mr.x;
    e.m *= mr.y;
  }
  if (mat.transTex >= 0) e.t *= sampleU8Mono(sc, mat.transTex, uv);
  if (mat.ccTex >= 0) {
    float cc = sampleU8Mono(sc, mat.ccTex, uv);  // float2(mono): same value on both (parametric.cpp:101-105)
    e.c *= cc;
    e.cr *= cc;
  }
  return e;
}

struct Bsdf {
  const DScene& sc;
  const YcMaterial& mat;
  YB_DEV Bsdf(const DScene& s, const YcMaterial& m) : sc(s), mat(m) {}

  // ---- metallic, parametric.cpp:260-352 ---------------------------------------------------
  YB_DEV V3 fMetallic(V3 wo, V3 wi, V3 base, const GGX& mf) const {
    if (mf.smooth()) return V3();
    const float cosTheta_o = fabsf(wo.z), cosTheta_i = fabsf(wi.z);
    if (cosTheta_i == 0 || cosTheta_o == 0) return V3();
    V3 wm = wo + wi;
    if (length2(wm) == 0.0f) return V3();
    wm = normalized(wm.z < 0.0f ? -wm : wm);
    const V3 Fss = fresnelSchlick(base, absDot(wo, wm));
    const V3 Mss = Fss * mf.mdf(wm) * mf.g(wo, wi) / (4 * cosTheta_o * cosTheta_i);
    const float Ess = ggxE(sc.lut, cosTheta_o, mf.r());
    const V3 Mms = Mss * base * (1.0f - Ess) / Ess;
    return Mss + Mms;
  }
  YB_DEV float pdfMetallic(V3 wo, V3 wi, const GGX& mf) const {
    if (mf.smooth()) return 0;
    V3 wm = wo + wi;
    if (length2(wm) == 0.0f) return 0;
    wm = normalized(wm.z < 0.0f ? -wm : wm);
    return mf.vmdf(wo, wm) / (4 * absDot(wo, wm));
  }
  YB_DEV BSDFSample sampleMetallic(V3 wo, V3 base, const GGX& mf, V2 u, float uc) const {
    if (mf.smooth()) {
      const V3 F = fresnelSchlick(base, wo.z);
      return BSDFSample(Reflected | Specular, F / fabsf(wo.z), V3(), V3(-wo.x, -wo.y, wo.z), 1.0f, 0.0f);
    }
    V3 wm = mf.sampleVisibleMicrofacet(wo, u);
    V3 wi = reflect(wo, wm);
    if (wo.z * wi.z < 0.0f) return BSDFSample();
    const float pdf = mf.vmdf(wo, wm) / (4 * absDot(wo, wm));
    const float cosTheta_o = fabsf(wo.z), cosTheta_i = fabsf(wi.z);
    const V3 Fss = fresnelSchlick(base, absDot(wo, wm));
    const V3 Mss = Fss * mf.mdf(wm) * mf.g(wo, wi) / (4 * cosTheta_o * cosTheta_i);
    const float Ess = ggxE(sc.lut, cosTheta_o, mf.r());
    const V3 Mms = Mss * base * (1.0f - Ess) / Ess;
    return BSDFSample(Reflected | Glossy, Mss + Mms, V3(), wi, pdf, mat.roughness);
  }

  // ---- dielectric, parametric.cpp:354-575 ---------------------------------------------------
  YB_DEV V3 fDielectric(V3 wo, V3 wi, V3 base, const GGX& mf) const {
    if (mf.smooth()) return V3();
    const float cosTheta_o = wo.z, cosTheta_i = wi.z;
    const bool isReflection = cosTheta_o * cosTheta_i > 0.0f;
    float ior = 1.0f;
    if (!isReflection) ior = cosTheta_o > 0.0f ? mat.ior : 1.0f / mat.ior;
    V3 wm = ior * wi + wo;
    if (cosTheta_i == 0.0f || cosTheta_o == 0.0f || length2(wm) == 0.0f) return V3();
    wm = normalized(wm.z < 0.0f ? -wm : wm);
    if (dot(wm, wi) * cosTheta_i < 0.0f || dot(wm, wo) * cosTheta_o < 0.0f) return V3();
    const float Fss = fresnelDielectric(absDot(wo, wm), ior);
    const float T = 1.0f - Fss;
    const float E_o = ggxGlassE(sc.lut, ior, mf.r(), fabsf(cosTheta_o));
    if (isReflection) {
      const float Mss = mf.mdf(wm) * mf.g(wo, wi) / (4 * cosTheta_o * cosTheta_i);
      return V3(Fss * Mss / E_o);
    } else if (mat.thinTransmission) {
      V3 wip = reflect(-wi, V3(0.0f, 0.0f, 1.0f));
      wm = normalized(wip + wo);
      const float cosTheta_ip = fabsf(wip.z);
      const float Tss = mf.mdf(wm) * mf.g(wo, wip) / (4 * cosTheta_o * cosTheta_ip);
      return T * base * Tss / E_o;
    } else {
      const float temp = dot(wi, wm) * ior + dot(wo, wm);
      const float dwm_dwi = absDot(wi, wm) * absDot(wo, wm) / (temp * temp);
      const float Tss = mf.mdf(wm) * mf.g(wo, wi) * dwm_dwi / (fabsf(cosTheta_i * cosTheta_o));
      return T * base * Tss / E_o;
    }
  }
  YB_DEV float pdfDielectric(V3 wo, V3 wi, const GGX& mf) const {
    if (mf.smooth()) return 0;
    const float cosTheta_o = wo.z, cosTheta_i = wi.z;
    const bool isReflection = cosTheta_o * cosTheta_i > 0.0f;
    float ior = 1.0f;
    if (!isReflection) ior = cosTheta_o > 0.0f ? mat.ior : 1.0f / mat.ior;
    V3 wm = ior * wi + wo;
    if (cosTheta_i == 0.0f || cosTheta_o == 0.0f || length2(wm) == 0.0f) return 0;
    wm = normalized(wm.z < 0.0f ? -wm : wm);
    if (dot(wm, wi) * cosTheta_i < 0.0f || dot(wm, wo) * cosTheta_o < 0.0f) return 0;
    const float F = fresnelDielectric(dot(wo, wm), mat.ior);
    const float T = 1.0f - F;
    float pdf;
    if (isReflection) {
      pdf = mf.vmdf(wo, wm) / (4 * absDot(wo, wm)) * F;
    } else if (mat.thinTransmission) {
      V3 wip = reflect(-wi, V3(0.0f, 0.0f, 1.0f));
      wm = normalized(wip + wo);
      pdf = mf.vmdf(wo, wm) / (4 * absDot(wo, wm)) * T;
    } else {
      const float temp = dot(wi, wm) + dot(wo, wm) / ior;
      const float dwm_dwi = absDot(wo, wm) / (temp * temp);
      pdf = mf.vmdf(wo, wm) * dwm_dwi * T;
    }
    return pdf;
  }
  YB_DEV BSDFSample sampleDielectric(V3 wo, V3 base, const GGX& mf, V2 u, float uc) const {
    const float ior = (mat.thinTransmission || wo.z > 0.0f) ? mat.ior : 1.0f / mat.ior;
    if (mf.smooth()) {
      float F = fresnelDielectric(fabsf(wo.z), ior);
      float T = 1.0f - F;
      if (uc < F) {
        V3 wi(-wo.x, -wo.y, wo.z);
        return BSDFSample(Reflected | Specular, V3(F / fabsf(wi.z)), V3(), wi, F, 0.0f);
      } else {
        V3 wi;
        if (mat.thinTransmission) wi = -wo;
        else if (!refract(wo, V3(0.0f, 0.0f, 1.0f), mat.ior, wi)) return BSDFSample();
        return BSDFSample(Transmitted | Specular, T * base / fabsf(wi.z), V3(), wi, T, 0.0f);
      }
    }
    V3 wm = mf.sampleVisibleMicrofacet(wo, u);
    const float Fss = fresnelDielectric(absDot(wo, wm), ior);
    const float cosTheta_o = fabsf(wo.z);
    const float E_o = ggxGlassE(sc.lut, ior, mf.r(), cosTheta_o);
    if (uc < Fss) {
      const V3 wi = reflect(wo, wm);
      if (wo.z * wi.z < 0.0f) return BSDFSample();
      const float cosTheta_i = fabsf(wi.z);
      const float Mss = mf.mdf(wm) * mf.g(wo, wi) / (4 * cosTheta_o * cosTheta_i);
      const float pdf = mf.vmdf(wo, wm) / (4 * absDot(wo, wm)) * Fss;
      return BSDFSample(Reflected | Glossy, V3(Fss * Mss / E_o), V3(), wi, pdf, mf.r());
    } else if (mat.thinTransmission) {
      const V3 wi = reflect(wo, wm) * V3(1.0f, 1.0f, -1.0f);
      const float cosTheta_i = fabsf(wi.z);
      const float Tss = mf.mdf(wm) * mf.g(wo, wi) / (4 * cosTheta_o * cosTheta_i);
      const float pdf = mf.vmdf(wo, wm) / (4 * absDot(wo, wm)) * (1.0f - Fss);
      return BSDFSample(Transmitted | Glossy, (1.0f - Fss) * Tss * base / E_o, V3(), wi, pdf, mf.r());
    } else {
      V3 wi;
      const bool tir = !refract(wo, wm, mat.ior, wi);
      if (tir || wo.z * wi.z > 0.0f || wi.z == 0.0f) return BSDFSample();
      const float temp = dot(wi, wm) * ior + dot(wo, wm);
      const float dwm_dwi = absDot(wi, wm) / (temp * temp);
      const float pdf = mf.vmdf(wo, wm) * dwm_dwi * (1.0f - Fss);
      const float Tss = mf.mdf(wm) * mf.g(wo, wi) * fabsf(dot(wi, wm) * dot(wo, wm) / (wi.z * wo.z * temp * temp));
      return BSDFSample(Transmitted | Glossy, (1.0f - Fss) * Tss * base / E_o, V3(), wi, pdf, mf.r());
    }
  }

  // ---- glossy (diffuse + dielectric spec), parametric.cpp:577-730 ---------------------------
  YB_DEV V3 fGlossy(V3 wo, V3 wi, V3 base, const GGX& mf) const {
    if (mf.smooth()) return V3();
    const float cosTheta_o = fabsf(wo.z), cosTheta_i = fabsf(wi.z);
    if (cosTheta_i == 0 || cosTheta_o == 0) return V3();
    V3 wm = wo + wi;
    if (length2(wm) == 0.0f) return V3();
    wm = normalized(wm.z < 0.0f ? -wm : wm);
    const float Fss = fresnelDielectric(dot(wo, wm), mat.ior);
    const float Mss = mf.mdf(wm) * mf.g(wo, wi) / (4 * cosTheta_o * cosTheta_i);
    const float Favg = FavgFit(mat.ior);
    const float Eavg = ggxEavg(sc.lut, mf.r());
    const float Mms = (1.0f - ggxE(sc.lut, cosTheta_o, mf.r())) * (1.0f - ggxE(sc.lut, cosTheta_i, mf.r())) /
                      (kPi * (1.0f - Eavg));
    const float Fms = Favg * Favg * Eavg / (1.0f - Favg * (1.0f - Eavg));
    const float r = (1.0f - mat.ior) / (1.0f + mat.ior);
    const float F0 = r * r;
    const float cDiffuse = (1.0f - ggxBaseE(sc.lut, F0, mf.r(), cosTheta_o)) *
                           (1.0f - ggxBaseE(sc.lut, F0, mf.r(), cosTheta_i)) /
                           (kPi * (1.0f - ggxBaseEavg(sc.lut, F0, mf.r())));
    const V3 diffuse = base * cDiffuse;
    return V3(Fss * Mss + Mms * Fms) + diffuse;
  }
  YB_DEV float pdfGlossy(V3 wo, V3 wi, const GGX& mf) const {
    if (mf.smooth()) return 0;
    const float cosTheta_o = fabsf(wo.z), cosTheta_i = fabsf(wi.z);
    V3 wm = wo + wi;
    if (length2(wm) == 0.0f) return 0;
    wm = normalized(wm.z < 0.0f ? -wm : wm);
    const float Fss = fresnelDielectric(dot(wo, wm), mat.ior);
    const float Favg = FavgFit(mat.ior);
    const float EmsAvg = ggxEavg(sc.lut, mf.r());
    const float Fms = Favg * Favg * EmsAvg / (1.0f - Favg * (1.0f - EmsAvg));
    const float Ems_o = ggxE(sc.lut, cosTheta_o, mf.r());
    const float kappa = 1.0f - (Favg * Ems_o + Fms * (1.0f - Ems_o));
    return (Fss + Fms) * mf.vmdf(wo, wm) / (4 * absDot(wo, wm)) + cosTheta_i * kappa;
  }
  YB_DEV BSDFSample sampleGlossy(V3 wo, V3 base, V3 emission, const GGX& mf, V2 u, float uc) const {
    const float cosTheta_o = wo.z;  // raw (may be negative): LUT UB emulated by lutIndex
    const float Favg = FavgFit(mat.ior);
    const float Eavg = ggxEavg(sc.lut, mf.r());
    const float Fms = Favg * Favg * Eavg / (1.0f - Favg * (1.0f - Eavg));
    const float E_o = ggxE(sc.lut, cosTheta_o, mf.r());
    const float kappa = 1.0f - (Favg * E_o + Fms * (1.0f - E_o));
    if (uc < kappa) {
      V3 wi = sampleCosineHemisphere(u);
      if (wo.z < 0) wi = wi * -1.0f;
      const float cosTheta_i = wi.z;
      const float r = (1.0f - mat.ior) / (1.0f + mat.ior);
      const float F0 = r * r;
      const float cDiffuse = (1.0f - ggxBaseE(sc.lut, F0, mf.r(), cosTheta_o)) *
                             (1.0f - ggxBaseE(sc.lut, F0, mf.r(), cosTheta_i)) /
                             (kPi * (1.0f - ggxBaseEavg(sc.lut, F0, mf.r())));
      return BSDFSample(Reflected | Diffuse | (length2(emission) > 0.0f ? Emitted : 0), base * cDiffuse, emission, wi,
                        fabsf(wi.z) * cDiffuse, 1.0f);
    }
    if (mf.smooth()) {
      const float F = fresnelDielectric(wo.z, mat.ior);
      V3 wi(-wo.x, -wo.y, wo.z);
      return BSDFSample(Reflected | Specular, V3(F / fabsf(wi.z)), V3(), wi, F, 0.0f);
    }
    V3 wm = mf.sampleVisibleMicrofacet(wo, u);
    const V3 wi = reflect(wo, wm);
    const float cosTheta_i = wi.z;
    if (wo.z * wi.z < 0.0f) return BSDFSample();
    const float Fss = fresnelDielectric(dot(wo, wm), mat.ior);
    const float Mss = mf.mdf(wm) * mf.g(wo, wi) / (4 * cosTheta_o * cosTheta_i);
    const float Mms = (1.0f - E_o) * (1.0f - ggxE(sc.lut, cosTheta_i, mf.r())) / (kPi * (1.0f - Eavg));
    const float pdf = mf.vmdf(wo, wm) / (4 * absDot(wo, wm)) * Fss;
    return BSDFSample(Reflected | Glossy, V3(Fss * Mss + Fms * Mms), V3(), wi, pdf, mf.r());
  }

  // ---- clearcoat, parametric.cpp:732-832 ---------------------------------------------------
  YB_DEV V3 fClearcoat(V3 wo, V3 wi, const GGX& mf, float* Fc) const {
    if (mf.smooth()) return V3();
    const float cosTheta_o = fabsf(wo.z), cosTheta_i = fabsf(wi.z);
    if (cosTheta_i == 0 || cosTheta_o == 0) return V3();
    V3 wm = wo + wi;
    if (length2(wm) == 0.0f) return V3();
    wm = normalized(wm.z < 0.0f ? -wm : wm);
    const float Fss = fresnelDielectric(dot(wo, wm), 1.5f);
    const float Mss = mf.mdf(wm) * mf.g(wo, wi) / (4 * cosTheta_o * cosTheta_i);
    *Fc = rmax(fresnelDielectric(cosTheta_o, 1.5f), fresnelDielectric(cosTheta_i, 1.5f));
    return V3(Fss * Mss);
  }
  YB_DEV float pdfClearcoat(V3 wo, V3 wi, const GGX& mf, float* Fc) const {
    if (mf.smooth()) return 0;
    V3 wm = wo + wi;
    if (length2(wm) == 0.0f) return 0;
    wm = normalized(wm.z < 0.0f ? -wm : wm);
    const float Fss = fresnelDielectric(dot(wo, wm), 1.5f);
    *Fc = rmax(fresnelDielectric(fabsf(wo.z), 1.5f), fresnelDielectric(fabsf(wi.z), 1.5f));
    return Fss * mf.vmdf(wo, wm) / (4 * absDot(wo, wm));
  }
  YB_DEV BSDFSample sampleClearcoat(V3 wo, const GGX& mf, V2 u, float uc) const {
    const float cosTheta_o = wo.z;
    if (mf.smooth()) {
      const float F = fresnelDielectric(wo.z, mat.ior);
      V3 wi(-wo.x, -wo.y, wo.z);
      return BSDFSample(Reflected | Specular, V3(F / fabsf(wi.z)), V3(), wi, F, 0.0f);
    }
    V3 wm = mf.sampleVisibleMicrofacet(wo, u);
    const V3 wi = reflect(wo, wm);
    const float cosTheta_i = wi.z;
    if (wo.z * wi.z < 0.0f) return BSDFSample();
    const float Fss = fresnelDielectric(dot(wo, wm), 1.5f);
    const float Mss = mf.mdf(wm) * mf.g(wo, wi) / (4 * cosTheta_o * cosTheta_i);
    const float pdf = mf.vmdf(wo, wm) / (4 * absDot(wo, wm)) * Fss;
    return BSDFSample(Reflected | Glossy, V3(Fss * Mss), V3(), wi, pdf, mat.clearcoatRoughness);
  }

  // ---- fImpl / pdfImpl / sampleImpl, parametric.cpp:84-258 -------------------------------------
  // The uv-taking entry points evaluate the material's textures and forward to the MatEval ones; shade
  // evaluates them (and the local frame) once per hit and shares them between sample / f / pdf — the
  // reference recomputes identical values in each call (core/bsdf.cpp:5-41).
  YB_DEV V3 fImpl(V3 _wo, V3 _wi, V2 uv) const { return fImpl(_wo, _wi, evalMaterialTextures(sc, mat, uv)); }
  YB_DEV float pdfImpl(V3 wo, V3 wi, V2 uv) const { return pdfImpl(wo, wi, evalMaterialTextures(sc, mat, uv)); }
  YB_DEV BSDFSample sampleImpl(V3 _wo, V2 uv, V2 u, float uc, float uc2, bool regularized) const {
    return sampleImpl(_wo, uv, evalMaterialTextures(sc, mat, uv), u, uc, uc2, regularized);
  }

  YB_DEV V3 fImpl(V3 _wo, V3 _wi, const MatEval& e) const {
    GGX mf(e.r, mat.anisotropic);
    const float cMetallic = e.m;
    const float cDielectric = (1.0f - e.m) * e.t;
    const float cGlossy = (1.0f - e.m) * (1.0f - e.t);
    V3 wo = mul3x3(mat.localRotation, _wo), wi = mul3x3(mat.localRotation, _wi);
    V3 val;
    if (cMetallic > 0.0f) val += cMetallic * fMetallic(wo, wi, e.base, mf);
    if (cDielectric > 0.0f) val += cDielectric * fDielectric(wo, wi, e.base, mf);
    if (cGlossy > 0.0f) val += cGlossy * fGlossy(wo, wi, e.base, mf);
    if (e.c > 0.0f) {
      GGX mfClearcoat(e.cr);
      float Fc = 0.0f;
      V3 valClear = fClearcoat(wo, wi, mfClearcoat, &Fc);
      val = (1.0f - e.c * Fc) * val + e.c * valClear;
    }
    return val;
  }
  YB_DEV float pdfImpl(V3 wo, V3 wi, const MatEval& e) const {
    GGX mf(e.r, mat.anisotropic);
    const float pMetallic = e.m;
    const float pDielectric = (1.0f - e.m) * e.t;
    const float pGlossy = (1.0f - e.m) * (1.0f - e.t);
    float pdf = 0.0f;
    if (pMetallic > 0.0f) pdf += pMetallic * pdfMetallic(wo, wi, mf);
    if (pDielectric > 0.0f) pdf += pDielectric * pdfDielectric(wo, wi, mf);
    if (pGlossy > 0.0f) pdf += pGlossy * pdfGlossy(wo, wi, mf);
    if (e.c > 0.0f) {
      GGX mfClearcoat(e.cr);
      float Fc = 0.0f;
      float pdfClear = pdfClearcoat(wo, wi, mfClearcoat, &Fc);
      pdf = (1.0f - e.c * Fc) * pdf + e.c * pdfClear;
    }
    return pdf;
  }
  YB_DEV BSDFSample sampleImpl(V3 _wo, V2 uv, const MatEval& e, V2 u, float uc, float uc2, bool regularized) const {
    float r = e.r, cr = e.cr;
    if (regularized) {
      r = roughen(r);
      cr = roughen(cr);
    }
    GGX mfCoat(cr);
    V3 wmCoat = mfCoat.sampleVisibleMicrofacet(_wo, u);
    const float Favg = FavgFit(1.5f);
    const float Eavg = ggxEavg(sc.lut, cr);
    const float Fms = Favg * Favg * Eavg / (1.0f - Favg * (1.0f - Eavg));
    const float E_o = ggxE(sc.lut, absDot(_wo, wmCoat), cr);
    const float kappa = 1.0f - (Favg * E_o + Fms * (1.0f - E_o));
    const float pClearcoat = float(double(e.c) * (1.0 - double(kappa)));  // `c * (1.0 - kappa)` in double
    const float pMetallic = (1.0f - pClearcoat) * e.m;
    const float pDielectric = (1.0f - pClearcoat) * (e.m + (1.0f - e.m) * e.t);
    BSDFSample s;
    if (uc2 < pClearcoat) {
      s = sampleClearcoat(_wo, mfCoat, u, uc);
    } else {
      GGX mf(r, mat.anisotropic);
      V3 wo = mul3x3(mat.localRotation, _wo);
      if (uc2 < pMetallic) {
        s = sampleMetallic(wo, e.base, mf, u, uc);
      } else if (uc2 < pDielectric) {
        s = sampleDielectric(wo, e.base, mf, u, uc);
      } else {
        V3 emission(mat.emission);
        if (mat.hasEmission && mat.emisTex >= 0) emission *= sampleU8RGB(sc, mat.emisTex, uv);
        s = sampleGlossy(wo, e.base, emission, mf, u, uc);
      }
      s.wi = mul3x3(mat.invRotation, s.wi);
    }
    return s;
  }

  // ---- world-space wrappers, core/bsdf.cpp:5-41 ---------------------------------------------
  static YB_DEV Frame localFrame(V3 n, V3 t) { return length2(t) > 0 ? Frame(n, t) : Frame(n); }
  YB_DEV V3 f(V3 wo, V3 wi, V3 n, V3 t, V2 uv) const {
    Frame fr = localFrame(n, t);
    return fImpl(fr.wtl(wo), fr.wtl(wi), uv);
  }
  YB_DEV float pdf(V3 wo, V3 wi, V3 n, V3 t, V2 uv) const {
    Frame fr = localFrame(n, t);
    return pdfImpl(fr.wtl(wo), fr.wtl(wi), uv);
  }
  YB_DEV BSDFSample sample(V3 wo, V3 n, V3 t, V2 uv, V2 u, float uc, float uc2, bool regularized) const {
    Frame fr = localFrame(n, t);
    BSDFSample s = sampleImpl(fr.wtl(wo), uv, u, uc, uc2, regularized);
    s.wi = fr.ltw(s.wi);
    return s;
  }
  // parametric.cpp:834-838
  YB_DEV V3 attenuation(float d) const {
    if (mat.thinTransmission) return V3(1.0f);
    V3 e = (V3(mat.volumeColor) - 1.0f) * d * mat.volumeDensity;
    return V3(expfExact(e.x), expfExact(e.y), expfExact(e.z));
  }
};

}  // namespace yb
