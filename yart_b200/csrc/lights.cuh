// lights.cuh — device lights, environment importance sampling and the power light sampler.
//
// Restates reference src/core/light.cpp:16-81 (AreaLight), :83-135 (UniformInfiniteLight),
// :137-243 (ImageInfiniteLight), src/math/sampling.cpp:5-60 (PiecewiseConstant1D/2D::sample/pdf,
// including the `m_cdf[0 + 1]` in-cell offset quirk), src/core/light-sampler.cpp:52-93
// (PowerLightSampler) and src/math/math_base.hpp:106-119 (findFirst).
#pragma once
#include "bsdf.cuh"

namespace yb {

struct LightSample {
  V3 Li, wi, p, n;
  float pdf;
  YB_DEV LightSample() : pdf(0.f) {}
};

// sampling.cpp:5-36 on one 1-D distribution stored as func[n], cdf[n+1], integral
YB_DEV float sampleDistribution1D(const float* func, const float* cdf, float integral, int n, float u, float* pdf,
                                  uint32_t* offset) {
  long long size = (long long)(n + 1) - 2, first = 1, half, middle;
  while (size > 0) {
    half = size >> 1;
    middle = first + half;
    if (cdf[middle] < u) {
      first = middle + 1;
      size -= half + 1;
    } else {
      size = half;
    }
  }
  long long o = first - 1;
  if (o < 0) o = 0;
  if (o > (long long)(n + 1) - 2) o = (long long)(n + 1) - 2;
  *offset = uint32_t(o);
  float du = u - cdf[o];
  if (cdf[0 + 1] - cdf[o] > 0) du /= cdf[0 + 1] - cdf[o];  // sic: reference quirk, sampling.cpp:28
  *pdf = (integral > 0) ? func[o] / integral : 0.0f;
  return lerpf(0.0f, 1.0f, (float(o) + du) / float(n));
}

struct EnvDist {
  const float *func, *cdf, *rowInt, *mfunc, *mcdf;
  float mInt;
  uint32_t w, h;
  YB_DEV EnvDist(const DScene& sc, const YcLight& l) : w(l.distW), h(l.distH) {
    func = sc.envDist + l.distOffset;
    cdf = func + size_t(w) * h;
    rowInt = cdf + size_t(w + 1) * h;
    mfunc = rowInt + h;
    mcdf = mfunc + h;
    mInt = mcdf[h + 1];
  }
  // PiecewiseConstant2D::sample, sampling.cpp:38-47
  YB_DEV V2 sample(V2 u, float* pdf) const {
    float p0, p1;
    uint32_t ux, uy;
    float d1 = sampleDistribution1D(mfunc, mcdf, mInt, int(h), u.y, &p1, &uy);
    float d0 = sampleDistribution1D(func + size_t(uy) * w, cdf + size_t(uy) * (w + 1), rowInt[uy], int(w), u.x, &p0, &ux);
    *pdf = p0 * p1;
    return V2(d0, d1);
  }
  // PiecewiseConstant2D::pdf, sampling.cpp:49-60 (domain {0,0}-{1,1})
  YB_DEV float pdf(V2 uv) const {
    V2 p((uv.x - 0.0f) / 1.0f, (uv.y - 0.0f) / 1.0f);
    uint32_t iu = uint32_t(p.x * float(w));
    uint32_t iv = uint32_t(p.y * float(h));
    if (iu > w - 1) iu = w - 1;
    if (iv > h - 1) iv = h - 1;
    return func[size_t(iv) * w + iu] / mInt;
  }
};

YB_DEV bool unitSquareIncludes(V2 v) { return !(v.x < 0.0f || v.x > 1.0f || v.y < 0.0f || v.y > 1.0f); }

// Light::Le(uv) for infinite lights (light.cpp:98-100, 198-203)
YB_DEV V3 lightLe(const DScene& sc, const YcLight& l, V2 uv) {
  if (l.type == YC_LIGHT_IMAGE_INFINITE) {
    if (!unitSquareIncludes(uv)) return V3();
    return sampleHDR(sc, l.hdrTex, uv);
  }
  return V3(l.emission);
}

// Light::pdf(wi) (light.cpp:42-44, 106-111, 210-217)
YB_DEV float lightPdf(const DScene& sc, const YcLight& l, V3 wi) {
  if (l.type == YC_LIGHT_AREA) return 1.0f / l.area;
  if (l.type == YC_LIGHT_UNIFORM_INFINITE) return 0.0f;
  V2 uv = octahedralUV(xformRows(l.envInv, wi, 0.0f));
  if (!unitSquareIncludes(uv)) return 0.0f;
  EnvDist d(sc, l);
  return d.pdf(uv) / (4.0f * kPi);
}

// Light::sample (light.cpp:46-72, 113-135, 219-239)
YB_DEV LightSample lightSample(const DScene& sc, const YcLight& l, V3 p, V2 u) {
  LightSample s;
  if (l.type == YC_LIGHT_AREA) {
    const V3 b = sampleTriUniform(u);
    V3 pos = b.x * V3(l.p0) + b.y * V3(l.p1) + b.z * V3(l.p2);
    V3 normal = b.x * V3(l.n0) + b.y * V3(l.n1) + b.z * V3(l.n2);
    pos = xformRows(l.fwd, pos, 1.0f);
    normal = normalized(mul3x3(l.nrm, normal));
    s.Li = V3(l.emission);
    s.wi = normalized(pos - p);
    s.p = pos;
    s.n = normal;
    s.pdf = 1.0f / l.area;
  } else if (l.type == YC_LIGHT_IMAGE_INFINITE) {
    EnvDist d(sc, l);
    float pdf;
    V2 uv = d.sample(u, &pdf);
    if (pdf == 0.0f) return s;
    V3 wi = xformRows(l.envFwd, invOctahedralUV(uv), 0.0f);
    pdf /= l.surfaceArea;
    s.Li = lightLe(sc, l, uv);
    s.wi = wi;
    s.p = wi * 2.0f * l.sceneRadius;
    s.n = -wi;
    s.pdf = pdf;
  }
  // UniformInfiniteLight::sample returns {} (light.cpp:113-121)
  return s;
}

struct PickedLight {
  uint32_t index;
  float p;
};

YB_DEV float pInfinite(const DScene& sc) {
  return sc.nArea == 0 ? 1.0f : float(sc.nInf) / float(sc.nInf + 1);
}

// UniformLightSampler::sample / p, light-sampler.cpp:11-31.  `size_t(u * nl - 0.01f)` as x86-64 evaluates it
// (truncation; a negative value in (-1, 0) gives 0).  Variants build only: the reference's MISIntegrator hard-codes
// PowerLightSampler (mis-integrator.hpp:20), so this one can be checked against the reference's class, not its frames.
YB_DEV PickedLight pickLightUniform(const DScene& sc, float u) {
  const uint32_t nl = sc.nLights;
  const float x = u * float(nl) - 0.01f;
  uint32_t idx = x != x ? 0u : uint32_t((long long)x);
  if (nl - 1 < idx) idx = nl - 1;
  PickedLight r;
  r.index = idx;
  r.p = 1.0f / float(nl);
  return r;
}

// PowerLightSampler::sample, light-sampler.cpp:52-78
YB_DEV PickedLight pickLight(const DScene& sc, float u) {
#ifdef YB_RNG_SAMPLERS
  if (sc.uniformLights) return pickLightUniform(sc, u);
#endif
  const uint32_t infCount = sc.nInf;
  const float pInf = pInfinite(sc);
  PickedLight r;
  if (u < pInf) {
    u /= pInf;
    uint32_t idx = uint32_t(u * float(infCount));
    if (infCount - 1 < idx) idx = infCount - 1;
    r.index = sc.infLights[idx];
    r.p = pInf / float(infCount);
    return r;
  }
  u = (u - pInf) / (1.0f - pInf);
  u *= sc.totalPower;
  // findFirst(size, cdf[i] < u), math_base.hpp:106-119
  long long size = sc.nArea;
  long long sz = size - 1, first = 0;
  while (sz > 0) {
    long long half = sz >> 1, middle = first + half;
    bool res = sc.powerCdf[middle] < u;
    first = res ? (middle + 1) : first;
    sz = res ? sz - (half + 1) : half;
  }
  if (first < 0) first = 0;
  if (first > size - 1) first = size - 1;
  r.index = sc.areaLights[first];
  r.p = sc.lights[r.index].power / sc.totalPower * (1.0f - pInf);
  return r;
}

// PowerLightSampler::p, light-sampler.cpp:80-93
YB_DEV float lightPickProbability(const DScene& sc, uint32_t lightIdx) {
#ifdef YB_RNG_SAMPLERS
  if (sc.uniformLights) return 1.0f / float(sc.nLights);
#endif
  const float pInf = pInfinite(sc);
  const YcLight& l = sc.lights[lightIdx];
  if (l.type != YC_LIGHT_AREA) return pInf / float(sc.nInf);
  return l.power / sc.totalPower * (1.0f - pInf);
}

}  // namespace yb
