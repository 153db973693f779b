// camera.cuh — primary-ray generation, the device side of yart::Camera.
//
// Restates reference src/core/camera.hpp:138-164 (Camera::getRay) on the derived members the host
// computes (ys_camera_make ↔ calcDerivedProperties, camera.hpp:25-59), with
// src/math/sampling.hpp:20-28 (pixelJitterGaussian, sigma = 0.3 px), :40-45 (sampleDiskUniform)
// and :72-89 (samplePolyUniform).  RayIntegrator::sample (src/cpu/ray-integrator.cpp:11-18) passes
// `getPixel2D()` and `get2D()` as call arguments; the x86-64 g++ oracle evaluates them right to
// left, so the LENS sample takes sampler dims 0-1 and the FILM sample dims 2-3.
#pragma once
#include "bsdf.cuh"
#include "sampler.cuh"

namespace yb {

// sampling.hpp:20-28
YB_DEV V2 pixelJitterGaussian(V2 u, float stdDev) {
  float a = sqrtf(-2.0f * logfExact(u.x)) * stdDev;
  float b = 2.0f * kPi * u.y;
  return V2(a * cosfExact(b), a * sinfExact(b));
}

// sampling.hpp:72-89
YB_DEV V2 samplePolyUniform(V2 u, uint32_t sides) {
  u.x *= float(sides);
  // math::min<uint32_t, uint32_t>(sides - 1, uint32_t(u.x))
  uint32_t ux = uint32_t(u.x);
  uint32_t side = (sides - 1) < ux ? (sides - 1) : ux;
  u.x -= float(side);
  V3 b = sampleTriUniform(u);
  float theta1 = float(side) / float(sides) * 2.0f * kPi;
  float theta2 = float(side + 1) / float(sides) * 2.0f * kPi;
  float c1 = cosfExact(theta1), s1 = sinfExact(theta1);
  float c2 = cosfExact(theta2), s2 = sinfExact(theta2);
  // float2(0,0)*b0 + float2(-s1,c1)*b1 + float2(-s2,c2)*b2, left to right
  V2 r = V2(0.0f, 0.0f) * b.x + V2(-s1, c1) * b.y;
  return r + V2(-s2, c2) * b.z;
}

// camera.hpp:138-164
YB_DEV void cameraRay(const YcCamera& c, uint32_t px, uint32_t py, V2 uvFilm, V2 uvLens, V3& origin, V3& dir) {
  V2 jitter = pixelJitterGaussian(uvFilm, 0.3f) + V2(float(px), float(py));
  V3 pixel = V3(c.topLeftPixel) + V3(c.pixelDeltaU) * jitter.x + V3(c.pixelDeltaV) * jitter.y;
  origin = V3(c.position);
  if (c.apertureRadius > 0.0f) {
    V2 s = c.apertureSides == 0 ? sampleDiskUniform(uvLens) : samplePolyUniform(uvLens, c.apertureSides);
    V3 lensPos(s.x, s.y, 0.0f);
    lensPos *= c.apertureRadius;
    V3 fx(c.frameX), fy(c.frameY), fz(c.frameZ);
    origin += lensPos.x * fx + lensPos.y * fy + lensPos.z * fz;  // Frame::ltw, frame.hpp:56-58
  }
  dir = normalized(pixel - origin);
}

// RayIntegrator::sample, ray-integrator.cpp:11-18 (lens draw first: see the header comment)
YB_DEV void primaryRay(const YcCamera& c, Sampler& smp, uint32_t px, uint32_t py, V3& origin, V3& dir) {
  V2 uvLens = smp.get2D();
  V2 uvFilm = smp.get2D();  // getPixel2D() == get2D() for SobolSampler (sampler.hpp:109-111)
  cameraRay(c, px, py, uvFilm, uvLens, origin, dir);
}

}  // namespace yb
