// texture.cuh — manual bilinear texture sampling, bit-compatible with the reference.
//
// Follows src/core/texture.hpp:106-160 (getValue / Texture::sample) and src/core/texture.cpp:21-35
// (getXY): repeat wrap by `uv - floor(uv)`, scale by (w-1, h-1), texel = min(w-2, uint(x)),
// four taps (x,y) (x,y+1) (x+1,y) (x+1,y+1) and math_base.hpp:46-58 bilerp.  8-bit texels are
// /255 and squared when the texture is sRGB (gamma-2 storage, RGB channels only).  No mips and
// no hardware filtering: the texture units' 9-bit weights would not reproduce these bits.
#pragma once
#include "scene_dev.cuh"

namespace yb {

struct TexTaps {
  size_t i00, i01, i10, i11;  // texel indices of (x,y) (x,y+1) (x+1,y) (x+1,y+1)
  float fu, fv;
};

YB_DEV TexTaps texTaps(const YcTexture& t, V2 uv) {
  float ux = uv.x - floorf(uv.x);
  float uy = uv.y - floorf(uv.y);
  ux *= float(t.width - 1);
  uy *= float(t.height - 1);
  // math::min<uint32_t, float>(w - 2, uv.x): compare in float, result truncated to uint32
  uint32_t x = (float(t.width - 2) < ux) ? (t.width - 2) : uint32_t(ux);
  uint32_t y = (float(t.height - 2) < uy) ? (t.height - 2) : uint32_t(uy);
  TexTaps r;
  r.fu = ux - float(x);
  r.fv = uy - float(y);
  r.i00 = size_t(y * t.width + x);
  r.i01 = size_t((y + 1) * t.width + x);
  r.i10 = size_t(y * t.width + (x + 1));
  r.i11 = size_t((y + 1) * t.width + (x + 1));
  return r;
}

YB_DEV float bilerp1(float a0, float a1, float b0, float b1, float u, float v) {
  return a0 * (1.0f - u) * (1.0f - v) + a1 * (1.0f - u) * v + b0 * u * (1.0f - v) + b1 * u * v;
}

// float(b) / 255.0f for a byte b, correctly rounded, without the general division sequence: one
// Newton correction of b * fl(1/255) with fused multiply-adds (the standard division fast path; its
// range checks are unnecessary for 0..255).  Checked against the division for all 256 inputs
// (tests/test_host_logic.py::test_u8_unit_conversion_is_the_division).
YB_DEV float u8ToUnit(uint32_t b) {
  const float fb = float(b), r = 1.0f / 255.0f;
  const float q = fb * r;
  const float rem = fmaf(-q, 255.0f, fb);
  return fmaf(rem, r, q);
}

// one channel of an 8-bit texture
YB_DEV float texelU8(const DScene& s, const YcTexture& t, size_t idx, uint32_t c, bool gamma2) {
  float v = u8ToUnit(s.texU8[t.offset + idx * t.channels + c]);
  return gamma2 ? v * v : v;
}

YB_DEV float sampleU8Channel(const DScene& s, const YcTexture& t, const TexTaps& k, uint32_t c) {
  bool g = (t.type == 1u) && c < 3u;  // sRGB decode never touches alpha (texture.hpp:108-116)
  return bilerp1(texelU8(s, t, k.i00, c, g), texelU8(s, t, k.i01, c, g), texelU8(s, t, k.i10, c, g),
                 texelU8(s, t, k.i11, c, g), k.fu, k.fv);
}

YB_DEV V3 sampleU8RGB(const DScene& s, int tex, V2 uv) {
  const YcTexture t = s.textures[tex];
  TexTaps k = texTaps(t, uv);
  return V3(sampleU8Channel(s, t, k, 0), sampleU8Channel(s, t, k, 1), sampleU8Channel(s, t, k, 2));
}
YB_DEV float sampleU8Mono(const DScene& s, int tex, V2 uv, uint32_t channel = 0) {
  const YcTexture t = s.textures[tex];
  TexTaps k = texTaps(t, uv);
  return sampleU8Channel(s, t, k, channel);
}
YB_DEV V2 sampleU8RG(const DScene& s, int tex, V2 uv) {
  const YcTexture t = s.textures[tex];
  TexTaps k = texTaps(t, uv);
  return V2(sampleU8Channel(s, t, k, 0), sampleU8Channel(s, t, k, 1));
}
YB_DEV V3 sampleHDR(const DScene& s, int tex, V2 uv) {
  const YcTexture t = s.textures[tex];
  TexTaps k = texTaps(t, uv);
  const float* d = s.texF32 + t.offset;
  V3 a0(d + 3 * k.i00), a1(d + 3 * k.i01), b0(d + 3 * k.i10), b1(d + 3 * k.i11);
  return V3(bilerp1(a0.x, a1.x, b0.x, b1.x, k.fu, k.fv), bilerp1(a0.y, a1.y, b0.y, b1.y, k.fu, k.fv),
            bilerp1(a0.z, a1.z, b0.z, b1.z, k.fu, k.fv));
}

// ParametricBSDF::alpha, parametric.cpp:68-72
YB_DEV float materialAlpha(const DScene& s, const YcMaterial& m, V2 uv) {
  if (m.hasAlpha && m.baseTex >= 0) return sampleU8Mono(s, m.baseTex, uv, 3);
  return 1.0f;
}
// ParametricBSDF::base, parametric.cpp:74-77
YB_DEV V3 materialBase(const DScene& s, const YcMaterial& m, V2 uv) {
  V3 b(m.base);
  if (m.baseTex >= 0) return b * sampleU8RGB(s, m.baseTex, uv);
  return b;
}

}  // namespace yb
