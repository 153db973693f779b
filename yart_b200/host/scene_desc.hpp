// scene_desc.hpp — neutral, data-only description of a yart scene (".ysc" file).
//
// This is the interchange between the synthetic-scene generators (tools/scenes.py),
// the product host library (which turns it into yartb::Scene → flattened GPU layout)
// and the oracle driver (which turns it into a reference yart::Scene through the
// reference's public C++ API).  It mirrors what yart's scene API takes:
//   ParametricBSDF ctor arguments        (reference src/bsdf/parametric.hpp:15-36)
//   Mesh(vertices, vertexData, faces)    (reference src/core/mesh.hpp:54-61)
//   Node(mesh*) + transform + children   (reference src/core/scene.hpp:11-64)
//   AreaLight / ImageInfiniteLight / UniformInfiniteLight ctor arguments
//                                        (reference src/core/light.hpp:76-171)
//   Texture<T,C>(w,h,type) + data        (reference src/core/texture.hpp:21-52)
// It contains no arithmetic — only containers and a reader/writer.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace ysc {

enum TextureType : uint32_t { LinearRGB = 0, sRGB = 1, NonColor = 2 };

struct TextureDesc {
  uint32_t channels = 0;  // 1,2,3,4
  uint32_t isFloat = 0;   // 1 → HDR float data (channels must be 3)
  uint32_t type = 0;      // TextureType
  uint32_t width = 0, height = 0;
  std::vector<uint8_t> u8;
  std::vector<float> f32;
};

struct MaterialDesc {
  float base[3] = {1, 1, 1};
  int32_t baseTex = -1, mrTex = -1, transTex = -1, normalTex = -1, ccTex = -1, emisTex = -1;
  float metallic = 0, roughness = 0, transmission = 0, ior = 1.5f;
  float anisotropic = 0, anisoRotation = 0, clearcoat = 0, clearcoatRoughness = 0;
  float emission[3] = {0, 0, 0};
  float normalScale = 1;
  int32_t thinTransmission = 0;
  float volumeColor[3] = {1, 1, 1};
  float volumeDensity = 0;
};

struct MeshDesc {
  std::vector<float> positions;   // 3 per vertex
  std::vector<float> vertexData;  // 9 per vertex: normal(3) tangent(4) uv(2)
  std::vector<uint32_t> faces;    // 4 per face: i0 i1 i2 material
  std::vector<int32_t> lightIdx;  // 1 per face, -1 = not a light
  size_t nVerts() const { return positions.size() / 3; }
  size_t nFaces() const { return faces.size() / 4; }
};

struct NodeDesc {
  int32_t parent = -1;  // node 0 is the root; parents precede children; sibling order = file order
  int32_t mesh = -1;
  int32_t hasTransform = 0;  // 0 → default (identity) Transform
  float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};  // row-major
};

enum LightType : int32_t { AreaLightT = 0, ImageInfiniteT = 1, UniformInfiniteT = 2 };

struct LightDesc {
  int32_t type = 0;
  int32_t mesh = -1, tri = -1;  // area
  float emission[3] = {0, 0, 0};  // area / uniform
  int32_t hasTransform = 0;
  float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  int32_t twoSided = 0;
  float sceneRadius = 100;  // infinite
  int32_t hdrTex = -1;      // image infinite
};

struct SceneDesc {
  std::vector<TextureDesc> textures;
  std::vector<MaterialDesc> materials;
  std::vector<MeshDesc> meshes;
  std::vector<NodeDesc> nodes;
  std::vector<LightDesc> lights;
};

namespace detail {
template <typename T>
inline bool rd(FILE* f, T* v, size_t n = 1) {
  return fread(v, sizeof(T), n, f) == n;
}
template <typename T>
inline bool wr(FILE* f, const T* v, size_t n = 1) {
  return fwrite(v, sizeof(T), n, f) == n;
}
}  // namespace detail

// File layout (little endian): "YSC1", then counts + raw arrays in the order below.
inline bool load(const std::string& path, SceneDesc& s, std::string* err = nullptr) {
  using namespace detail;
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    if (err) *err = "cannot open " + path;
    return false;
  }
  auto fail = [&](const char* m) {
    if (err) *err = std::string(m) + " in " + path;
    fclose(f);
    return false;
  };
  // every count read from the file is checked against the bytes the file still holds before anything is sized by it
  fseek(f, 0, SEEK_END);
  const long fileSize = ftell(f);
  fseek(f, 0, SEEK_SET);
  auto holds = [&](uint64_t count, uint64_t elemBytes) {
    const long at = ftell(f);
    if (fileSize < 0 || at < 0 || at > fileSize) return false;
    const uint64_t left = uint64_t(fileSize - at);
    return elemBytes == 0 || count <= left / elemBytes;
  };
  char magic[4];
  if (!rd(f, magic, 4) || memcmp(magic, "YSC1", 4) != 0) return fail("bad magic");
  uint32_t n;
  if (!rd(f, &n)) return fail("truncated");
  if (!holds(n, 20)) return fail("texture count exceeds the file");
  s.textures.resize(n);
  for (auto& t : s.textures) {
    uint32_t hdr[5];
    if (!rd(f, hdr, 5)) return fail("truncated texture header");
    t.channels = hdr[0], t.isFloat = hdr[1], t.type = hdr[2], t.width = hdr[3], t.height = hdr[4];
    if (t.channels == 0 || t.channels > 4 || t.width > (1u << 20) || t.height > (1u << 20)) return fail("bad texture header");
    size_t cnt = size_t(t.width) * t.height * t.channels;
    if (!holds(cnt, t.isFloat ? 4 : 1)) return fail("texture size exceeds the file");
    if (t.isFloat) {
      t.f32.resize(cnt);
      if (!rd(f, t.f32.data(), cnt)) return fail("truncated texture data");
    } else {
      t.u8.resize(cnt);
      if (!rd(f, t.u8.data(), cnt)) return fail("truncated texture data");
    }
  }
  if (!rd(f, &n)) return fail("truncated");
  if (!holds(n, sizeof(MaterialDesc))) return fail("material count exceeds the file");
  s.materials.resize(n);
  static_assert(sizeof(MaterialDesc) == 26 * 4, "MaterialDesc must be packed 4-byte fields");
  if (n && !rd(f, s.materials.data(), n)) return fail("truncated materials");
  if (!rd(f, &n)) return fail("truncated");
  if (!holds(n, 8)) return fail("mesh count exceeds the file");
  s.meshes.resize(n);
  for (auto& m : s.meshes) {
    uint32_t nv, nf;
    if (!rd(f, &nv) || !rd(f, &nf)) return fail("truncated mesh header");
    if (!holds(uint64_t(nv) * 12 + uint64_t(nf) * 5, 4)) return fail("mesh size exceeds the file");
    m.positions.resize(size_t(nv) * 3);
    m.vertexData.resize(size_t(nv) * 9);
    m.faces.resize(size_t(nf) * 4);
    m.lightIdx.resize(nf);
    if (!rd(f, m.positions.data(), m.positions.size()) || !rd(f, m.vertexData.data(), m.vertexData.size()) ||
        !rd(f, m.faces.data(), m.faces.size()) || !rd(f, m.lightIdx.data(), m.lightIdx.size()))
      return fail("truncated mesh data");
  }
  if (!rd(f, &n)) return fail("truncated");
  if (!holds(n, sizeof(NodeDesc))) return fail("node count exceeds the file");
  s.nodes.resize(n);
  static_assert(sizeof(NodeDesc) == 19 * 4, "NodeDesc must be packed");
  if (n && !rd(f, s.nodes.data(), n)) return fail("truncated nodes");
  if (!rd(f, &n)) return fail("truncated");
  if (!holds(n, sizeof(LightDesc))) return fail("light count exceeds the file");
  s.lights.resize(n);
  static_assert(sizeof(LightDesc) == 26 * 4, "LightDesc must be packed");
  if (n && !rd(f, s.lights.data(), n)) return fail("truncated lights");
  fclose(f);
  return true;
}

inline bool save(const std::string& path, const SceneDesc& s) {
  using namespace detail;
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  wr(f, "YSC1", 4);
  uint32_t n = uint32_t(s.textures.size());
  wr(f, &n);
  for (auto& t : s.textures) {
    uint32_t hdr[5] = {t.channels, t.isFloat, t.type, t.width, t.height};
    wr(f, hdr, 5);
    if (t.isFloat) wr(f, t.f32.data(), t.f32.size());
    else wr(f, t.u8.data(), t.u8.size());
  }
  n = uint32_t(s.materials.size());
  wr(f, &n);
  wr(f, s.materials.data(), n);
  n = uint32_t(s.meshes.size());
  wr(f, &n);
  for (auto& m : s.meshes) {
    uint32_t nv = uint32_t(m.nVerts()), nf = uint32_t(m.nFaces());
    wr(f, &nv);
    wr(f, &nf);
    wr(f, m.positions.data(), m.positions.size());
    wr(f, m.vertexData.data(), m.vertexData.size());
    wr(f, m.faces.data(), m.faces.size());
    wr(f, m.lightIdx.data(), m.lightIdx.size());
  }
  n = uint32_t(s.nodes.size());
  wr(f, &n);
  wr(f, s.nodes.data(), n);
  n = uint32_t(s.lights.size());
  wr(f, &n);
  wr(f, s.lights.data(), n);
  fclose(f);
  return true;
}

}  // namespace ysc
