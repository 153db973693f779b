// glb.cpp — from-scratch binary-glTF (GLB) reader producing the same scene yart's gltf::load builds.
//
// The reference's loader (src/gltf/gltf.cpp:62-358) sits on fastgltf, which the reference does not
// vendor and this image does not have.  This file restates what gltf::load does with the parsed
// asset — it does not restate fastgltf — and emits the neutral SceneDesc the rest of the host layer
// (and the oracle driver) consume:
//   processMaterial  gltf.cpp:62-176   pbrMetallicRoughness + KHR_materials_{emissive_strength,
//                                      transmission, ior, anisotropy, clearcoat, volume};
//                                      thinTransmission is always true (:105); MR texture = channels {1,2}
//   loadTexture      core/texture.hpp:62-90   decode to RGBA8, pick channels, sRGB → gamma-2 8-bit
//   processMesh      gltf.cpp:178-270  all triangle primitives of a mesh merged into one Mesh;
//                                      POSITION, NORMAL, TEXCOORD_0 and indices are required (the
//                                      reference dereferences them unconditionally), TANGENT optional
//   processNode      gltf.cpp:272-317  TRS → translation * rotationFromQuat * scaling; children first;
//                                      one AreaLight per emissive triangle with the node's accumulated
//                                      transform `node.transform * globalTransform`; lightIdx restarts at
//                                      0 for every node while lights are appended globally (:301,309)
//   load             gltf.cpp:319-358  materials → meshes → scene nodes under an identity root
// Containers: binary .glb, and .gltf JSON whose buffers are files next to the asset or base64 data URIs
// (Options::LoadExternalBuffers, gltf.cpp:335).  Images: PNG and JPEG in bufferViews of the GLB's BIN chunk
// (images.cpp); `uri` images and images inside external buffers are dropped, as the reference drops them.
// Not covered (an error, or noted): sparse accessors, `matrix` nodes are used as given (fastgltf would decompose them
// to TRS first), cameras / skins / animations are ignored.
#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "hmath.hpp"
#include "scene_desc.hpp"

namespace yartb {

// ---------------------------------------------------------------------------------------
// minimal JSON
// ---------------------------------------------------------------------------------------
struct Json {
  enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
  double num = 0;
  bool b = false;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;

  const Json* get(const char* key) const {
    if (type != Obj) return nullptr;
    for (const auto& kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  double number(const char* key, double def) const {
    const Json* j = get(key);
    return j && j->type == Num ? j->num : def;
  }
  long index(const char* key) const {  // -1 when absent
    const Json* j = get(key);
    return j && j->type == Num && j->num >= 0.0 && j->num < 2147483648.0 ? long(j->num) : -1;  // NaN / negative / huge: absent
  }
  size_t size() const { return type == Arr ? arr.size() : 0; }
};

class JsonParser {
 public:
  JsonParser(const char* s, size_t n) : p_(s), end_(s + n) {}
  bool parse(Json& out, std::string& err) {
    begin_ = p_;
    if (!value(out, 0)) {
      err = "malformed JSON chunk near byte " + std::to_string(p_ - begin_);
      return false;
    }
    return true;
  }

 private:
  const char *p_, *end_, *begin_ = nullptr;
  void ws() {
    while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) p_++;
  }
  bool lit(const char* s) {
    size_t n = strlen(s);
    if (size_t(end_ - p_) < n || memcmp(p_, s, n) != 0) return false;
    p_ += n;
    return true;
  }
  bool string(std::string& out) {
    if (p_ >= end_ || *p_ != '"') return false;
    p_++;
    while (p_ < end_ && *p_ != '"') {
      if (*p_ == '\\') {
        if (++p_ >= end_) return false;
        switch (*p_) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            if (end_ - p_ < 5) return false;
            unsigned cp = 0;
            for (int i = 1; i <= 4; i++) {
              char c = p_[i];
              cp = cp * 16 + (c >= '0' && c <= '9' ? c - '0' : (c | 32) >= 'a' && (c | 32) <= 'f' ? (c | 32) - 'a' + 10 : 0);
            }
            p_ += 4;
            if (cp < 0x80) out += char(cp);
            else if (cp < 0x800) out += char(0xc0 | (cp >> 6)), out += char(0x80 | (cp & 0x3f));
            else out += char(0xe0 | (cp >> 12)), out += char(0x80 | ((cp >> 6) & 0x3f)), out += char(0x80 | (cp & 0x3f));
            break;
          }
          default: out += *p_;
        }
        p_++;
      } else {
        out += *p_++;
      }
    }
    if (p_ >= end_) return false;
    p_++;
    return true;
  }
  bool value(Json& out, int depth) {
    if (depth > 64) return false;
    ws();
    if (p_ >= end_) return false;
    if (*p_ == '{') {
      out.type = Json::Obj;
      p_++;
      ws();
      if (p_ < end_ && *p_ == '}') return p_++, true;
      while (true) {
        ws();
        std::string k;
        if (!string(k)) return false;
        ws();
        if (p_ >= end_ || *p_++ != ':') return false;
        Json v;
        if (!value(v, depth + 1)) return false;
        out.obj.emplace_back(std::move(k), std::move(v));
        ws();
        if (p_ >= end_) return false;
        if (*p_ == ',') { p_++; continue; }
        if (*p_ == '}') return p_++, true;
        return false;
      }
    }
    if (*p_ == '[') {
      out.type = Json::Arr;
      p_++;
      ws();
      if (p_ < end_ && *p_ == ']') return p_++, true;
      while (true) {
        Json v;
        if (!value(v, depth + 1)) return false;
        out.arr.push_back(std::move(v));
        ws();
        if (p_ >= end_) return false;
        if (*p_ == ',') { p_++; continue; }
        if (*p_ == ']') return p_++, true;
        return false;
      }
    }
    if (*p_ == '"') {
      out.type = Json::Str;
      return string(out.str);
    }
    if (lit("true")) return out.type = Json::Bool, out.b = true, true;
    if (lit("false")) return out.type = Json::Bool, out.b = false, true;
    if (lit("null")) return out.type = Json::Null, true;
    char* e = nullptr;
    std::string tmp(p_, size_t(std::min<ptrdiff_t>(end_ - p_, 64)));
    double v = strtod(tmp.c_str(), &e);
    if (e == tmp.c_str()) return false;
    p_ += e - tmp.c_str();
    out.type = Json::Num;
    out.num = v;
    return true;
  }
};

// ---------------------------------------------------------------------------------------
// minimal PNG → RGBA8 (what stbi_load_from_memory(..., 4) returns for the supported subset)
// ---------------------------------------------------------------------------------------
// image decoding (stb_image's results restated): images.cpp
bool decodeImageRGBA8(const uint8_t* data, size_t len, int& w, int& h, std::vector<uint8_t>& rgba, std::string& err);

// core/color-utils.hpp:12-15
static float sRGBDecode(float val) {
  if (val <= 0.04045f) return val / 12.92f;
  return std::pow((val + 0.055f) / 1.055f, 2.4f);
}

// loadTexture<C>, core/texture.hpp:62-90
static void convertTexture(const std::vector<uint8_t>& rgba, int w, int h, uint32_t type, int C, const int* channels,
                           ysc::TextureDesc& out) {
  out.channels = uint32_t(C), out.isFloat = 0, out.type = type, out.width = uint32_t(w), out.height = uint32_t(h);
  out.u8.resize(size_t(w) * h * C);
  for (size_t i = 0; i < size_t(w) * h; i++)
    for (int j = 0; j < C; j++) {
      uint8_t pixel = rgba[i * 4 + channels[j]];
      if (type == ysc::sRGB) {
        float val = sRGBDecode(float(pixel) / 255.0f);
        val = std::sqrt(val);
        pixel = uint8_t(val * 255.0f);
      }
      out.u8[i * C + j] = pixel;
    }
}

// ---------------------------------------------------------------------------------------
// the loader
// ---------------------------------------------------------------------------------------
namespace {

// A JSON number used as a size / offset: finite, non-negative, integral enough, and small enough that products with
// strides cannot wrap (files are < 2^40 bytes).
static bool toSize(double v, size_t& out) {
  if (!(v >= 0.0) || !(v < 1099511627776.0)) return false;  // also rejects NaN
  out = size_t(v);
  return true;
}

struct Glb {
  Json root;
  std::vector<uint8_t> bin;                      // the GLB's BIN chunk = buffer 0 of a .glb
  std::vector<std::vector<uint8_t>> external;    // buffers loaded from `uri` (fastgltf Options::LoadExternalBuffers, gltf.cpp:335)
  std::vector<bool> isExternal;                  // per buffer
  std::string baseDir;

  static bool base64(const std::string& in, std::vector<uint8_t>& out) {
    uint32_t acc = 0;
    int bits = 0;
    for (char ch : in) {
      int v;
      if (ch >= 'A' && ch <= 'Z') v = ch - 'A';
      else if (ch >= 'a' && ch <= 'z') v = ch - 'a' + 26;
      else if (ch >= '0' && ch <= '9') v = ch - '0' + 52;
      else if (ch == '+' || ch == '-') v = 62;
      else if (ch == '/' || ch == '_') v = 63;
      else if (ch == '=' || ch == '\n' || ch == '\r') continue;
      else return false;
      acc = (acc << 6) | uint32_t(v), bits += 6;
      if (bits >= 8) out.push_back(uint8_t(acc >> (bits -= 8)));
    }
    return true;
  }
  // `buffers[i].uri`: a file next to the asset or a base64 data URI.  Buffer 0 without a uri is the BIN chunk.
  bool loadBuffers() {
    const Json* bufs = root.get("buffers");
    const size_t n = bufs ? bufs->size() : 0;
    external.resize(n), isExternal.assign(n, false);
    for (size_t i = 0; i < n; i++) {
      const Json* uri = bufs->arr[i].get("uri");
      if (!uri || uri->type != Json::Str) continue;
      isExternal[i] = true;
      const std::string& u = uri->str;
      if (u.compare(0, 5, "data:") == 0) {
        const size_t comma = u.find(',');
        if (comma == std::string::npos || u.find(";base64") == std::string::npos || !base64(u.substr(comma + 1), external[i]))
          return err = "bad data URI in buffers", false;
      } else {
        if (u.find("..") != std::string::npos || (!u.empty() && u[0] == '/')) return err = "buffer uri must stay next to the asset", false;
        FILE* f = fopen((baseDir + u).c_str(), "rb");
        if (!f) return err = "cannot open external buffer " + u, false;
        uint8_t chunk[65536];
        for (size_t k; (k = fread(chunk, 1, sizeof chunk, f)) > 0;) external[i].insert(external[i].end(), chunk, chunk + k);
        fclose(f);
      }
    }
    return true;
  }
  std::string err;

  // `count` elements of `elem` bytes, `stride` apart, starting at `off`, inside a view of `n` bytes — written so that
  // nothing can wrap: off + (count - 1) * stride + elem <= n
  static bool spanFits(size_t off, size_t count, size_t stride, size_t elem, size_t n) {
    if (count == 0) return true;
    if (off > n || elem > n - off) return false;
    return stride == 0 ? count == 1 : (count - 1) <= (n - off - elem) / stride;
  }

  bool view(long bv, const uint8_t*& p, size_t& n, size_t& stride) {
    const Json* views = root.get("bufferViews");
    if (bv < 0 || !views || size_t(bv) >= views->size()) return err = "bufferView index out of range", false;
    const Json& v = views->arr[bv];
    const long bi = std::max<long>(0, v.index("buffer"));
    if (size_t(bi) >= isExternal.size() && bi != 0) return err = "bufferView names a missing buffer", false;
    const std::vector<uint8_t>& data = (size_t(bi) < isExternal.size() && isExternal[bi]) ? external[bi] : bin;
    if (&data == &bin && bi != 0) return err = "buffer without a uri that is not the GLB's BIN chunk", false;
    size_t off;
    if (!toSize(v.number("byteOffset", 0), off) || !toSize(v.number("byteLength", 0), n) || !toSize(v.number("byteStride", 0), stride))
      return err = "bufferView with a negative, non-finite or absurd offset / length / stride", false;
    if (off > data.size() || n > data.size() - off) return err = "bufferView exceeds its buffer", false;
    p = data.data() + off;
    return true;
  }
  // An image is decoded only when it sits in a bufferView of the GLB's own BIN chunk: the reference takes
  // `sources::BufferView` images whose buffer is a `sources::ByteView` and returns nullptr otherwise (gltf.cpp:33-41) —
  // external or data-URI buffers arrive as owned arrays there, and `uri` images are never opened.
  bool viewIsEmbedded(long bv) const {
    const Json* views = root.get("bufferViews");
    if (bv < 0 || !views || size_t(bv) >= views->size()) return false;
    const long bi = std::max<long>(0, views->arr[bv].index("buffer"));
    return !(size_t(bi) < isExternal.size() && isExternal[bi]);
  }

  // reads accessor `idx` as `comps` floats per element
  bool floats(long idx, int comps, std::vector<float>& out) {
    const Json* accs = root.get("accessors");
    if (idx < 0 || !accs || size_t(idx) >= accs->size()) return err = "accessor index out of range", false;
    const Json& a = accs->arr[idx];
    if (a.get("sparse")) return err = "sparse accessors are not supported", false;
    if (long(a.number("componentType", 0)) != 5126) return err = "vertex attributes must be float (5126)", false;
    const Json* ty = a.get("type");
    const int have = !ty ? 0 : ty->str == "SCALAR" ? 1 : ty->str == "VEC2" ? 2 : ty->str == "VEC3" ? 3 : ty->str == "VEC4" ? 4 : 0;
    if (have < comps) return err = "accessor has too few components", false;
    size_t count, off;
    if (!toSize(a.number("count", 0), count) || !toSize(a.number("byteOffset", 0), off)) return err = "accessor with a bad count / offset", false;
    const uint8_t* p;
    size_t n, stride;
    if (!view(a.index("bufferView"), p, n, stride)) return false;
    if (!stride) stride = size_t(have) * 4;
    if (!spanFits(off, count, stride, size_t(comps) * 4, n)) return err = "accessor exceeds its bufferView", false;
    out.resize(count * comps);
    for (size_t i = 0; i < count; i++) memcpy(&out[i * comps], p + off + i * stride, size_t(comps) * 4);
    return true;
  }

  bool indices(long idx, std::vector<uint32_t>& out) {
    const Json* accs = root.get("accessors");
    if (idx < 0 || !accs || size_t(idx) >= accs->size()) return err = "primitive without indices (required by gltf.cpp:246)", false;
    const Json& a = accs->arr[idx];
    const long ct = long(a.number("componentType", 0));
    const size_t sz = ct == 5121 ? 1 : ct == 5123 ? 2 : ct == 5125 ? 4 : 0;
    if (!sz) return err = "index accessor must be u8/u16/u32", false;
    size_t count, off;
    if (!toSize(a.number("count", 0), count) || !toSize(a.number("byteOffset", 0), off)) return err = "index accessor with a bad count / offset", false;
    const uint8_t* p;
    size_t n, stride;
    if (!view(a.index("bufferView"), p, n, stride)) return false;
    if (!stride) stride = sz;
    if (!spanFits(off, count, stride, sz, n)) return err = "index accessor exceeds its bufferView", false;
    out.resize(count);
    for (size_t i = 0; i < count; i++) {
      const uint8_t* q = p + off + i * stride;
      out[i] = sz == 1 ? q[0] : sz == 2 ? uint32_t(q[0] | (q[1] << 8)) : uint32_t(q[0] | (q[1] << 8) | (q[2] << 16) | (uint32_t(q[3]) << 24));
    }
    return true;
  }
};

// float4x4 * float4x4 (mat.hpp:262-273) and Transform::operator* (transform.hpp:52-57) live in hmath.hpp

static Mat4 trsMatrix(const float t[3], const float q[4], const float s[3]) {
  // rotationFromQuat, gltf.cpp:6-18: `half * 2.0f`
  const float qr = q[3], qi = q[0], qj = q[1], qk = q[2];
  Mat4 half;
  const float hm[16] = {0.5f - (qj * qj + qk * qk), (qi * qj - qr * qk), (qi * qk + qr * qj), 0.0f,
                        (qi * qj + qr * qk), 0.5f - (qi * qi + qk * qk), (qj * qk - qr * qi), 0.0f,
                        (qi * qk - qr * qj), (qj * qk + qr * qi), 0.5f - (qi * qi + qj * qj), 0.0f,
                        0.0f, 0.0f, 0.0f, 0.5f};
  for (int i = 0; i < 16; i++) half.m[i] = hm[i] * 2.0f;
  Mat4 T, S;
  T.m[3] = t[0], T.m[7] = t[1], T.m[11] = t[2];
  S.m[0] = s[0], S.m[5] = s[1], S.m[10] = s[2];
  return mul(mul(T, half), S);  // translation * rotation * scaling, gltf.cpp:285-288
}

struct Loader {
  Glb g;
  ysc::SceneDesc& d;
  std::vector<bool> materialEmissive;
  explicit Loader(ysc::SceneDesc& out) : d(out) {}

  int texture(long texIdx, uint32_t type, int C, const int* channels) {
    const Json* texs = g.root.get("textures");
    const Json* imgs = g.root.get("images");
    if (texIdx < 0 || !texs || size_t(texIdx) >= texs->size()) return -1;
    const long src = texs->arr[texIdx].index("source");
    if (src < 0 || !imgs || size_t(src) >= imgs->size()) return -1;  // `if (!gltfTex.imageIndex) return nullptr`
    const long bv = imgs->arr[src].index("bufferView");
    if (bv < 0) return -1;  // external URI: the reference returns nullptr too (gltf.cpp:33-34)
    if (!g.viewIsEmbedded(bv)) return -1;  // image inside an external / data-URI buffer: nullptr in the reference as well
    const uint8_t* p;
    size_t n, stride;
    if (!g.view(bv, p, n, stride)) return -2;
    int w, h;
    std::vector<uint8_t> rgba;
    if (!decodeImageRGBA8(p, n, w, h, rgba, g.err)) return -2;
    if (w < 2 || h < 2) return g.err = "textures must be at least 2x2", -2;
    ysc::TextureDesc t;
    convertTexture(rgba, w, h, type, C, channels, t);
    d.textures.push_back(std::move(t));
    return int(d.textures.size()) - 1;
  }

  bool materials() {
    const Json* mats = g.root.get("materials");
    static const int ch0123[4] = {0, 1, 2, 3}, ch12[2] = {1, 2};
    for (size_t i = 0; mats && i < mats->size(); i++) {
      const Json& m = mats->arr[i];
      ysc::MaterialDesc o;  // defaults below are glTF's / fastgltf's
      o.roughness = 1.0f, o.metallic = 1.0f, o.clearcoatRoughness = 0.03f, o.thinTransmission = 1;
      auto texOf = [&](const Json* parent, const char* key, uint32_t type, int C, const int* chn, int32_t& slot) {
        const Json* t = parent ? parent->get(key) : nullptr;
        if (!t) return true;
        int r = texture(t->index("index"), type, C, chn);
        if (r == -2) return false;
        slot = r;
        return true;
      };
      const Json* pbr = m.get("pbrMetallicRoughness");
      if (pbr) {
        if (const Json* f = pbr->get("baseColorFactor"))
          for (int k = 0; k < 3 && k < int(f->size()); k++) o.base[k] = float(f->arr[k].num);
        o.roughness = float(pbr->number("roughnessFactor", 1.0));
        o.metallic = float(pbr->number("metallicFactor", 1.0));
      }
      if (!texOf(pbr, "baseColorTexture", ysc::sRGB, 4, ch0123, o.baseTex)) return false;
      if (!texOf(pbr, "metallicRoughnessTexture", ysc::NonColor, 2, ch12, o.mrTex)) return false;
      const Json* ext = m.get("extensions");
      auto extension = [&](const char* name) { return ext ? ext->get(name) : nullptr; };
      if (const Json* tr = extension("KHR_materials_transmission")) {
        o.transmission = float(tr->number("transmissionFactor", 0.0));
        if (!texOf(tr, "transmissionTexture", ysc::NonColor, 1, ch0123, o.transTex)) return false;
      }
      if (const Json* an = extension("KHR_materials_anisotropy")) {
        o.anisotropic = float(an->number("anisotropyStrength", 0.0));
        o.anisoRotation = float(an->number("anisotropyRotation", 0.0));
      }
      if (const Json* cc = extension("KHR_materials_clearcoat")) {
        o.clearcoat = float(cc->number("clearcoatFactor", 0.0));
        o.clearcoatRoughness = float(cc->number("clearcoatRoughnessFactor", 0.0));
      }
      float em[3] = {0, 0, 0}, strength = 1.0f;
      if (const Json* f = m.get("emissiveFactor"))
        for (int k = 0; k < 3 && k < int(f->size()); k++) em[k] = float(f->arr[k].num);
      if (const Json* es = extension("KHR_materials_emissive_strength")) strength = float(es->number("emissiveStrength", 1.0));
      for (int k = 0; k < 3; k++) o.emission[k] = em[k] * strength;
      if (!texOf(&m, "emissiveTexture", ysc::sRGB, 3, ch0123, o.emisTex)) return false;
      if (const Json* nt = m.get("normalTexture")) {
        if (!texOf(&m, "normalTexture", ysc::NonColor, 3, ch0123, o.normalTex)) return false;
        o.normalScale = float(nt->number("scale", 1.0));
      }
      if (const Json* io = extension("KHR_materials_ior")) o.ior = float(io->number("ior", 1.5));
      if (const Json* vol = extension("KHR_materials_volume")) {
        if (const Json* c = vol->get("attenuationColor"))
          for (int k = 0; k < 3 && k < int(c->size()); k++) o.volumeColor[k] = float(c->arr[k].num);
        // fastgltf's default attenuationDistance is +infinity → density 0
        const Json* ad = vol->get("attenuationDistance");
        o.volumeDensity = ad ? 1.0f / float(ad->num) : 0.0f;
      }
      const f3 e(o.emission);
      materialEmissive.push_back(length2(e) > 0.0f);  // ParametricBSDF::emission(), parametric.hpp:43-45
      d.materials.push_back(o);
    }
    return true;
  }

  bool meshes() {
    const Json* ms = g.root.get("meshes");
    for (size_t mi = 0; ms && mi < ms->size(); mi++) {
      ysc::MeshDesc out;
      const Json* prims = ms->arr[mi].get("primitives");
      for (size_t pi = 0; prims && pi < prims->size(); pi++) {
        const Json& p = prims->arr[pi];
        const size_t idxOffset = out.nVerts();
        const long matIdx = std::max<long>(0, p.index("material"));  // materialIndex.value_or(0)
        if (size_t(matIdx) >= d.materials.size()) return g.err = "primitive references a missing material", false;
        if (long(p.number("mode", 4)) != 4) continue;  // triangles only (gltf.cpp:197)
        const Json* at = p.get("attributes");
        if (!at || at->index("POSITION") < 0 || at->index("NORMAL") < 0 || at->index("TEXCOORD_0") < 0)
          return g.err = "primitives need POSITION, NORMAL and TEXCOORD_0 (the reference reads them unconditionally)", false;
        std::vector<float> pos, nrm, uv, tan;
        if (!g.floats(at->index("POSITION"), 3, pos) || !g.floats(at->index("NORMAL"), 3, nrm) ||
            !g.floats(at->index("TEXCOORD_0"), 2, uv))
          return false;
        const size_t nv = pos.size() / 3;
        if (nrm.size() / 3 > nv || uv.size() / 2 > nv) return g.err = "attribute longer than POSITION", false;
        if (at->index("TANGENT") >= 0 && !g.floats(at->index("TANGENT"), 4, tan)) return false;
        out.positions.insert(out.positions.end(), pos.begin(), pos.end());
        for (size_t v = 0; v < nv; v++) {
          float vd[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
          if (v < nrm.size() / 3) memcpy(vd, &nrm[3 * v], 12);
          if (v < tan.size() / 4) memcpy(vd + 3, &tan[4 * v], 16);
          if (v < uv.size() / 2) memcpy(vd + 7, &uv[2 * v], 8);
          out.vertexData.insert(out.vertexData.end(), vd, vd + 9);
        }
        std::vector<uint32_t> idx;
        if (!g.indices(p.index("indices"), idx)) return false;
        for (size_t i = 0; i + 2 < idx.size(); i += 3) {
          for (int k = 0; k < 3; k++)
            if (idx[i + k] >= nv) return g.err = "index out of range", false;
          const uint32_t f[4] = {uint32_t(idx[i] + idxOffset), uint32_t(idx[i + 1] + idxOffset),
                                 uint32_t(idx[i + 2] + idxOffset), uint32_t(matIdx)};
          out.faces.insert(out.faces.end(), f, f + 4);
        }
      }
      out.lightIdx.assign(out.nFaces(), -1);
      d.meshes.push_back(std::move(out));
    }
    return true;
  }

  // processNode: returns false on error.  `global` = accumulated transform of the parents.
  bool node(long idx, int parentOut, const Transform& global, int depth) {
    const Json* ns = g.root.get("nodes");
    if (idx < 0 || !ns || size_t(idx) >= ns->size()) return g.err = "node index out of range", false;
    if (depth > 14) return g.err = "node hierarchy too deep", false;
    if (d.nodes.size() > (1u << 20)) return g.err = "more than 2^20 node instances (a node graph that is not a tree?)", false;
    const Json& n = ns->arr[idx];
    ysc::NodeDesc nd;
    nd.parent = parentOut;
    nd.mesh = int32_t(n.index("mesh"));
    if (nd.mesh >= int32_t(d.meshes.size())) return g.err = "node references a missing mesh", false;
    Mat4 m;
    if (const Json* mj = n.get("matrix")) {
      if (mj->size() != 16) return g.err = "node matrix must have 16 elements", false;
      // column-major in glTF.  (fastgltf would decompose this into TRS first; used as given here.)
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) m(r, c) = float(mj->arr[size_t(c) * 4 + r].num);
    } else {
      float t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
      auto rd = [&](const char* key, float* dst, int cnt) {
        const Json* j = n.get(key);
        for (int k = 0; j && k < cnt && k < int(j->size()); k++) dst[k] = float(j->arr[k].num);
      };
      rd("translation", t, 3), rd("rotation", q, 4), rd("scale", s, 3);
      m = trsMatrix(t, q, s);
    }
    nd.hasTransform = 1;
    memcpy(nd.m, m.m, sizeof nd.m);
    const int self = int(d.nodes.size());
    d.nodes.push_back(nd);
    const Transform own(m);
    // Transform::operator*: (m * rhs.m, rhs.inv * inv)
    const Transform local(mul(own.fwd, global.fwd), mul(global.inv, own.inv));
    if (const Json* ch = n.get("children"))
      for (const Json& c : ch->arr)
        if (!node(c.type == Json::Num && c.num >= 0.0 && c.num < 2147483648.0 ? long(c.num) : -1, self, local, depth + 1)) return false;
    if (nd.mesh >= 0) {
      ysc::MeshDesc& mesh = d.meshes[nd.mesh];
      int32_t li = 0;
      for (size_t i = 0; i < mesh.nFaces(); i++) {
        const uint32_t mat = mesh.faces[4 * i + 3];
        if (!materialEmissive[mat]) continue;
        ysc::LightDesc l;
        l.type = ysc::AreaLightT, l.mesh = nd.mesh, l.tri = int32_t(i);
        memcpy(l.emission, d.materials[mat].emission, 12);
        l.hasTransform = 1;
        memcpy(l.m, local.fwd.m, sizeof l.m);
        d.lights.push_back(l);
        mesh.lightIdx[i] = li++;  // sic: restarts per node (gltf.cpp:301,309)
      }
    }
    return true;
  }
};

}  // namespace

bool loadGlb(const std::string& path, ysc::SceneDesc& out, std::string* err) {
  auto fail = [&](const std::string& m) {
    if (err) *err = m + " in " + path;
    return false;
  };
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return fail("cannot open file");
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> buf(sz > 0 ? size_t(sz) : 0);
  const bool readOk = buf.empty() || fread(buf.data(), 1, buf.size(), f) == buf.size();
  fclose(f);
  if (!readOk || buf.size() < 2) return fail("truncated file");
  auto le32 = [&](size_t o) { return uint32_t(buf[o] | (buf[o + 1] << 8) | (buf[o + 2] << 16) | (uint32_t(buf[o + 3]) << 24)); };
  Loader L(out);
  const size_t slash = path.find_last_of('/');
  L.g.baseDir = slash == std::string::npos ? "" : path.substr(0, slash + 1);
  bool haveJson = false;
  if (buf.size() >= 20 && le32(0) == 0x46546c67u) {  // binary container
    if (le32(4) != 2) return fail("unsupported glTF version");
    size_t pos = 12;
    while (pos + 8 <= buf.size()) {
      const uint32_t n = le32(pos), type = le32(pos + 4);
      if (n > buf.size() || pos + 8 + n > buf.size()) return fail("truncated GLB chunk");
      if (type == 0x4e4f534au) {  // JSON
        JsonParser jp(reinterpret_cast<const char*>(&buf[pos + 8]), n);
        std::string e;
        if (!jp.parse(L.g.root, e)) return fail(e);
        haveJson = true;
      } else if (type == 0x004e4942u && L.g.bin.empty()) {  // BIN
        L.g.bin.assign(buf.begin() + long(pos) + 8, buf.begin() + long(pos) + 8 + n);
      }
      pos += 8 + size_t(n);
    }
  } else {  // a .gltf: the JSON document itself, geometry in external / data-URI buffers
    size_t k = 0;
    while (k < buf.size() && (buf[k] == ' ' || buf[k] == '\n' || buf[k] == '\r' || buf[k] == '\t' || buf[k] == 0xef || buf[k] == 0xbb || buf[k] == 0xbf)) k++;
    if (k >= buf.size() || buf[k] != '{') return fail("not a glTF file (neither the GLB magic nor a JSON document)");
    JsonParser jp(reinterpret_cast<const char*>(&buf[k]), buf.size() - k);
    std::string e;
    if (!jp.parse(L.g.root, e)) return fail(e);
    haveJson = true;
  }
  if (!haveJson) return fail("GLB without a JSON chunk");
  if (!L.g.loadBuffers()) return fail(L.g.err);
  if (!L.materials() || !L.meshes()) return fail(L.g.err);
  if (out.materials.empty()) {  // the reference would index material 0 out of range; give it a default one
    ysc::MaterialDesc def;
    def.roughness = 1.0f, def.metallic = 1.0f, def.thinTransmission = 1, def.clearcoatRoughness = 0.03f;
    out.materials.push_back(def);
    L.materialEmissive.push_back(false);
  }
  out.nodes.push_back(ysc::NodeDesc{});  // `Node root; Scene scene(std::move(root));`
  const long sceneIdx = std::max<long>(0, L.g.root.index("scene"));
  const Json* scenes = L.g.root.get("scenes");
  if (scenes && size_t(sceneIdx) < scenes->size())
    if (const Json* roots = scenes->arr[sceneIdx].get("nodes"))
      for (const Json& r : roots->arr)
        if (!L.node(r.type == Json::Num && r.num >= 0.0 && r.num < 2147483648.0 ? long(r.num) : -1, 0, Transform(), 1)) return fail(L.g.err);
  return true;
}

bool decodeTextureForTest(const uint8_t* png, size_t len, uint32_t type, int C, const int* channels, ysc::TextureDesc& out,
                          std::string& err) {
  int w, h;
  std::vector<uint8_t> rgba;
  if (!decodeImageRGBA8(png, len, w, h, rgba, err)) return false;
  convertTexture(rgba, w, h, type, C, channels, out);
  return true;
}

// writePPM, src/output/ppm.cpp:6-21: gamma 1/2.2, clamp, uint8(mapped * 255.999f)
bool writePpm(const std::string& path, const float* rgba, uint32_t w, uint32_t h) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  fprintf(f, "P6\n%u %u\n255\n", w, h);
  const float gamma = 1.0f / 2.2f;
  std::vector<uint8_t> row(size_t(w) * 3);
  for (uint32_t y = 0; y < h; y++) {
    for (uint32_t x = 0; x < w; x++)
      for (int c = 0; c < 3; c++) {
        const float v = std::pow(rgba[(size_t(y) * w + x) * 4 + c], gamma);
        const float mapped = v < 0.0f ? 0.0f : (1.0f < v ? 1.0f : v);  // std::clamp (NaN passes through → 0 byte on x86)
        row[size_t(x) * 3 + c] = (mapped != mapped) ? 0 : uint8_t(mapped * 255.999f);
      }
    fwrite(row.data(), 1, row.size(), f);
  }
  fclose(f);
  return true;
}

}  // namespace yartb
