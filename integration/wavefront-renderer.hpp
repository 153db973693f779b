// src/cuda/wavefront-renderer.hpp — the file a yart maintainer adds to teofum/yart to render on B200 through
// libyart_b200.so (C ABI: include/yart_cuda.h).  It is compiled and tested in this repository against the
// UNMODIFIED reference sources (oracle/Makefile, target `adapter`; tests/test_adapter.py renders a scene built
// through yart's own API with cpu::TileRenderer and with this class in the same process and compares the frames).
//
//   yart::cuda::WavefrontRenderer : yart::Renderer        drop-in for cpu::TileRenderer<Sampler, Integrator>
//       same knobs (samples, firstWaveSamples, maxWaveSamples, tileSize, tonemapper, backgroundColor, scene),
//       same callbacks (onRenderWaveComplete / onRenderTileComplete / onRenderComplete / onRenderAborted),
//       same render() / abort() / wait() / renderSync()                     (src/cpu/tile-renderer.hpp:25-116)
//   yart::cuda::SceneFlattener                             yart::Scene → YcScene (node tree in DFS pre-order,
//       the Mesh's own SahBVH re-laid out with both child boxes per node, leaf-ordered triangles, materials,
//       textures, lights with the constants the reference's constructors derived, PowerLightSampler tables)
//
// The flattener reads private members of a few reference classes.  In-tree that is one line per class,
//     friend class yart::cuda::SceneFlattener;
// in Scene (scene.hpp:66), Transform (transform.hpp:34), BVH (bvh.hpp:39), Mesh (mesh.hpp:15), BSDF (bsdf.hpp:44),
// ParametricBSDF (parametric.hpp:15), Camera (camera.hpp:10), AreaLight / UniformInfiniteLight / ImageInfiniteLight
// (light.hpp:75,119,145), samplers::PiecewiseConstant1D / 2D (sampling.hpp:118,160) — or, as this repository's build
// does to leave the reference untouched, compiling this translation unit with g++ -fno-access-control.
#pragma once
#include <atomic>
#include <map>
#include <thread>

#include <core/core.hpp>
#include <bsdf/parametric.hpp>

#include <yart_cuda.h>

namespace yart::cuda {

class SceneFlattener {
public:
  // Owns every array YcScene points into; valid as long as the flattener lives.
  std::vector<YcNode> nodes;
  std::vector<YcMesh> meshes;
  std::vector<YcBvhNode> bvhNodes;
  std::vector<YcBvhTri> bvhTris;
  std::vector<float> positions, normals, tangents, uvs;
  std::vector<uint32_t> primIndices, primMaterial;
  std::vector<int32_t> primLight;
  std::vector<YcMaterial> materials;
  std::vector<YcTexture> textures;
  std::vector<uint8_t> texelsU8;
  std::vector<float> texelsF32;
  std::vector<YcLight> lights;
  std::vector<float> envDist;
  std::vector<uint32_t> infiniteLights, areaLights;
  std::vector<float> lightPowerCdf;
  YcScene flat{};
  std::string error;

  bool flatten(const Scene& scene) {
    m_scene = &scene;
    // meshes, in the scene's order (Mesh::material(i) indexes the scene's material list, gltf.cpp:300-316)
    for (const auto& m : scene.m_meshes) {
      m_meshIndex[m.get()] = int32_t(meshes.size());
      if (!addMesh(*m)) return false;
    }
    for (const auto& b : scene.m_materials) {
      const auto* p = dynamic_cast<const ParametricBSDF*>(b.get());
      if (!p) return fail("only ParametricBSDF materials are supported");
      addMaterial(*p);
    }
    // triangle flags need the materials
    for (size_t mi = 0; mi < meshes.size(); mi++) {
      const YcMesh& ym = meshes[mi];
      for (uint32_t i = 0; i < ym.nTris; i++) {
        YcBvhTri& t = bvhTris[size_t(ym.triOffset) + i];
        const uint32_t mat = primMaterial[size_t(ym.primOffset) + t.prim];
        if (mat >= materials.size()) return fail("triangle material index out of range");
        const YcMaterial& m = materials[mat];
        t.flags |= (m.hasAlpha ? YC_TRI_ALPHA : 0u) | ((m.thinTransmission && m.transmission > 0.0f) ? YC_TRI_TRANSPARENT : 0u);
      }
    }
    addNode(scene.root(), -1, 0);
    if (!error.empty()) return false;
    for (const auto& l : scene.m_lights)
      if (!addLight(*l)) return false;
    // PowerLightSampler::init, light-sampler.cpp:32-50
    float totalPower = 0.0f;
    for (size_t i = 0; i < lights.size(); i++) {
      if (lights[i].type != YC_LIGHT_AREA) {
        infiniteLights.push_back(uint32_t(i));
      } else {
        areaLights.push_back(uint32_t(i));
        lightPowerCdf.push_back(totalPower + lights[i].power);
        totalPower += lights[i].power;
      }
    }
    int anyAlpha = 0;
    for (const YcMaterial& m : materials) anyAlpha |= m.hasAlpha;

    flat.nodes = nodes.data(), flat.nNodes = uint32_t(nodes.size());
    flat.meshes = meshes.data(), flat.nMeshes = uint32_t(meshes.size());
    flat.bvhNodes = bvhNodes.data(), flat.nBvhNodes = bvhNodes.size();
    flat.bvhTris = bvhTris.data(), flat.nBvhTris = bvhTris.size();
    flat.positions = positions.data(), flat.normals = normals.data();
    flat.tangents = tangents.data(), flat.uvs = uvs.data(), flat.nVerts = positions.size() / 3;
    flat.primIndices = primIndices.data(), flat.primMaterial = primMaterial.data();
    flat.primLight = primLight.data(), flat.nPrims = primMaterial.size();
    flat.materials = materials.data(), flat.nMaterials = uint32_t(materials.size());
    flat.textures = textures.data(), flat.nTextures = uint32_t(textures.size());
    flat.texelsU8 = texelsU8.data(), flat.nTexelsU8 = texelsU8.size();
    flat.texelsF32 = texelsF32.data(), flat.nTexelsF32 = texelsF32.size();
    flat.lights = lights.data(), flat.nLights = uint32_t(lights.size());
    flat.envDist = envDist.data(), flat.nEnvDist = envDist.size();
    flat.infiniteLights = infiniteLights.data(), flat.nInfinite = uint32_t(infiniteLights.size());
    flat.areaLights = areaLights.data(), flat.nArea = uint32_t(areaLights.size());
    flat.lightPowerCdf = lightPowerCdf.data();
    flat.totalPower = totalPower;
    size_t nLut = 0;
    flat.lutTables = ys_lut_tables(&nLut);  // the library's copy of bsdf/luts.hpp + the Sobol tables it needs
    flat.hasAlpha = anyAlpha;
    return flat.lutTables != nullptr;
  }

  // Camera's derived members (camera.hpp:17-22, filled by calcDerivedProperties :25-59)
  static YcCamera camera(const Camera& c) {
    YcCamera y{};
    put3(y.position, c.m_position);
    put3(y.topLeftPixel, c.m_topLeftPixel);
    put3(y.pixelDeltaU, c.m_pixelDeltaU);
    put3(y.pixelDeltaV, c.m_pixelDeltaV);
    put3(y.frameX, c.m_cameraFrame.x), put3(y.frameY, c.m_cameraFrame.y), put3(y.frameZ, c.m_cameraFrame.z);
    y.apertureRadius = c.m_apertureRadius;
    y.apertureSides = c.apertureSides;
    y.exposure = c.exposure;
    return y;
  }

private:
  const Scene* m_scene = nullptr;
  std::map<const Mesh*, int32_t> m_meshIndex;
  std::map<const void*, int32_t> m_texIndex;

  bool fail(const char* what) {
    error = what;
    return false;
  }
  static void put3(float* d, const float3& v) { d[0] = v[0], d[1] = v[1], d[2] = v[2]; }
  static void rows(float* d, const float4x4& m) {
    for (int i = 0; i < 12; i++) d[i] = m[size_t(i)];  // row-major storage (mat.hpp:192)
  }
  static void mat3(float* d, const float3x3& m) {
    for (int i = 0; i < 9; i++) d[i] = m[size_t(i)];
  }

  template <typename T, size_t C>
  int32_t texture(const Texture<T, C>* t) {
    if (!t) return -1;
    auto it = m_texIndex.find(t);
    if (it != m_texIndex.end()) return it->second;
    YcTexture y{};
    y.width = t->width(), y.height = t->height(), y.channels = uint32_t(C);
    y.isFloat = std::is_floating_point_v<T> ? 1u : 0u;
    y.type = uint32_t(t->type());
    if constexpr (std::is_floating_point_v<T>) {
      y.offset = texelsF32.size();
      texelsF32.insert(texelsF32.end(), t->data.begin(), t->data.end());
    } else {
      y.offset = texelsU8.size();
      texelsU8.insert(texelsU8.end(), t->data.begin(), t->data.end());
    }
    textures.push_back(y);
    return m_texIndex[t] = int32_t(textures.size()) - 1;
  }

  void addMaterial(const ParametricBSDF& b) {
    YcMaterial y{};
    put3(y.base, b.m_base);
    y.metallic = b.m_cMetallic, y.roughness = b.m_roughness, y.transmission = b.m_cTrans, y.ior = b.m_ior;
    y.anisotropic = b.m_anisotropic, y.clearcoat = b.m_clearcoat, y.clearcoatRoughness = b.m_clearcoatRoughness;
    put3(y.emission, b.m_emission);
    y.normalScale = b.m_normalScale;
    put3(y.volumeColor, b.m_volumeColor);
    y.volumeDensity = b.m_volumeDensity;
    mat3(y.localRotation, b.m_localRotation);
    mat3(y.invRotation, b.m_invRotation);
    y.baseTex = texture(b.m_baseTexture), y.mrTex = texture(b.m_mrTexture), y.transTex = texture(b.m_transmissionTexture);
    y.normalTex = texture(b.m_normalTexture), y.ccTex = texture(b.m_clearcoatTexture), y.emisTex = texture(b.m_emissionTexture);
    y.thinTransmission = b.m_thinTransmission, y.hasAlpha = b.m_hasAlpha, y.hasEmission = b.m_hasEmission;
    materials.push_back(y);
  }

  // Mesh + its SahBVH (bvh.hpp:21-33: children of an inner node are `left`, `left + 1`; leaves are runs of
  // m_indices) → YcBvhNode (both child boxes inlined) + leaf-ordered YcBvhTri
  bool addMesh(const Mesh& m) {
    const size_t nv = m.m_vertices.size(), nf = m.m_triangles.size();
    if (!nv || !nf) return fail("empty mesh");
    YcMesh ym{};
    ym.vertOffset = uint32_t(positions.size() / 3);
    ym.primOffset = uint32_t(primMaterial.size());
    ym.nodeOffset = uint32_t(bvhNodes.size());
    ym.triOffset = uint32_t(bvhTris.size());
    ym.nTris = uint32_t(nf), ym.nVerts = uint32_t(nv);
    for (size_t i = 0; i < nv; i++) {
      const float3& p = m.m_vertices[i];
      const VertexData& v = m.m_vertexData[i];
      positions.insert(positions.end(), {p[0], p[1], p[2]});
      normals.insert(normals.end(), {v.normal[0], v.normal[1], v.normal[2]});
      tangents.insert(tangents.end(), {v.tangent[0], v.tangent[1], v.tangent[2], v.tangent[3]});
      uvs.insert(uvs.end(), {v.texCoords[0], v.texCoords[1]});
    }
    for (size_t i = 0; i < nf; i++) {
      const Triangle& t = m.m_triangles[i];
      primIndices.insert(primIndices.end(), {t.i0, t.i1, t.i2});
      primMaterial.push_back(m.m_materials[i]);
      primLight.push_back(m.m_lights[i]);
    }
    const BVH& bvh = m.m_bvh;
    const std::vector<BVHNode>& rn = bvh.m_nodes;
    // inner nodes reachable from the root, numbered in array order
    std::vector<uint32_t> rank(rn.size(), 0xffffffffu), todo{uint32_t(bvh.m_rootIdx)};
    std::vector<uint32_t> inner;
    while (!todo.empty()) {
      const uint32_t i = todo.back();
      todo.pop_back();
      if (rn[i].span != 0) continue;
      inner.push_back(i);
      todo.push_back(rn[i].left), todo.push_back(rn[i].left + 1);
    }
    std::sort(inner.begin(), inner.end());
    for (size_t k = 0; k < inner.size(); k++) rank[inner[k]] = uint32_t(k);
    ym.nInner = uint32_t(inner.size());
    auto refOf = [&](uint32_t node) { return rn[node].span == 0 ? rank[node] : (YC_REF_LEAF | rn[node].first); };
    const BVHNode& root = rn[bvh.m_rootIdx];
    put3(ym.rootMin, root.bounds.min), put3(ym.rootMax, root.bounds.max);
    ym.rootRef = refOf(uint32_t(bvh.m_rootIdx));
    for (uint32_t i : inner) {
      YcBvhNode o{};
      const BVHNode &c0 = rn[rn[i].left], &c1 = rn[rn[i].left + 1];
      put3(o.c0min, c0.bounds.min), put3(o.c0max, c0.bounds.max);
      put3(o.c1min, c1.bounds.min), put3(o.c1max, c1.bounds.max);
      o.ref0 = refOf(rn[i].left), o.ref1 = refOf(rn[i].left + 1);
      bvhNodes.push_back(o);
    }
    const size_t tbase = bvhTris.size();
    for (size_t i = 0; i < nf; i++) {
      const uint32_t prim = uint32_t(bvh.m_indices[i]);
      const Triangle& tri = m.m_triangles[prim];
      YcBvhTri t{};
      put3(t.p0, m.m_vertices[tri.i0]), put3(t.p1, m.m_vertices[tri.i1]), put3(t.p2, m.m_vertices[tri.i2]);
      t.prim = prim;
      bvhTris.push_back(t);
    }
    todo.assign(1, uint32_t(bvh.m_rootIdx));
    while (!todo.empty()) {
      const uint32_t i = todo.back();
      todo.pop_back();
      if (rn[i].span != 0) bvhTris[tbase + rn[i].first + rn[i].span - 1].flags |= YC_TRI_LAST;
      else todo.push_back(rn[i].left), todo.push_back(rn[i].left + 1);
    }
    meshes.push_back(ym);
    return true;
  }

  // Node tree → DFS pre-order with skip links (what RayIntegrator::testNode's recursion visits, ray-integrator.cpp:20-54)
  void addNode(const Node& n, int parent, int depth) {
    if (depth >= YC_MAX_NODE_DEPTH) {
      error = "scene graph deeper than YC_MAX_NODE_DEPTH";
      return;
    }
    const int self = int(nodes.size());
    nodes.emplace_back();
    {
      YcNode& y = nodes[size_t(self)];
      rows(y.inv, n.transform.m_inverseTransform);
      rows(y.fwd, n.transform.m_transform);
      mat3(y.nrm, n.transform.m_normalTransform);
      put3(y.bmin, n.boundingBox().min), put3(y.bmax, n.boundingBox().max);
      y.mesh = n.mesh() ? m_meshIndex.at(n.mesh()) : -1;
      y.parent = parent, y.depth = depth;
      static const float kIdentityRows[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
      const bool selfIdentity = memcmp(y.inv, kIdentityRows, sizeof kIdentityRows) == 0 && memcmp(y.fwd, kIdentityRows, sizeof kIdentityRows) == 0;
      y.identityChain = selfIdentity && (parent < 0 || nodes[size_t(parent)].identityChain);
    }
    for (const Node& c : n.children()) addNode(c, self, depth + 1);
    nodes[size_t(self)].skip = int(nodes.size());
  }

  bool addLight(const Light& l) {
    YcLight y{};
    y.hdrTex = -1;
    y.power = l.power();
    if (const auto* a = dynamic_cast<const AreaLight*>(&l)) {
      y.type = YC_LIGHT_AREA, y.twoSided = a->twoSided;
      put3(y.emission, a->m_emission);
      y.area = a->m_area;
      const Mesh& m = *a->m_mesh;
      const Triangle& t = *a->m_tri;
      put3(y.p0, m.vertex(t.i0)), put3(y.p1, m.vertex(t.i1)), put3(y.p2, m.vertex(t.i2));
      put3(y.n0, m.vertexData(t.i0).normal), put3(y.n1, m.vertexData(t.i1).normal), put3(y.n2, m.vertexData(t.i2).normal);
      rows(y.fwd, a->m_transform.m_transform);
      mat3(y.nrm, a->m_transform.m_normalTransform);
    } else if (const auto* e = dynamic_cast<const ImageInfiniteLight*>(&l)) {
      y.type = YC_LIGHT_IMAGE_INFINITE;
      y.sceneRadius = e->m_sceneRadius, y.surfaceArea = e->m_surfaceArea;
      put3(y.Lavg, e->m_Lavg);
      rows(y.envFwd, e->transform.m_transform), rows(y.envInv, e->transform.m_inverseTransform);
      y.hdrTex = texture(e->m_emissionTexture);
      // PiecewiseConstant2D (sampling.hpp:160-196) → func[W*H] cdf[(W+1)*H] rowIntegral[H] mfunc[H] mcdf[H+1] mIntegral[1]
      const auto& dist = e->m_distribution;
      const size_t h = dist.m_conditional.size(), w = h ? dist.m_conditional[0].m_func.size() : 0;
      y.distW = uint32_t(w), y.distH = uint32_t(h);
      y.distOffset = envDist.size();
      for (const auto& c : dist.m_conditional) envDist.insert(envDist.end(), c.m_func.begin(), c.m_func.end());
      for (const auto& c : dist.m_conditional) envDist.insert(envDist.end(), c.m_cdf.begin(), c.m_cdf.end());
      for (const auto& c : dist.m_conditional) envDist.push_back(c.m_integral);
      envDist.insert(envDist.end(), dist.m_marginal.m_func.begin(), dist.m_marginal.m_func.end());
      envDist.insert(envDist.end(), dist.m_marginal.m_cdf.begin(), dist.m_marginal.m_cdf.end());
      envDist.push_back(dist.m_marginal.m_integral);
    } else if (const auto* u = dynamic_cast<const UniformInfiniteLight*>(&l)) {
      y.type = YC_LIGHT_UNIFORM_INFINITE;
      y.sceneRadius = u->m_sceneRadius;
      put3(y.emission, u->m_emission);
      put3(y.Lavg, u->m_emission);
      y.surfaceArea = 4.0f * float(pi);
    } else {
      return fail("unknown light class");
    }
    lights.push_back(y);
    return true;
  }
};

/**
 * Drop-in for cpu::TileRenderer<SobolSampler<FastOwenScrambler>, cpu::MISIntegrator> on one or more B200s.
 */
class WavefrontRenderer : public Renderer {
public:
  uint32_t samples = 64;           // tile-renderer.hpp:27-30, same defaults
  uint32_t firstWaveSamples = 64;
  uint32_t maxWaveSamples = 128;
  uint32_t tileSize = 64;
  const tonemap::Tonemap* tonemapper = nullptr;
  uint32_t maxDepth = 30;          // RayIntegrator::m_maxDepth, ray-integrator.hpp:14
  std::vector<int> devices = {0};  // more than one: the frame's tiles are split across these GPUs (yr_create_multi)
  uint32_t traversal = YC_TRAVERSAL_AUTO;

  WavefrontRenderer(Buffer&& buffer, const Camera& camera) noexcept : Renderer(std::move(buffer), camera) {}
  ~WavefrontRenderer() {
    if (m_r) yr_destroy(m_r);
  }

  void render() override {
    if (!prepare()) return;
    yr_render(m_r);
  }
  void abort() override {
    if (m_r) yr_abort(m_r);
  }
  void wait() override {
    if (m_r) yr_wait(m_r);
  }
  RenderData renderSync() override {
    YrRenderData d{};
    if (prepare()) yr_render_sync(m_r, &d);
    return {m_buffer, size_t(d.samplesTaken), size_t(d.totalSamples), d.totalRays,
            std::chrono::milliseconds(int64_t(d.totalTimeMs))};
  }

  [[nodiscard]] const char* lastError() const { return m_error.empty() ? (m_r ? yr_last_error(m_r) : "") : m_error.c_str(); }
  // The HDR accumulation (TileRenderer keeps it in its private m_hdrBuffer)
  bool readHdr(Buffer& out) { return m_r && yr_read(m_r, const_cast<float*>(out.data({0u, 0u})->data()), nullptr, nullptr) == YC_OK; }

private:
  yr_renderer* m_r = nullptr;
  const Scene* m_flattened = nullptr;
  SceneFlattener m_flat;
  YrSettings m_settings{};
  std::string m_error;

  bool prepare() {
    if (!scene) return false;  // integrator.cpp:6 `if (!scene) return;`
    YrSettings s{};
    s.width = m_buffer.width(), s.height = m_buffer.height();
    s.samples = samples, s.firstWaveSamples = firstWaveSamples, s.maxWaveSamples = maxWaveSamples, s.tileSize = tileSize;
    s.maxDepth = maxDepth;
    for (int k = 0; k < 3; k++) s.background[k] = backgroundColor[size_t(k)];
    s.tonemap = YC_TONEMAP_NONE;
    if (const auto* agx = dynamic_cast<const tonemap::AgX*>(tonemapper)) {
      const float p = agx->look.power[0], sat = agx->look.sat;
      s.tonemap = p == tonemap::AgX::golden.power[0] && sat == tonemap::AgX::golden.sat   ? YC_TONEMAP_AGX_GOLDEN
                  : p == tonemap::AgX::punchy.power[0] && sat == tonemap::AgX::punchy.sat ? YC_TONEMAP_AGX_PUNCHY
                                                                                          : YC_TONEMAP_AGX;
    }
    s.estimator = YC_ESTIMATOR_GMON;  // integrator.cpp:17
    s.device = devices.empty() ? 0 : devices[0];
    s.traversal = traversal;
    const bool sameSettings = m_r && memcmp(&s, &m_settings, sizeof s) == 0 && m_flattened == scene;
    if (!sameSettings) {
      if (m_r) yr_destroy(m_r), m_r = nullptr;
      if (m_flattened != scene) {
        m_flat = SceneFlattener();
        if (!m_flat.flatten(*scene)) {
          m_error = m_flat.error;
          return false;
        }
        m_flattened = scene;
      }
      const YcCamera cam = SceneFlattener::camera(m_camera);
      const int rc = devices.size() > 1
                       ? yr_create_multi_flat(&s, &m_flat.flat, &cam, devices.data(), uint32_t(devices.size()), &m_r)
                       : yr_create_flat(&s, &m_flat.flat, &cam, &m_r);
      if (rc != YC_OK) {
        m_error = "yr_create failed";
        return false;
      }
      m_settings = s;
      yr_set_frame_target(m_r, const_cast<float*>(m_buffer.data({0u, 0u})->data()));  // waves land in m_buffer, like finishTile
      yr_set_wave_callback(m_r, &WavefrontRenderer::waveDone, this);
      yr_set_tile_callback(m_r, &WavefrontRenderer::tileDone, this);
      yr_set_done_callback(m_r, &WavefrontRenderer::renderDone, this);
    } else {
      const YcCamera cam = SceneFlattener::camera(m_camera);  // the camera may have moved between renders
      yr_set_camera(m_r, &cam);
    }
    return true;
  }

  RenderData data(const YrRenderData* d) const {
    return {m_buffer, size_t(d->samplesTaken), size_t(d->totalSamples), d->totalRays, std::chrono::milliseconds(int64_t(d->totalTimeMs))};
  }
  static void waveDone(const YrRenderData* d, const YrWaveData* w, void* self) {
    auto* r = static_cast<WavefrontRenderer*>(self);
    if (r->onRenderWaveComplete)
      r->onRenderWaveComplete.value()(r->data(d), {size_t(w->wave), size_t(w->waveSamples), w->rays, std::chrono::milliseconds(int64_t(w->timeMs))});
  }
  static void tileDone(const YrRenderData* d, const YrTileData* t, void* self) {
    auto* r = static_cast<WavefrontRenderer*>(self);
    if (r->onRenderTileComplete)
      r->onRenderTileComplete.value()(r->data(d), {uint2(t->x, t->y), uint2(t->w, t->h), size_t(t->index), size_t(t->total), t->rays,
                                                   std::chrono::milliseconds(int64_t(t->timeMs))});
  }
  static void renderDone(const YrRenderData* d, int aborted, void* self) {
    auto* r = static_cast<WavefrontRenderer*>(self);
    auto& cb = aborted ? r->onRenderAborted : r->onRenderComplete;
    if (cb) cb.value()(r->data(d));
  }
};

}  // namespace yart::cuda
