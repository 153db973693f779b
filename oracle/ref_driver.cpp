// ref_driver.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Command-line harness around the UNMODIFIED reference sources (teofum/yart, compiled
// where they lie under $YART_REF by oracle/Makefile into oracle/_ref/oracle_ref).
// It builds reference `yart::Scene` objects through the reference's public C++ API from
// a neutral ".ysc" scene description and exposes the reference's own implementation of the
// hot path as: whole renders (TileRenderer<SobolSampler<FastOwenScrambler>, MISIntegrator>),
// ray-level traces (RayIntegrator::testNode via a deriving harness class, because testNode
// is protected: src/cpu/ray-integrator.hpp:27-32), BVH dumps, data-table dumps and
// function-level known-answer vectors.  All arithmetic on the path is the reference's;
// this file only moves bytes in and out.
#include <cstdio>
#include <cstdlib>
#include <map>

#include <core/core.hpp>
#include <bsdf/parametric.hpp>
#include <bsdf/luts.hpp>
#include <cpu/mis-integrator.hpp>
#include <cpu/naive-integrator.hpp>
#include <cpu/tile-renderer.hpp>
#include <output/ppm.hpp>
#include <fstream>

#include "../yart_b200/host/scene_desc.hpp"

using namespace yart;
using namespace yart::math;

// ---------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------
static std::vector<uint8_t> readFile(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) {
    fprintf(stderr, "cannot open %s\n", path);
    exit(2);
  }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> d(n);
  if (n && fread(d.data(), 1, n, f) != size_t(n)) exit(2);
  fclose(f);
  return d;
}

struct Writer {
  FILE* f;
  explicit Writer(const char* path) : f(fopen(path, "wb")) {
    if (!f) {
      fprintf(stderr, "cannot write %s\n", path);
      exit(2);
    }
  }
  ~Writer() { fclose(f); }
  template <typename T>
  void put(const T& v) { fwrite(&v, sizeof(T), 1, f); }
  template <typename T>
  void putn(const T* v, size_t n) { fwrite(v, sizeof(T), n, f); }
  void f3(const float3& v) { float a[3] = {v[0], v[1], v[2]}; putn(a, 3); }
};

struct Args {
  std::map<std::string, std::string> kv;
  Args(int argc, char** argv, int first) {
    for (int i = first; i < argc; i++) {
      std::string a = argv[i];
      auto eq = a.find('=');
      if (eq == std::string::npos) kv[a] = "1";
      else kv[a.substr(0, eq)] = a.substr(eq + 1);
    }
  }
  bool has(const char* k) const { return kv.count(k) > 0; }
  double num(const char* k, double d) const { return has(k) ? atof(kv.at(k).c_str()) : d; }
  std::string str(const char* k, const char* d) const { return has(k) ? kv.at(k) : d; }
  float3 vec3(const char* k, float3 d) const {
    if (!has(k)) return d;
    float a = 0, b = 0, c = 0;
    sscanf(kv.at(k).c_str(), "%f,%f,%f", &a, &b, &c);
    return {a, b, c};
  }
};

static float4x4 mat16(const float* m) {
  return float4x4(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12],
                  m[13], m[14], m[15]);
}

// ---------------------------------------------------------------------------------------
// .ysc → reference Scene, through the reference's public API
// ---------------------------------------------------------------------------------------
struct RefScene {
  std::unique_ptr<Scene> scene;
  std::vector<std::unique_ptr<HDRTexture>> hdr;  // env maps are not owned by Scene (main.cpp:81-84)
  std::vector<Mesh*> meshes;
  std::vector<const BSDF*> materials;
  double buildMs = 0;
};

static Node buildNode(const ysc::SceneDesc& d, int idx, const std::vector<Mesh*>& meshes) {
  const auto& nd = d.nodes[idx];
  Node node = nd.mesh >= 0 ? Node(meshes[nd.mesh]) : Node();
  if (nd.hasTransform) node.transform = Transform(mat16(nd.m));
  for (size_t c = 0; c < d.nodes.size(); c++)
    if (d.nodes[c].parent == idx) node.appendChild(buildNode(d, int(c), meshes));
  return node;
}

static RefScene buildScene(const ysc::SceneDesc& d, bool quiet = true) {
  RefScene rs;
  // textures: keep typed pointers by index
  std::vector<std::unique_ptr<MonoTexture>> t1(d.textures.size());
  std::vector<std::unique_ptr<SDRTexture<2>>> t2(d.textures.size());
  std::vector<std::unique_ptr<RGBTexture>> t3(d.textures.size());
  std::vector<std::unique_ptr<RGBATexture>> t4(d.textures.size());
  std::vector<HDRTexture*> thdr(d.textures.size(), nullptr);
  for (size_t i = 0; i < d.textures.size(); i++) {
    const auto& t = d.textures[i];
    auto type = TextureType(t.type);
    if (t.isFloat) {
      auto h = std::make_unique<HDRTexture>(t.width, t.height, type);
      h->data = t.f32;
      thdr[i] = h.get();
      rs.hdr.push_back(std::move(h));
    } else if (t.channels == 1) {
      t1[i] = std::make_unique<MonoTexture>(t.width, t.height, type);
      t1[i]->data = t.u8;
    } else if (t.channels == 2) {
      t2[i] = std::make_unique<SDRTexture<2>>(t.width, t.height, type);
      t2[i]->data = t.u8;
    } else if (t.channels == 3) {
      t3[i] = std::make_unique<RGBTexture>(t.width, t.height, type);
      t3[i]->data = t.u8;
    } else {
      t4[i] = std::make_unique<RGBATexture>(t.width, t.height, type);
      t4[i]->data = t.u8;
    }
  }

  std::streambuf* oldBuf = nullptr;
  std::ostringstream sink;
  if (quiet) oldBuf = std::cout.rdbuf(sink.rdbuf());  // BVH::printStats talks to stdout

  auto t0 = std::chrono::high_resolution_clock::now();
  std::vector<std::unique_ptr<Mesh>> meshOwners;
  for (const auto& m : d.meshes) {
    std::vector<float3> verts(m.nVerts());
    std::vector<VertexData> vdata(m.nVerts());
    std::vector<Face> faces(m.nFaces());
    for (size_t i = 0; i < verts.size(); i++) {
      verts[i] = float3(m.positions[3 * i], m.positions[3 * i + 1], m.positions[3 * i + 2]);
      const float* v = &m.vertexData[9 * i];
      vdata[i].normal = float3(v[0], v[1], v[2]);
      vdata[i].tangent = float4(v[3], v[4], v[5], v[6]);
      vdata[i].texCoords = float2(v[7], v[8]);
    }
    for (size_t i = 0; i < faces.size(); i++)
      faces[i] = {m.faces[4 * i], m.faces[4 * i + 1], m.faces[4 * i + 2], m.faces[4 * i + 3]};
    meshOwners.push_back(std::make_unique<Mesh>(Mesh(verts, vdata, faces)));
    rs.meshes.push_back(meshOwners.back().get());
  }
  auto t1c = std::chrono::high_resolution_clock::now();
  rs.buildMs = std::chrono::duration<double, std::milli>(t1c - t0).count();
  if (quiet) std::cout.rdbuf(oldBuf);

  for (size_t i = 0; i < d.meshes.size(); i++)
    for (size_t f = 0; f < d.meshes[i].nFaces(); f++) rs.meshes[i]->lightIdx(uint32_t(f)) = d.meshes[i].lightIdx[f];

  Node root = buildNode(d, 0, rs.meshes);
  rs.scene = std::make_unique<Scene>(std::move(root));
  for (auto& m : meshOwners) rs.scene->addMesh(std::move(m));

  for (const auto& m : d.materials) {
    auto tex = [&](int i, auto& pool) -> decltype(pool[0].get()) { return i >= 0 ? pool[i].get() : nullptr; };
    auto bsdf = std::make_unique<ParametricBSDF>(
      float3(m.base[0], m.base[1], m.base[2]), tex(m.baseTex, t4), tex(m.mrTex, t2), tex(m.transTex, t1),
      tex(m.normalTex, t3), tex(m.ccTex, t1), tex(m.emisTex, t3), m.metallic, m.roughness, m.transmission, m.ior,
      m.anisotropic, m.anisoRotation, m.clearcoat, m.clearcoatRoughness,
      float3(m.emission[0], m.emission[1], m.emission[2]), m.normalScale, m.thinTransmission != 0,
      float3(m.volumeColor[0], m.volumeColor[1], m.volumeColor[2]), m.volumeDensity);
    rs.materials.push_back(bsdf.get());
    rs.scene->addMaterial(std::unique_ptr<BSDF>(std::move(bsdf)));
  }
  // hand texture ownership to the scene (keeps them alive like gltf.cpp does)
  for (auto& t : t1) if (t) rs.scene->addTexture(std::move(t));
  for (auto& t : t2) if (t) rs.scene->addTexture(std::move(t));
  for (auto& t : t3) if (t) rs.scene->addTexture(std::move(t));
  for (auto& t : t4) if (t) rs.scene->addTexture(std::move(t));

  for (const auto& l : d.lights) {
    Transform xf = l.hasTransform ? Transform(mat16(l.m)) : Transform();
    if (l.type == ysc::AreaLightT) {
      Mesh* mesh = rs.meshes[l.mesh];
      AreaLight light(&mesh->triangles()[l.tri], mesh, float3(l.emission[0], l.emission[1], l.emission[2]), xf);
      light.twoSided = l.twoSided != 0;
      rs.scene->addLight(std::move(light));
    } else if (l.type == ysc::ImageInfiniteT) {
      ImageInfiniteLight light(l.sceneRadius, thdr[l.hdrTex]);
      light.transform = xf;
      rs.scene->addLight(std::move(light));
    } else {
      UniformInfiniteLight light(l.sceneRadius, float3(l.emission[0], l.emission[1], l.emission[2]));
      rs.scene->addLight(std::move(light));
    }
  }
  return rs;
}

static Camera makeCamera(const Args& a, uint32_t w, uint32_t h) {
  Camera cam({w, h}, float(a.num("focal", 35)), float(a.num("fnum", 0)));
  cam.exposure = float(a.num("exposure", 0));
  cam.apertureSides = uint32_t(a.num("sides", 0));
  float3 pos = a.vec3("pos", {0, 0, 5}), target = a.vec3("target", {0, 0, 0}), up = a.vec3("up", {0, 0, 0});
  cam.moveAndLookAt(pos, target, up);
  return cam;
}

// ---------------------------------------------------------------------------------------
// integrator subclasses (public m_maxDepth: src/cpu/ray-integrator.hpp:14)
// ---------------------------------------------------------------------------------------
static uint32_t g_maxDepth = 30;
using RefSampler = SobolSampler<FastOwenScrambler>;

struct DepthIntegrator : cpu::MISIntegrator {
  DepthIntegrator(Buffer& b, const Camera& c, Sampler& s) noexcept : cpu::MISIntegrator(b, c, s) {
    m_maxDepth = g_maxDepth;
  }
};

struct DepthNaiveIntegrator : cpu::NaiveIntegrator {  // src/cpu/naive-integrator.cpp
  DepthNaiveIntegrator(Buffer& b, const Camera& c, Sampler& s) noexcept : cpu::NaiveIntegrator(b, c, s) {
    m_maxDepth = g_maxDepth;
  }
};

struct TraceHarness : cpu::MISIntegrator {
  TraceHarness(Buffer& b, const Camera& c, Sampler& s) noexcept : cpu::MISIntegrator(b, c, s) {}
  bool closest(const Ray& r, float tMin, cpu::Hit& hit) { return testNode(r, tMin, hit, scene->root()); }
};

// ---------------------------------------------------------------------------------------
// commands
// ---------------------------------------------------------------------------------------
template <class IntegratorT, class SamplerT>
static int cmdRenderWith(int argc, char** argv);

// integrator=mis (default) | naive: the `Integrator` template argument of TileRenderer (src/main.cpp:17)
// scrambler=fastowen (default) | owen | binary: the R of `Sampler = SobolSampler<R>` (src/main.cpp:16)
static int cmdRender(int argc, char** argv) {
  if (argc < 4) return 1;
  Args a(argc, argv, 4);
  const std::string scr = a.str("scrambler", "fastowen");
  if (a.str("integrator", "mis") == "naive") return cmdRenderWith<DepthNaiveIntegrator, RefSampler>(argc, argv);
  const std::string smp = a.str("sampler", "sobol");  // the `Sampler` template argument (src/main.cpp:16)
  if (smp == "naive") return cmdRenderWith<DepthIntegrator, NaiveSampler>(argc, argv);
  if (smp == "stratified") return cmdRenderWith<DepthIntegrator, StratifiedSampler>(argc, argv);
  if (scr == "owen") return cmdRenderWith<DepthIntegrator, SobolSampler<OwenScrambler>>(argc, argv);
  if (scr == "binary") return cmdRenderWith<DepthIntegrator, SobolSampler<BinaryPermuteScrambler>>(argc, argv);
  return cmdRenderWith<DepthIntegrator, RefSampler>(argc, argv);
}

template <class IntegratorT, class SamplerT>
static int cmdRenderWith(int argc, char** argv) {
  Args a(argc, argv, 4);
  ysc::SceneDesc d;
  std::string err;
  if (!ysc::load(argv[2], d, &err)) {
    fprintf(stderr, "%s\n", err.c_str());
    return 2;
  }
  RefScene rs = buildScene(d);
  uint32_t w = uint32_t(a.num("w", 64)), h = uint32_t(a.num("h", 64));
  Camera cam = makeCamera(a, w, h);
  g_maxDepth = uint32_t(a.num("maxdepth", 30));

  cpu::TileRenderer<SamplerT, IntegratorT> r(Buffer(w, h), cam);
  r.scene = rs.scene.get();
  r.samples = uint32_t(a.num("spp", 16));
  r.firstWaveSamples = uint32_t(a.num("first", r.samples));
  r.maxWaveSamples = uint32_t(a.num("max", r.samples));
  r.tileSize = uint32_t(a.num("tile", 64));
  if (a.has("threads")) r.threadCount = uint32_t(a.num("threads", 1));
  r.backgroundColor = a.vec3("bg", {0, 0, 0});
  r.tonemapper = nullptr;  // m_buffer then carries the HDR accumulation (tile-renderer.hpp:238)

  // Ray count of a render: the reference's worker threads add their tiles' counts into m_totalRays OUTSIDE its mutex
  // (tile-renderer.hpp:217-218), so RenderData::totalRays can lose an update when two tiles finish together (seen as a
  // rare off-by-one-tile count on a loaded box).  The per-tile counts are exact, and the tile callback runs under
  // m_bufferMutex (tile-renderer.hpp:225, 243-262): their sum is the render's ray count.
  uint64_t tileRaySum = 0;
  r.onRenderTileComplete = [&](Renderer::RenderData, Renderer::TileData t) { tileRaySum += t.rays; };

  // repeat=N: time N back-to-back renderSync() calls of the same frame (bench.py's reference arm);
  // the last one is written out.  Each line of "steps" is one call: rays and wall milliseconds.
  int repeat = int(a.num("repeat", 1));
  std::vector<std::pair<uint64_t, double>> steps;
  auto t0 = std::chrono::high_resolution_clock::now();
  auto res = r.renderSync();
  auto t1 = std::chrono::high_resolution_clock::now();
  double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  const uint64_t reportedRays = res.totalRays;  // what the reference itself reports (kept in the JSON line: "rays_reported")
  res.totalRays = tileRaySum;
  steps.push_back({tileRaySum, ms});
  for (int it = 1; it < repeat; it++) {  // (the sum keeps growing over the calls, like the reference's m_totalRays)
    t0 = std::chrono::high_resolution_clock::now();
    auto again = r.renderSync();
    t1 = std::chrono::high_resolution_clock::now();
    ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    (void)again;
    steps.push_back({tileRaySum, ms});
  }

  tonemap::AgX agx;
  std::string tm = a.str("tonemap", "agx");
  if (tm == "golden") agx.look = tonemap::AgX::golden;
  if (tm == "punchy") agx.look = tonemap::AgX::punchy;

  Writer out(argv[3]);
  out.put(w);
  out.put(h);
  out.put(uint64_t(res.totalRays));
  out.put(ms);
  out.put(uint32_t(r.threadCount));
  out.put(rs.buildMs);
  for (uint32_t y = 0; y < h; y++)
    for (uint32_t x = 0; x < w; x++) {
      float4 p = res.buffer(x, y);
      out.putn(p.data(), 4);
    }
  for (uint32_t y = 0; y < h; y++)
    for (uint32_t x = 0; x < w; x++) {
      float3 hdr = float3(res.buffer(x, y));
      float4 p = tm == "none" ? res.buffer(x, y) : float4(agx(hdr), 1.0f);
      out.putn(p.data(), 4);
    }
  if (a.has("ppm")) {
    // the reference's own writer (src/output/ppm.cpp:6-21) on the tonemapped frame
    Buffer ldr(w, h);
    for (uint32_t y = 0; y < h; y++)
      for (uint32_t x = 0; x < w; x++) {
        float3 hdr = float3(res.buffer(x, y));
        ldr(x, y) = tm == "none" ? res.buffer(x, y) : float4(agx(hdr), 1.0f);
      }
    std::ofstream os(a.str("ppm", "out.ppm"), std::ios::binary);
    output::writePPM(os, ldr);
  }
  printf("{\"rays\": %llu, \"rays_reported\": %llu, \"ms\": %.3f, \"threads\": %u, \"build_ms\": %.3f, \"w\": %u, \"h\": %u, \"spp\": %u, \"steps\": [",
         (unsigned long long) res.totalRays, (unsigned long long) reportedRays, ms, r.threadCount, rs.buildMs, w, h, r.samples);
  for (size_t i = 0; i < steps.size(); i++)
    printf("%s[%llu, %.3f]", i ? ", " : "", (unsigned long long) steps[i].first, steps[i].second);
  printf("]}\n");
  return 0;
}

// rays.bin: u32 n, then n × {o[3], tmin, d[3], tmax}.  mode=closest|any
// hits.bin: n × {f32 t; u32 prim; i32 material; i32 lightIdx; u32 backSide; u32 didHit;
//                f32 p[3], n[3], tg[3], uv[2], att[3]}   (20 × 4 bytes)
static int cmdTrace(int argc, char** argv) {
  if (argc < 5) return 1;
  Args a(argc, argv, 5);
  ysc::SceneDesc d;
  std::string err;
  if (!ysc::load(argv[2], d, &err)) {
    fprintf(stderr, "%s\n", err.c_str());
    return 2;
  }
  RefScene rs = buildScene(d);
  auto raw = readFile(argv[3]);
  uint32_t n = *reinterpret_cast<uint32_t*>(raw.data());
  const float* rays = reinterpret_cast<const float*>(raw.data() + 4);
  bool any = a.str("mode", "closest") == "any";

  Buffer buf(1, 1);
  Camera cam({64u, 64u}, 35.0f);
  RefSampler sampler(16, {64u, 64u});
  TraceHarness h(buf, cam, sampler);
  h.scene = rs.scene.get();

  std::map<const BSDF*, int> matIdx;
  for (size_t i = 0; i < rs.materials.size(); i++) matIdx[rs.materials[i]] = int(i);

  Writer out(argv[4]);
  auto t0 = std::chrono::high_resolution_clock::now();
  for (uint32_t i = 0; i < n; i++) {
    const float* r = rays + 8 * size_t(i);
    Ray ray(float3(r[0], r[1], r[2]), float3(r[4], r[5], r[6]));
    ray.nee = any;
    cpu::Hit hit;
    if (any || a.has("usetmax")) hit.t = r[7];
    sampler.startPixelSample({0u, 0u}, 0);
    bool did = h.closest(ray, r[3], hit);
    out.put(hit.t);
    out.put(uint32_t(did ? hit.idx : 0xffffffffu));
    out.put(int32_t(did && hit.bsdf ? matIdx[hit.bsdf] : -1));
    out.put(int32_t(did ? hit.lightIdx : -1));
    out.put(uint32_t(did ? hit.backSide : 0));
    out.put(uint32_t(did));
    out.f3(did ? hit.p : float3());
    out.f3(did ? hit.n : float3());
    out.f3(did ? hit.tg : float3());
    float uv[2] = {did ? hit.uv[0] : 0, did ? hit.uv[1] : 0};
    out.putn(uv, 2);
    out.f3(hit.attenuation);
  }
  auto t1 = std::chrono::high_resolution_clock::now();
  printf("{\"rays\": %u, \"ms\": %.3f, \"build_ms\": %.3f}\n", n,
         std::chrono::duration<double, std::milli>(t1 - t0).count(), rs.buildMs);
  return 0;
}

// bvh.bin: per mesh: u32 nNodes(reachable, numbered as in the reference array), u32 nTris,
//          nodes × {f32 min[3], max[3]; u32 leftFirst; u32 span}, indices u32[nTris]
static int cmdBvh(int argc, char** argv) {
  if (argc < 4) return 1;
  ysc::SceneDesc d;
  std::string err;
  if (!ysc::load(argv[2], d, &err)) {
    fprintf(stderr, "%s\n", err.c_str());
    return 2;
  }
  RefScene rs = buildScene(d);
  Writer out(argv[3]);
  out.put(uint32_t(rs.meshes.size()));
  // kind=median: MedianSplitBVH (bvh.hpp:237-264) built over each mesh's data the way Mesh's constructor builds
  // its BVHType (mesh.hpp:54-61: m_bvh.init(&m_vertices, &m_triangles, &m_centroids)); Mesh does not expose
  // m_centroids, so they are recomputed with createTriangle's expression (primitives.hpp:46).
  Args a(argc, argv, 4);
  const bool median = a.str("kind", "sah") == "median";
  for (Mesh* m : rs.meshes) {
    MedianSplitBVH medianBvh;
    std::vector<float3> centroids;
    if (median) {
      for (const Triangle& t : m->triangles())
        centroids.push_back((m->vertices()[t.i0] + m->vertices()[t.i1] + m->vertices()[t.i2]) / 3.0f);
      medianBvh.init(&m->vertices(), &m->triangles(), &centroids);
    }
    const BVH& bvh = median ? static_cast<const BVH&>(medianBvh) : m->bvh();
    // nodes are allocated contiguously in creation order; highest reachable index + 1 = nodes used
    uint32_t maxIdx = 0;
    std::vector<uint32_t> stack{0};
    while (!stack.empty()) {
      uint32_t i = stack.back();
      stack.pop_back();
      maxIdx = std::max(maxIdx, i);
      if (bvh[i].span == 0) {
        stack.push_back(bvh[i].left);
        stack.push_back(bvh[i].left + 1);
      }
    }
    uint32_t nNodes = maxIdx + 1, nTris = uint32_t(m->triangles().size());
    out.put(nNodes);
    out.put(nTris);
    for (uint32_t i = 0; i < nNodes; i++) {
      const BVHNode& nd = bvh[i];
      out.f3(nd.bounds.min);
      out.f3(nd.bounds.max);
      out.put(uint32_t(nd.left));
      out.put(uint32_t(nd.span));
    }
    for (uint32_t i = 0; i < nTris; i++) out.put(uint32_t(bvh.idx(i)));
  }
  printf("{\"build_ms\": %.3f}\n", rs.buildMs);
  return 0;
}

// tables.bin: the data tables the path reads (values only), in this order:
//   ggx_E[32*32] ggx_Eavg[32] ggx_base_E[16^3] ggx_base_Eavg[16^2]
//   glass_E[16^3] glass_Eavg[16^2] glass_inv_E[16^3] glass_inv_Eavg[16^2]   (f32)
//   sobol matrices dims 0..1 (2*52 u32)
static int cmdTables(int argc, char** argv) {
  if (argc < 3) return 1;
  Writer out(argv[2]);
  out.putn(&lut::table_ggx_E[0][0], 32 * 32);
  out.putn(&lut::table_ggx_Eavg[0], 32);
  out.putn(&lut::table_ggx_base_E[0][0][0], 16 * 16 * 16);
  out.putn(&lut::table_ggx_base_Eavg[0][0], 16 * 16);
  out.putn(&lut::table_ggx_glass_E[0][0][0], 16 * 16 * 16);
  out.putn(&lut::table_ggx_glass_Eavg[0][0], 16 * 16);
  out.putn(&lut::table_ggx_glass_inv_E[0][0][0], 16 * 16 * 16);
  out.putn(&lut::table_ggx_glass_inv_Eavg[0][0], 16 * 16);
  out.putn(&sobol::matrices[0], 2 * sobol::sobolMatrixSize);
  return 0;
}

// ---------------------------------------------------------------------------------------
// function-level known-answer vectors
// ---------------------------------------------------------------------------------------
static int cmdKat(int argc, char** argv) {
  if (argc < 5) return 1;
  std::string kind = argv[2];
  auto raw = readFile(argv[3]);
  const uint8_t* p = raw.data();
  auto u32 = [&]() { uint32_t v; memcpy(&v, p, 4); p += 4; return v; };
  auto f32 = [&]() { float v; memcpy(&v, p, 4); p += 4; return v; };
  auto v3 = [&]() { float a = f32(), b = f32(), c = f32(); return float3(a, b, c); };
  auto v2 = [&]() { float a = f32(), b = f32(); return float2(a, b); };
  Writer out(argv[4]);

  if (kind == "sampler") {
    // in: u32 spp, u32 n, n×{u32 x,y,s}.  out: n×8 floats: get2D get2D get1D get1D get2D
    uint32_t spp = u32(), n = u32();
    RefSampler s(spp, {64u, 64u});
    for (uint32_t i = 0; i < n; i++) {
      uint32_t x = u32(), y = u32(), smp = u32();
      s.startPixelSample({x, y}, smp);
      float2 a = s.get2D(), b = s.get2D();
      float c = s.get1D(), d = s.get1D();
      float2 e = s.get2D();
      float o[8] = {a[0], a[1], b[0], b[1], c, d, e[0], e[1]};
      out.putn(o, 8);
    }
  } else if (kind == "lut") {
    // in: u32 n, n×{a,b,c,ior}.  out: n×8 floats
    uint32_t n = u32();
    for (uint32_t i = 0; i < n; i++) {
      float a = f32(), b = f32(), c = f32(), ior = f32();
      float o[8] = {lut::ggxE(a, b), lut::ggxEavg(b), lut::ggxBaseE(a, b, c), lut::ggxBaseEavg(a, b),
                    lut::ggxGlassE(ior, b, c), lut::ggxGlassEavg(ior, b), fresnelDielectric(a * 2.0f - 1.0f, ior),
                    roughen(b)};
      out.putn(o, 8);
    }
  } else if (kind == "ggx") {
    // in: u32 n, n×{r, aniso, w[3], wm[3], u[2]}. out: n×{mdf(wm), g1(w), g(w,wm), vmdf(w,wm), smooth, svm[3]}
    uint32_t n = u32();
    for (uint32_t i = 0; i < n; i++) {
      float r = f32(), an = f32();
      float3 w = v3(), wm = v3();
      float2 u = v2();
      GGX g(r, an);
      float3 s = g.sampleVisibleMicrofacet(w, u);
      float o[8] = {g.mdf(wm), g.g1(w), g.g(w, wm), g.vmdf(w, wm), g.smooth() ? 1.0f : 0.0f, s[0], s[1], s[2]};
      out.putn(o, 8);
    }
  } else if (kind == "bsdf") {
    // argv[5] = scene.ysc (materials + textures)
    // in: u32 n, n×{u32 mat; wo[3] wi[3] n[3] t[3] uv[2] u[2] uc uc2; u32 regularized; f32 tgw; f32 dist}
    // out: n×{f[3] pdf | s.scatter(i32) s.f[3] s.Le[3] s.wi[3] s.pdf s.rough | alpha base[3] transparent(i32)
    //         normal[3] atten[3]}  = 4 + 12 + 5 + 6 = 27 words
    if (argc < 6) return 1;
    ysc::SceneDesc d;
    if (!ysc::load(argv[5], d)) return 2;
    RefScene rs = buildScene(d);
    uint32_t n = u32();
    for (uint32_t i = 0; i < n; i++) {
      uint32_t mat = u32();
      float3 wo = v3(), wi = v3(), nn = v3(), t = v3();
      float2 uv = v2(), u = v2();
      float uc = f32(), uc2 = f32();
      uint32_t reg = u32();
      float tgw = f32(), dist = f32();
      const BSDF* b = rs.materials[mat];
      float3 f = b->f(wo, wi, nn, t, uv);
      float pdf = b->pdf(wo, wi, nn, t, uv);
      BSDFSample s = b->sample(wo, nn, t, uv, u, uc, uc2, reg != 0);
      out.f3(f);
      out.put(pdf);
      out.put(int32_t(s.scatter));
      out.f3(s.f);
      out.f3(s.Le);
      out.f3(s.wi);
      out.put(s.pdf);
      out.put(s.roughness);
      out.put(b->alpha(uv));
      out.f3(b->base(uv));
      out.put(int32_t(b->transparent()));
      out.f3(b->normal(nn, float4(t, tgw), uv));
      out.f3(b->attenuation(dist));
    }
  } else if (kind == "light") {
    // argv[5] = scene.ysc.  in: u32 n, n×{u32 light; p[3] n[3] u[2] wi[3] uc}
    // out: n×{Li[3] wi[3] p[3] n[3] pdf | pdf(wi) power type(i32) Le[3] | picked(i32) pPick pOf(light)} = 13+6+3 = 22
    if (argc < 6) return 1;
    ysc::SceneDesc d;
    if (!ysc::load(argv[5], d)) return 2;
    RefScene rs = buildScene(d);
    PowerLightSampler ls;
    ls.init(rs.scene.get());
    std::map<const Light*, int> lidx;
    for (size_t i = 0; i < rs.scene->nLights(); i++) lidx[&rs.scene->light(i)] = int(i);
    uint32_t n = u32();
    for (uint32_t i = 0; i < n; i++) {
      uint32_t li = u32();
      float3 pp = v3(), nn = v3();
      float2 u = v2();
      float3 wi = v3();
      float uc = f32();
      const Light& L = rs.scene->light(li);
      LightSample s = L.sample(pp, nn, u, 0.0f);
      out.f3(s.Li);
      out.f3(s.wi);
      out.f3(s.p);
      out.f3(s.n);
      out.put(s.pdf);
      out.put(L.pdf(wi));
      out.put(L.power());
      out.put(int32_t(L.type() == Light::Type::Area ? 0 : 1));
      out.f3(L.Le(octahedralUV(wi)));
      SampledLight sl = ls.sample(pp, nn, uc);
      out.put(int32_t(lidx[&sl.light]));
      out.put(sl.p);
      out.put(ls.p(pp, nn, li));
    }
  } else if (kind == "gmon") {
    // in: u32 n (samples per pixel), u32 npix, float3[npix*n].  out: npix×{gmon[3] mon[3] mean[3]}
    uint32_t n = u32(), npix = u32();
    for (uint32_t i = 0; i < npix; i++) {
      GMoNEstimator g(int32_t(n), 15);
      MoNEstimator m(int32_t(n), 15);
      MeanEstimator mean(n);
      for (uint32_t s = 0; s < n; s++) {
        float3 v = v3();
        g.addSample(v);
        m.addSample(v);
        mean.addSample(v);
      }
      out.f3(g.getValue());
      out.f3(m.getValue());
      out.f3(mean.getValue());
    }
  } else if (kind == "lightuniform") {
    // argv[5] = scene.ysc.  in: as "light".  out: n×{picked(i32) pPick pOf(light)}  (UniformLightSampler, light-sampler.cpp:11-31)
    if (argc < 6) return 1;
    ysc::SceneDesc d;
    if (!ysc::load(argv[5], d)) return 2;
    RefScene rs = buildScene(d);
    UniformLightSampler ls;
    ls.init(rs.scene.get());
    std::map<const Light*, int> lidx;
    for (size_t i = 0; i < rs.scene->nLights(); i++) lidx[&rs.scene->light(i)] = int(i);
    uint32_t n = u32();
    for (uint32_t i = 0; i < n; i++) {
      uint32_t li = u32();
      float3 pp = v3(), nn = v3();
      v2();
      v3();
      float uc = f32();
      SampledLight sl = ls.sample(pp, nn, uc);
      out.put(int32_t(lidx[&sl.light]));
      out.put(sl.p);
      out.put(ls.p(pp, nn, li));
    }
  } else if (kind == "gmonb") {
    // in: as "gmon".  out: npix×gmonb[3]  (GMoNbEstimator, src/core/estimator.hpp:94-141)
    uint32_t n = u32(), npix = u32();
    for (uint32_t i = 0; i < npix; i++) {
      GMoNbEstimator g(int32_t(n), 15);
      for (uint32_t s = 0; s < n; s++) g.addSample(v3());
      out.f3(g.getValue());
    }
  } else if (kind == "agx") {
    // in: u32 look, u32 n, float3[n]. out float3[n]
    uint32_t look = u32(), n = u32();
    tonemap::AgX agx;
    if (look == 1) agx.look = tonemap::AgX::golden;
    if (look == 2) agx.look = tonemap::AgX::punchy;
    for (uint32_t i = 0; i < n; i++) out.f3(agx(v3()));
  } else if (kind == "camera") {
    // in: u32 w,h; focal fnum; u32 sides; pos[3] target[3] up[3]; u32 n; n×{u32 px,py; film[2] lens[2]}
    // out: n×{o[3] d[3]}
    uint32_t w = u32(), h = u32();
    float focal = f32(), fnum = f32();
    uint32_t sides = u32();
    float3 pos = v3(), target = v3(), up = v3();
    Camera cam({w, h}, focal, fnum);
    cam.apertureSides = sides;
    cam.moveAndLookAt(pos, target, up);
    uint32_t n = u32();
    for (uint32_t i = 0; i < n; i++) {
      uint32_t px = u32(), py = u32();
      float2 film = v2(), lens = v2();
      Ray r = cam.getRay({px, py}, film, lens);
      out.f3(r.origin);
      out.f3(r.dir);
    }
  } else if (kind == "xform") {
    // in: u32 n, n×{m[16], v[3]}.  out: n×{inv[16] (via Transform::inverse of basis), point[3] vector[3] normal[3],
    //                                      invpoint[3] invvector[3]}
    uint32_t n = u32();
    for (uint32_t i = 0; i < n; i++) {
      float m[16];
      for (float& x : m) x = f32();
      float3 v = v3();
      Transform t(mat16(m));
      float4 cols[4] = {t.inverse(float4(1, 0, 0, 0)), t.inverse(float4(0, 1, 0, 0)), t.inverse(float4(0, 0, 1, 0)),
                        t.inverse(float4(0, 0, 0, 1))};
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) out.put(cols[c][r]);
      out.f3(t(v, Transform::Type::Point));
      out.f3(t(v, Transform::Type::Vector));
      out.f3(t(v, Transform::Type::Normal));
      out.f3(t.inverse(v, Transform::Type::Point));
      out.f3(t.inverse(v, Transform::Type::Vector));
    }
  } else if (kind == "texture") {
    // argv[5] = scene.ysc. in: u32 n, n×{u32 tex; uv[2]}. out: n×4 floats (unused channels 0)
    if (argc < 6) return 1;
    ysc::SceneDesc d;
    if (!ysc::load(argv[5], d)) return 2;
    uint32_t n = u32();
    for (uint32_t i = 0; i < n; i++) {
      uint32_t ti = u32();
      float2 uv = v2();
      const auto& t = d.textures[ti];
      float o[4] = {0, 0, 0, 0};
      auto type = TextureType(t.type);
      if (t.isFloat) {
        HDRTexture x(t.width, t.height, type);
        x.data = t.f32;
        float3 v = x.sample(uv);
        o[0] = v[0], o[1] = v[1], o[2] = v[2];
      } else if (t.channels == 1) {
        MonoTexture x(t.width, t.height, type);
        x.data = t.u8;
        o[0] = x.sample(uv);
      } else if (t.channels == 2) {
        SDRTexture<2> x(t.width, t.height, type);
        x.data = t.u8;
        float2 v = x.sample(uv);
        o[0] = v[0], o[1] = v[1];
      } else if (t.channels == 3) {
        RGBTexture x(t.width, t.height, type);
        x.data = t.u8;
        float3 v = x.sample(uv);
        o[0] = v[0], o[1] = v[1], o[2] = v[2];
      } else {
        RGBATexture x(t.width, t.height, type);
        x.data = t.u8;
        float4 v = x.sample(uv);
        o[0] = v[0], o[1] = v[1], o[2] = v[2], o[3] = v[3];
      }
      out.putn(o, 4);
    }
  } else {
    fprintf(stderr, "unknown kat kind %s\n", kind.c_str());
    return 1;
  }
  return 0;
}

// texload in.png type(0 linear,1 sRGB,2 noncolor) C ch0,ch1,... out.bin
// out: u32 w, h, C then w*h*C bytes — the reference's loadTexture<C> (core/texture.hpp:62-90, stb_image decode)
template <size_t C>
static int texloadImpl(const std::vector<uint8_t>& png, int type, const std::vector<uint32_t>& ch, const char* outPath) {
  std::array<uint32_t, C> channels{};
  for (size_t i = 0; i < C; i++) channels[i] = ch[i];
  SDRTexture<C> t = loadTexture<C>(png.data(), int32_t(png.size()), TextureType(type), channels);
  Writer out(outPath);
  out.put(uint32_t(t.width()));
  out.put(uint32_t(t.height()));
  out.put(uint32_t(C));
  out.putn(t.data.data(), t.data.size());
  return 0;
}
static int cmdTexload(int argc, char** argv) {
  if (argc < 7) return 1;
  auto png = readFile(argv[2]);
  int type = atoi(argv[3]), C = atoi(argv[4]);
  std::vector<uint32_t> ch;
  for (const char* p = argv[5]; *p;) {
    ch.push_back(uint32_t(strtoul(p, const_cast<char**>(&p), 10)));
    if (*p == ',') p++;
  }
  if (int(ch.size()) < C) return 1;
  switch (C) {
    case 1: return texloadImpl<1>(png, type, ch, argv[6]);
    case 2: return texloadImpl<2>(png, type, ch, argv[6]);
    case 3: return texloadImpl<3>(png, type, ch, argv[6]);
    case 4: return texloadImpl<4>(png, type, ch, argv[6]);
  }
  return 1;
}

// hdrload in.hdr out.bin: loadTextureHDR (src/core/texture.cpp:21-35) → u32 w, h; f32 rgb[w*h*3]
static int cmdHdrload(int argc, char** argv) {
  if (argc < 4) return 1;
  HDRTexture t = loadTextureHDR(argv[2]);
  Writer out(argv[3]);
  out.put(uint32_t(t.width()));
  out.put(uint32_t(t.height()));
  out.putn(t.data.data(), t.data.size());
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr,
            "usage: oracle_ref render scene.ysc out.bin [k=v...]\n"
            "       oracle_ref trace scene.ysc rays.bin hits.bin [mode=closest|any]\n"
            "       oracle_ref bvh scene.ysc out.bin\n"
            "       oracle_ref tables out.bin\n"
            "       oracle_ref kat <kind> in.bin out.bin [scene.ysc]\n"
            "       oracle_ref texload in.png type C ch0,ch1,.. out.bin\n"
            "       oracle_ref hdrload in.hdr out.bin\n");
    return 1;
  }
  std::string cmd = argv[1];
  if (cmd == "render") return cmdRender(argc, argv);
  if (cmd == "trace") return cmdTrace(argc, argv);
  if (cmd == "bvh") return cmdBvh(argc, argv);
  if (cmd == "tables") return cmdTables(argc, argv);
  if (cmd == "kat") return cmdKat(argc, argv);
  if (cmd == "texload") return cmdTexload(argc, argv);
  if (cmd == "hdrload") return cmdHdrload(argc, argv);
  fprintf(stderr, "unknown command %s\n", cmd.c_str());
  return 1;
}
