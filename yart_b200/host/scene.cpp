// scene.cpp — SceneDesc → HostScene (derived constants + flat GPU layout).  See scene.hpp.
#include "scene.hpp"

#include <atomic>
#include <chrono>
#include <cmath>
#include <limits>
#include <memory>
#include <thread>
#include <cstring>
#include <functional>

extern "C" {
extern const unsigned char yb_tables_start[];
extern const unsigned char yb_tables_end[];
}

namespace yartb {

static const float kPi = float(M_PI);  // float(pi), math_base.hpp:12

static void put3(float* d, f3 v) { d[0] = v.x, d[1] = v.y, d[2] = v.z; }
static void putRows(float* d, const Mat4& m) { memcpy(d, m.m, 12 * sizeof(float)); }

static Mat4 mat16(const float* m) {
  Mat4 r;
  memcpy(r.m, m, sizeof(r.m));
  return r;
}

bool HostScene::loadLuts(std::string* err) {
  size_t n = size_t(yb_tables_end - yb_tables_start);
  if (n < 14112 * 4) {
    if (err) *err = "embedded LUT tables missing";
    return false;
  }
  lutTables.resize(14112);
  memcpy(lutTables.data(), yb_tables_start, 14112 * 4);
  return true;
}

// PiecewiseConstant1D ctor, math/sampling.hpp:122-144.  Appends func[n], cdf[n+1]; returns integral.
static float buildDistribution1D(const float* f, size_t n, float mn, float mx, float* func, float* cdf) {
  for (size_t i = 0; i < n; i++) func[i] = std::fabs(f[i]);
  cdf[0] = 0.0f;
  for (size_t i = 1; i < n + 1; i++) cdf[i] = cdf[i - 1] + func[i - 1] * (mx - mn) / float(n);
  float integral = cdf[n];
  if (integral == 0.0f) {
    for (size_t i = 1; i < n + 1; i++) cdf[i] = float(i) / float(n);
  } else {
    for (size_t i = 1; i < n + 1; i++) cdf[i] /= integral;
  }
  return integral;
}

// fn(first, last) over [0, n) on the host's cores (big meshes only: the loops below are copies and gathers)
template <class F>
static void parallelRanges(size_t n, const F& fn) {
  unsigned threads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if (n < 100000 || threads == 1) {
    fn(size_t(0), n);
    return;
  }
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < threads; t++) pool.emplace_back([&, t] { fn(n * t / threads, n * (t + 1) / threads); });
  for (auto& th : pool) th.join();
}

// ---- SAH build on the device (csrc/bvh_build.cuh through the C ABI) ---------------------------------------------------
static std::atomic<int> gBuildDevice{0};
int setBuildDevice(int device) {
  gBuildDevice.store(device);
  return 0;
}

// YS_BVH_SAH: the GPU for meshes worth the round trip, if there is one and the vertex data is finite (the device
// build's min / max atomics assume an ordered set; the host builder folds NaNs the way the reference does).
// YART_B200_BVH_DEVICE=0 keeps every build on the host cores.
static bool autoDeviceBuild(const std::vector<float>& positions, size_t nTris) {
#ifdef YB_HOSTSIM
  (void)positions, (void)nTris;
  return false;  // the CPU build of the device layer runs the stages as plain loops: the host builder is faster there
#else
  const char* e = getenv("YART_B200_BVH_DEVICE");
  if (e && *e == '0') return false;
  if (gBuildDevice.load() < 0 || nTris < 32768) return false;
  for (float v : positions)
    if (!(std::fabs(v) <= std::numeric_limits<float>::max())) return false;
  return true;
#endif
}

// yc_build_bvh_sah delivers the reference's node array and index order as they are.
static_assert(sizeof(RefBvhNode) == sizeof(YcBuildNode), "RefBvhNode and YcBuildNode are the reference's BVHNode");
static bool buildOnDevice(const float* positions, size_t nVerts, const uint32_t* faces4, size_t nTris, BvhBuildResult& out,
                          std::string* err) {
  uint32_t nNodes = 0, levels = 0;
  out.indices.resize(nTris);
  out.nodes.resize(2 * nTris + 2);
  const int dev = std::max(0, gBuildDevice.load());
  const int rc = yc_build_bvh_sah(dev, positions, nVerts, faces4, nTris, reinterpret_cast<YcBuildNode*>(out.nodes.data()), &nNodes,
                                  out.indices.data(), &levels);
  if (rc != YC_OK) {
    if (err) *err = yc_build_last_error();
    return false;
  }
  out.nodes.resize(nNodes);
  return true;
}

bool HostScene::build(const ysc::SceneDesc& d, std::string* err) {
  auto fail = [&](const std::string& m) {
    if (err) *err = m;
    return false;
  };
  if (d.nodes.empty()) return fail("scene has no root node");
  if (!loadLuts(err)) return false;

  // ---- textures ----------------------------------------------------------------------
  for (const auto& t : d.textures) {
    YcTexture y{};
    y.width = t.width, y.height = t.height, y.channels = t.channels, y.isFloat = t.isFloat, y.type = t.type;
    if (t.width < 2 || t.height < 2) return fail("textures must be at least 2x2 (bilinear taps x+1,y+1)");
    if (t.isFloat) {
      y.offset = texelsF32.size();
      texelsF32.insert(texelsF32.end(), t.f32.begin(), t.f32.end());
    } else {
      y.offset = texelsU8.size();
      texelsU8.insert(texelsU8.end(), t.u8.begin(), t.u8.end());
    }
    textures.push_back(y);
  }
  auto checkTex = [&](int idx, uint32_t ch, bool isFloat) {
    return idx < 0 || (size_t(idx) < d.textures.size() && d.textures[idx].channels == ch &&
                       (d.textures[idx].isFloat != 0) == isFloat);
  };

  // ---- materials: ParametricBSDF ctor (parametric.cpp:11-66) -----------------------------
  int anyAlpha = 0;
  for (const auto& m : d.materials) {
    if (!checkTex(m.baseTex, 4, false) || !checkTex(m.mrTex, 2, false) || !checkTex(m.transTex, 1, false) ||
        !checkTex(m.normalTex, 3, false) || !checkTex(m.ccTex, 1, false) || !checkTex(m.emisTex, 3, false))
      return fail("material texture index / channel count mismatch");
    YcMaterial y{};
    memcpy(y.base, m.base, 12);
    y.metallic = m.metallic, y.roughness = m.roughness, y.transmission = m.transmission, y.ior = m.ior;
    y.anisotropic = m.anisotropic, y.clearcoat = m.clearcoat, y.clearcoatRoughness = m.clearcoatRoughness;
    memcpy(y.emission, m.emission, 12);
    y.normalScale = m.normalScale;
    memcpy(y.volumeColor, m.volumeColor, 12);
    y.volumeDensity = m.volumeDensity;
    Mat4 lr = rotation(-m.anisoRotation, f3(0, 0, 1)), ir = rotation(m.anisoRotation, f3(0, 0, 1));
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        y.localRotation[i * 3 + j] = lr(i, j);
        y.invRotation[i * 3 + j] = ir(i, j);
      }
    y.baseTex = m.baseTex, y.mrTex = m.mrTex, y.transTex = m.transTex, y.normalTex = m.normalTex;
    y.ccTex = m.ccTex, y.emisTex = m.emisTex;
    y.thinTransmission = m.thinTransmission != 0;
    y.hasAlpha = 0;
    if (m.baseTex >= 0) {
      const auto& t = d.textures[m.baseTex];
      for (size_t i = 3; i < t.u8.size(); i += 4)
        if (t.u8[i] < 255) y.hasAlpha = 1;
    }
    y.hasEmission = length2(f3(m.emission)) > 0.0f;
    anyAlpha |= y.hasAlpha;
    materials.push_back(y);
  }

  // ---- meshes + BVH ------------------------------------------------------------------
  auto t0 = std::chrono::high_resolution_clock::now();
  std::vector<Bounds3> meshVertexBounds;
  for (const auto& m : d.meshes) {
    size_t nv = m.nVerts(), nf = m.nFaces();
    if (nf == 0 || nv == 0) return fail("empty mesh");
    for (size_t i = 0; i < nf; i++) {
      for (int k = 0; k < 3; k++)
        if (m.faces[4 * i + k] >= nv) return fail("face index out of range");
      if (m.faces[4 * i + 3] >= d.materials.size()) return fail("material index out of range");
      if (m.lightIdx[i] >= int32_t(d.lights.size())) return fail("light index out of range");
    }
    const char* traceEnv = getenv("YART_B200_BUILD_TRACE");
    const bool trace = traceEnv && *traceEnv && *traceEnv != '0';
    auto tick = [] { return std::chrono::high_resolution_clock::now(); };
    auto since = [&](std::chrono::high_resolution_clock::time_point t) { return std::chrono::duration<double, std::milli>(tick() - t).count(); };
    const auto tm0 = tick();
    YcMesh ym{};
    ym.vertOffset = uint32_t(positions.size() / 3);
    ym.primOffset = uint32_t(primMaterial.size());
    ym.nodeOffset = uint32_t(bvhNodes.size());
    ym.triOffset = uint32_t(bvhTris.size());
    ym.nTris = uint32_t(nf), ym.nVerts = uint32_t(nv);

    positions.insert(positions.end(), m.positions.begin(), m.positions.end());
    {
      // interleaved (normal, tangent, uv) per vertex → the three SoA arrays; (i0, i1, i2, material) per face → two
      const size_t n0 = normals.size(), t0 = tangents.size(), u0 = uvs.size();
      normals.resize(n0 + 3 * nv), tangents.resize(t0 + 4 * nv), uvs.resize(u0 + 2 * nv);
      float *pn = normals.data() + n0, *pt = tangents.data() + t0, *pu = uvs.data() + u0;
      const float* vd = m.vertexData.data();
      parallelRanges(nv, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
          const float* v = vd + 9 * i;
          pn[3 * i] = v[0], pn[3 * i + 1] = v[1], pn[3 * i + 2] = v[2];
          pt[4 * i] = v[3], pt[4 * i + 1] = v[4], pt[4 * i + 2] = v[5], pt[4 * i + 3] = v[6];
          pu[2 * i] = v[7], pu[2 * i + 1] = v[8];
        }
      });
      const size_t i0 = primIndices.size(), m0 = primMaterial.size();
      primIndices.resize(i0 + 3 * nf), primMaterial.resize(m0 + nf);
      uint32_t *pi = primIndices.data() + i0, *pm = primMaterial.data() + m0;
      const uint32_t* fd = m.faces.data();
      parallelRanges(nf, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) pi[3 * i] = fd[4 * i], pi[3 * i + 1] = fd[4 * i + 1], pi[3 * i + 2] = fd[4 * i + 2], pm[i] = fd[4 * i + 3];
      });
      primLight.insert(primLight.end(), m.lightIdx.begin(), m.lightIdx.end());
    }
    Bounds3 vb;  // Node(Mesh*) ctor: unpadded vertex bounds (scene.hpp:17-22)
    for (size_t i = 0; i < nv; i++) vb.expandToInclude(f3(&m.positions[3 * i]));
    meshVertexBounds.push_back(vb);

    const double msAttr = since(tm0);
    const auto tm1 = tick();
    BvhBuildResult ref;
    bool onDevice = false;
    if (bvhKind == YS_BVH_SAH_DEVICE || (bvhKind == YS_BVH_SAH && autoDeviceBuild(m.positions, nf))) {
      std::string berr;
      onDevice = buildOnDevice(m.positions.data(), nv, m.faces.data(), nf, ref, &berr);
      if (!onDevice && bvhKind == YS_BVH_SAH_DEVICE) return fail(("device BVH build: " + berr).c_str());
    }
    if (!onDevice) {
      SahBvhBuilder builder;
      builder.kind = bvhKind == YS_BVH_MEDIAN_SPLIT ? SahBvhBuilder::kMedianSplit : SahBvhBuilder::kSah;
      ref = builder.build(m.positions.data(), nv, m.faces.data(), nf);
    }
    deviceBuilds += onDevice ? 1u : 0u;
    const double msBvh = since(tm1);
    const auto tm2 = tick();

    // reference nodes → inner-node records with both children inlined; leaves → contiguous tri runs
    std::vector<uint32_t> innerRank(ref.nodes.size(), 0);
    uint32_t nInner = 0;
    for (size_t i = 0; i < ref.nodes.size(); i++)
      if (ref.nodes[i].span == 0) innerRank[i] = nInner++;
    ym.nInner = nInner;
    auto refOf = [&](uint32_t node) -> uint32_t {
      const RefBvhNode& n = ref.nodes[node];
      return n.span == 0 ? innerRank[node] : (YC_REF_LEAF | n.leftFirst);
    };
    for (int k = 0; k < 3; k++) {
      ym.rootMin[k] = ref.nodes[0].mn[k];
      ym.rootMax[k] = ref.nodes[0].mx[k];
    }
    ym.rootRef = refOf(0);
    size_t base = bvhNodes.size();
    bvhNodes.resize(base + nInner);
    for (size_t i = 0; i < ref.nodes.size(); i++) {
      const RefBvhNode& n = ref.nodes[i];
      if (n.span != 0) continue;
      YcBvhNode& o = bvhNodes[base + innerRank[i]];
      const RefBvhNode &c0 = ref.nodes[n.leftFirst], &c1 = ref.nodes[n.leftFirst + 1];
      memcpy(o.c0min, c0.mn, 12), memcpy(o.c0max, c0.mx, 12);
      memcpy(o.c1min, c1.mn, 12), memcpy(o.c1max, c1.mx, 12);
      o.ref0 = refOf(n.leftFirst), o.ref1 = refOf(n.leftFirst + 1);
      o.pad0 = o.pad1 = 0;
    }
    size_t tbase = bvhTris.size();
    bvhTris.resize(tbase + nf);
    parallelRanges(nf, [&](size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; i++) {
        uint32_t prim = ref.indices[i];
        YcBvhTri& t = bvhTris[tbase + i];
        memcpy(t.p0, &m.positions[3 * size_t(m.faces[4 * prim + 0])], 12);
        memcpy(t.p1, &m.positions[3 * size_t(m.faces[4 * prim + 1])], 12);
        memcpy(t.p2, &m.positions[3 * size_t(m.faces[4 * prim + 2])], 12);
        t.prim = prim;
        const YcMaterial& mat = materials[m.faces[4 * prim + 3]];
        t.flags = (mat.hasAlpha ? YC_TRI_ALPHA : 0) |
                  ((mat.thinTransmission && mat.transmission > 0.0f) ? YC_TRI_TRANSPARENT : 0);
        t.pad = 0;
      }
    });
    for (const RefBvhNode& n : ref.nodes)
      if (n.span != 0) bvhTris[tbase + n.leftFirst + n.span - 1].flags |= YC_TRI_LAST;
    meshes.push_back(ym);
    refBvh.push_back(std::move(ref));
    if (trace && nf >= 1000)
      fprintf(stderr, "yart_b200 scene build: mesh of %zu tris: attributes %.1f ms, BVH (%s) %.1f ms, flatten %.1f ms\n", nf, msAttr,
              onDevice ? "GPU" : "host", msBvh, since(tm2));
  }
  buildMs = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();

  // ---- node tree → DFS pre-order (children in file order) ----------------------------------
  struct Built {
    Transform xf;
    Bounds3 bounds;
  };
  std::vector<std::vector<int>> children(d.nodes.size());
  for (size_t i = 1; i < d.nodes.size(); i++) {
    int p = d.nodes[i].parent;
    if (p < 0 || size_t(p) >= i) return fail("node parents must precede children; node 0 is the root");
    children[p].push_back(int(i));
  }
  std::string nodeErr;
  std::function<Built(int, int, int)> visit = [&](int idx, int parentFlat, int depth) -> Built {
    const auto& nd = d.nodes[idx];
    Built b;
    if (nd.hasTransform) b.xf = Transform(mat16(nd.m));
    if (nd.mesh >= 0) {
      if (size_t(nd.mesh) >= meshes.size()) nodeErr = "node mesh index out of range";
      else b.bounds = meshVertexBounds[nd.mesh];
    }
    if (depth >= YC_MAX_NODE_DEPTH) nodeErr = "scene graph deeper than YC_MAX_NODE_DEPTH";
    int self = int(nodes.size());
    nodes.push_back(YcNode{});
    {
      // known before the children are visited: they look at their parent's flag
      static const float kIdentityRows[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
      const bool selfIdentity = memcmp(b.xf.inv.m, kIdentityRows, sizeof kIdentityRows) == 0 &&
                                memcmp(b.xf.fwd.m, kIdentityRows, sizeof kIdentityRows) == 0;
      nodes[self].identityChain = selfIdentity && (parentFlat < 0 || nodes[parentFlat].identityChain);
    }
    for (int c : children[idx]) {
      Built cb = visit(c, self, depth + 1);
      b.bounds = Bounds3::join(b.bounds, cb.xf.bounds(cb.bounds));  // Node::appendChild, scene.hpp:54-58
    }
    YcNode& y = nodes[self];
    putRows(y.inv, b.xf.inv);
    putRows(y.fwd, b.xf.fwd);
    memcpy(y.nrm, b.xf.nrm.m, sizeof(y.nrm));
    put3(y.bmin, b.bounds.mn);
    put3(y.bmax, b.bounds.mx);
    y.mesh = nd.mesh, y.parent = parentFlat, y.skip = int(nodes.size()), y.depth = depth;

    return b;
  };
  visit(0, -1, 0);
  if (!nodeErr.empty()) return fail(nodeErr);

  // ---- lights --------------------------------------------------------------------------
  for (const auto& l : d.lights) {
    YcLight y{};
    y.type = l.type, y.twoSided = l.twoSided;
    memcpy(y.emission, l.emission, 12);
    y.hdrTex = -1;
    Transform xf = l.hasTransform ? Transform(mat16(l.m)) : Transform();
    if (l.type == ysc::AreaLightT) {
      if (l.mesh < 0 || size_t(l.mesh) >= d.meshes.size() || l.tri < 0 || size_t(l.tri) >= d.meshes[l.mesh].nFaces())
        return fail("area light references a missing triangle");
      const auto& m = d.meshes[l.mesh];
      const uint32_t* f = &m.faces[4 * size_t(l.tri)];
      f3 p[3], n[3];
      for (int k = 0; k < 3; k++) {
        p[k] = f3(&m.positions[3 * size_t(f[k])]);
        n[k] = f3(&m.vertexData[9 * size_t(f[k])]);
      }
      put3(y.p0, p[0]), put3(y.p1, p[1]), put3(y.p2, p[2]);
      put3(y.n0, n[0]), put3(y.n1, n[1]), put3(y.n2, n[2]);
      putRows(y.fwd, xf.fwd);
      memcpy(y.nrm, xf.nrm.m, sizeof(y.nrm));
      // AreaLight ctor, light.cpp:16-32 (triangleArea: primitives.hpp:24-32)
      f3 t0p = xf.point(p[0]), t1p = xf.point(p[1]), t2p = xf.point(p[2]);
      y.area = length(cross(t1p - t0p, t2p - t0p)) * 0.5f;
      // AreaLight::power, light.cpp:38-40
      y.power = length(f3(l.emission)) * y.area * kPi * (l.twoSided ? 2.0f : 1.0f);
    } else if (l.type == ysc::ImageInfiniteT) {
      if (l.hdrTex < 0 || size_t(l.hdrTex) >= d.textures.size() || !d.textures[l.hdrTex].isFloat ||
          d.textures[l.hdrTex].channels != 3)
        return fail("image infinite light needs a 3-channel float texture");
      const auto& t = d.textures[l.hdrTex];
      y.hdrTex = l.hdrTex;
      y.sceneRadius = l.sceneRadius;
      putRows(y.envFwd, xf.fwd);
      putRows(y.envInv, xf.inv);
      // ImageInfiniteLight ctor, light.cpp:137-196 with the default bounds {0,0}-{1,1}
      uint32_t w = t.width, h = t.height;
      std::vector<float> dd(size_t(w) * h);
      f3 Lavg;
      for (uint32_t yy = 0; yy < h; yy++) {
        float v = (float(yy) + 0.5f) / float(h);
        float z = 1.0f - v * 2.0f;
        float sinTheta = std::sqrt(1.0f - z * z);
        for (uint32_t x = 0; x < w; x++) {
          f3 s(&t.f32[3 * (size_t(x) + size_t(yy) * w)]);
          float sum = 0.0f;
          sum += s.x, sum += s.y, sum += s.z;
          dd[size_t(yy) * w + x] = (sum / 3.0f) * sinTheta;
          Lavg = Lavg + s;
        }
      }
      Lavg = Lavg / float(w * h);
      put3(y.Lavg, Lavg);
      y.distW = w, y.distH = h;
      y.distOffset = envDist.size();
      size_t blk = size_t(w) * h + size_t(w + 1) * h + h + h + (h + 1) + 1;
      envDist.resize(envDist.size() + blk);
      float* func = &envDist[y.distOffset];
      float* cdf = func + size_t(w) * h;
      float* rowInt = cdf + size_t(w + 1) * h;
      float* mfunc = rowInt + h;
      float* mcdf = mfunc + h;
      float* mInt = mcdf + (h + 1);
      for (uint32_t yy = 0; yy < h; yy++)
        rowInt[yy] = buildDistribution1D(&dd[size_t(yy) * w], w, 0.0f, 1.0f, func + size_t(yy) * w,
                                         cdf + size_t(yy) * (w + 1));
      *mInt = buildDistribution1D(rowInt, h, 0.0f, 1.0f, mfunc, mcdf);
      // surface area, light.cpp:190-195
      float phi0 = 0.0f * 2.0f * kPi, phi1 = 1.0f * 2.0f * kPi;
      float theta0 = 0.0f * kPi, theta1 = 1.0f * kPi;
      y.surfaceArea = (phi1 - phi0) * (std::cos(theta0) - std::cos(theta1));
      float lsum = 0.0f;
      lsum += Lavg.x, lsum += Lavg.y, lsum += Lavg.z;
      y.power = y.surfaceArea * kPi * l.sceneRadius * l.sceneRadius * lsum / 3.0f;  // light.cpp:205-208
    } else if (l.type == ysc::UniformInfiniteT) {
      y.sceneRadius = l.sceneRadius;
      put3(y.Lavg, f3(l.emission));
      y.surfaceArea = 4.0f * kPi;
      y.power = 4.0f * kPi * kPi * l.sceneRadius * l.sceneRadius * length(f3(l.emission));  // light.cpp:99-104
    } else {
      return fail("unknown light type");
    }
    lights.push_back(y);
  }

  // ---- PowerLightSampler::init, light-sampler.cpp:32-50 ---------------------------------
  totalPower = 0.0f;
  for (size_t i = 0; i < lights.size(); i++) {
    if (lights[i].type != YC_LIGHT_AREA) {
      infiniteLights.push_back(uint32_t(i));
    } else {
      areaLights.push_back(uint32_t(i));
      lightPowerCdf.push_back(totalPower + lights[i].power);
      totalPower += lights[i].power;
    }
  }

  // ---- flat view ---------------------------------------------------------------------
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
  flat.lutTables = lutTables.data();
  flat.hasAlpha = anyAlpha;
  return true;
}

}  // namespace yartb
