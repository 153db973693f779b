// wide_bvh.cuh — the GPU-friendly wide-node layout of the reference's SAH BVH (BASELINE.json north_star:
// "flattens the same SAH BVH into a GPU-friendly wide-node layout") and its traversal.
//
// yc_upload_scene collapses every mesh's BVH2 (the reference's tree, src/core/bvh.hpp:21-33, as it arrives in
// YcScene::bvhNodes) into 4-wide nodes: same leaves (the same contiguous triangle runs of YcScene::bvhTris, so
// the triangle arithmetic and every accepted hit's t / u / v are the reference's), child boxes taken verbatim
// from the BVH2 nodes they came from, inner levels removed greedily by surface area.
//
//   WideNode (128 B, one L1 line):   row 0  lo.x of children 0..3      row 1  hi.x
//                                    row 2  lo.y                        row 3  hi.y
//                                    row 4  lo.z                        row 5  hi.z
//                                    row 6  child refs (YcBvhNode::ref encoding, kWideEmpty = unused slot)
//   A ray precomputes, per axis, which row of the pair holds its NEAR planes (row 2a + (d[a] < 0)), so a lane
//   loads near and far planes directly (6 x LDG.128, addresses differ by the sign bit) and the slab test has no
//   selects:  t = fma(plane, 1/d, -o/d)  (one rounding),  tn = max(tnx, tny, tnz, tMin),  tf = min(tfx, tfy, tfz, hit.t).
//
// What differs from the reference-order BVH2 walk (traverse.cuh), and why results still match:
//   * box culling only.  The set of triangles tested is a superset / reordering of the reference's, every
//     triangle test is the reference's arithmetic (testTriangle, -fmad=false), and the closest hit is the
//     minimum over accepted tests.  Hit ids / t / u / v are therefore identical except where two triangles tie
//     in t to the last bit, or where a hit sits on the boundary of a box that one of the two slab
//     formulations culls (the reference's `bmin * idir + odir` has two roundings and is not conservative
//     either).  tests/ count those and list them.
//   * a zero direction component is clamped to ±1e-20 for the BOX test only (Aila-Laine), so a ray with
//     d.x == 0 and o.x == 0 no longer has NaN slabs on that axis and no longer walks every box that overlaps
//     it in y and z (14 K boxes for one ray of the C2 camera, profiles/README.md "Pathological rays").
//   * used only for scenes without alpha-tested materials: there the order of triangle tests cannot move a
//     sampler draw (ray-integrator.cpp:211).  Alpha scenes keep the reference-order BVH2 walk.
//   * NEE rays through thin transmissive surfaces multiply Hit::attenuation in a different order
//     (floating-point product order; last-bit differences in those pixels).
#pragma once
#include <vector>

#include "traverse.cuh"

namespace yb {

constexpr uint32_t kWideEmpty = 0xfffffffdu;  // unused child slot (has the leaf bit; never equals a real leaf ref)

struct WideNode {
  float plane[6][4];
  uint32_t ref[4];
  uint32_t pad[4];
};
static_assert(sizeof(WideNode) == 128, "WideNode is one 128-byte line");

// ---- host: BVH2 → BVH4 collapse --------------------------------------------------------------------
// Appends mesh `m`'s wide nodes to `out` (children of a node are allocated together, parents before children)
// and returns the deepest level (root = 1; 0 for a single-leaf mesh).
inline int collapseToWide(const YcBvhNode* bvh2, const YcMesh& m, std::vector<WideNode>& out, WideMesh& wm) {
  wm.nodeOffset = uint32_t(out.size());
  wm.rootRef = m.rootRef;
  if (m.rootRef & YC_REF_LEAF) return 0;
  struct Child {
    float lo[3], hi[3];
    uint32_t ref;
  };
  struct Work {
    uint32_t ref2, wide;
    int depth;
  };
  const size_t base = out.size();
  out.emplace_back();
  std::vector<Work> todo{{m.rootRef, 0u, 1}};
  int maxDepth = 1;
  auto area = [](const Child& c) {
    const float dx = c.hi[0] - c.lo[0], dy = c.hi[1] - c.lo[1], dz = c.hi[2] - c.lo[2];
    return dx * dy + dy * dz + dz * dx;
  };
  while (!todo.empty()) {
    const Work w = todo.back();
    todo.pop_back();
    maxDepth = std::max(maxDepth, w.depth);
    Child c[4];
    int n = 0;
    auto put = [&](Child& dst, const float* lo, const float* hi, uint32_t ref) {
      memcpy(dst.lo, lo, 12), memcpy(dst.hi, hi, 12), dst.ref = ref;
    };
    {
      const YcBvhNode& nd = bvh2[w.ref2];
      put(c[0], nd.c0min, nd.c0max, nd.ref0), put(c[1], nd.c1min, nd.c1max, nd.ref1);
      n = 2;
    }
    while (n < 4) {
      int best = -1;
      float bestArea = -1.0f;
      for (int k = 0; k < n; k++)
        if (!(c[k].ref & YC_REF_LEAF) && area(c[k]) > bestArea) best = k, bestArea = area(c[k]);
      if (best < 0) break;
      // the opened child's slot receives its left child, its right child is inserted after it (left-to-right
      // order of the reference tree is kept)
      const YcBvhNode& nd = bvh2[c[best].ref];
      for (int k = n - 1; k > best; k--) c[k + 1] = c[k];
      put(c[best], nd.c0min, nd.c0max, nd.ref0), put(c[best + 1], nd.c1min, nd.c1max, nd.ref1);
      n++;
    }
    WideNode node{};
    for (int k = 0; k < 4; k++) {
      if (k < n) {
        for (int a = 0; a < 3; a++) node.plane[2 * a][k] = c[k].lo[a], node.plane[2 * a + 1][k] = c[k].hi[a];
        if (c[k].ref & YC_REF_LEAF) {
          node.ref[k] = c[k].ref;
        } else {
          node.ref[k] = uint32_t(out.size() - base);
          todo.push_back({c[k].ref, node.ref[k], w.depth + 1});
          out.emplace_back();
        }
      } else {
        for (int a = 0; a < 3; a++) node.plane[2 * a][k] = INFINITY, node.plane[2 * a + 1][k] = -INFINITY;
        node.ref[k] = kWideEmpty;
      }
    }
    out[base + w.wide] = node;
  }
  return maxDepth;
}

// ---- device ------------------------------------------------------------------------------------------
// Per-ray constants of the wide slab test.
struct WideRay {
  V3 idir, odir;     // 1 / d (zero components clamped to ±1e-20) and -o * idir
  uint32_t nx, ny, nz;  // row (float4 index inside the node) of the NEAR planes per axis; far = near ^ 1
  YB_DEV void set(V3 o, V3 d) {
    const float eps = 1e-20f;
    const float dx = fabsf(d.x) > eps ? d.x : copysignf(eps, d.x), dy = fabsf(d.y) > eps ? d.y : copysignf(eps, d.y),
                dz = fabsf(d.z) > eps ? d.z : copysignf(eps, d.z);
    idir = V3(1.0f / dx, 1.0f / dy, 1.0f / dz);
    odir = V3(-o.x * idir.x, -o.y * idir.y, -o.z * idir.z);
    nx = 0u + (dx < 0.0f ? 1u : 0u), ny = 2u + (dy < 0.0f ? 1u : 0u), nz = 4u + (dz < 0.0f ? 1u : 0u);
  }
};

#ifdef YB_HOSTSIM
YB_DEV float fmaExact(float a, float b, float c) { return fmaf(a, b, c); }
#else
YB_DEV float fmaExact(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#endif

// One box against the ray: entry distance in `tn`, true when the slab interval [max(tn, tMin), min(tf, tmx)] is
// not empty.  fmaxf / fminf drop NaN operands, like the reference's NaN-tolerant rmin / rmax.
YB_DEV bool slabWide(const WideRay& w, V3 nearP, V3 farP, float tmn, float tmx, float& tn) {
  const float tnx = fmaExact(nearP.x, w.idir.x, w.odir.x), tny = fmaExact(nearP.y, w.idir.y, w.odir.y),
              tnz = fmaExact(nearP.z, w.idir.z, w.odir.z);
  const float tfx = fmaExact(farP.x, w.idir.x, w.odir.x), tfy = fmaExact(farP.y, w.idir.y, w.odir.y),
              tfz = fmaExact(farP.z, w.idir.z, w.odir.z);
  tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmn));
  const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmx));
  return tn <= tf;
}
// lo / hi form (scene-graph node boxes, mesh root boxes)
YB_DEV bool slabWideBox(const WideRay& w, V3 lo, V3 hi, float tmn, float tmx, float& tn) {
  const bool sx = w.nx & 1u, sy = w.ny & 1u, sz = w.nz & 1u;
  return slabWide(w, V3(sx ? hi.x : lo.x, sy ? hi.y : lo.y, sz ? hi.z : lo.z),
                  V3(sx ? lo.x : hi.x, sy ? lo.y : hi.y, sz ? lo.z : hi.z), tmn, tmx, tn);
}

// Ordering key of a hit child: entry distance with the child slot in the two lowest mantissa bits (distances are
// >= tMin > 0, so keys order like the distances; ties go to the lower slot).  kWideMiss for a missed child.
constexpr uint32_t kWideMiss = 0xffffffffu;
YB_DEV uint32_t wideKey(bool hit, float tn, uint32_t slot) { return hit ? ((__float_as_uint(tn) & ~3u) | slot) : kWideMiss; }

// Sequential walk of one mesh's wide BVH (per-path tail kernel, CPU build of the product sources); the
// persistent-warp kernels (trace_wide.cuh) make the same decisions: nearest hit child first (by key), the other
// hit children pushed in slot order with their entry distances, popped entries culled by `d < hit.t`.
template <bool NEE, bool COUNT, bool EARLY_OUT>
YB_DEV bool testBVHWide(const DScene& sc, const YcMesh& mesh, const WideMesh& wm, const LocalRay& r, const WideRay& w,
                        int nodeIdx, TraceState& st, TravStack& stack, TraceCounters& cnt) {
  float d;
  if (COUNT) cnt.box++;
  if (!slabWideBox(w, V3(mesh.rootMin), V3(mesh.rootMax), kTMin, st.hit.t, d)) return false;
  const float4* __restrict__ nodes = sc.wideNodes + 8 * size_t(wm.nodeOffset);
  const float4* __restrict__ tris = sc.bvhTris + 3 * size_t(mesh.triOffset);
  uint32_t cur = wm.rootRef;
  int sp = 0;
  bool didHit = false;
  while (true) {
    if (d < st.hit.t) {
      if (cur & YC_REF_LEAF) {
        uint32_t ti = cur & ~YC_REF_LEAF;
        while (true) {
          const float4 a = __ldg(tris + 3 * size_t(ti)), b = __ldg(tris + 3 * size_t(ti) + 1),
                       c = __ldg(tris + 3 * size_t(ti) + 2);
          didHit |= testTriangle<NEE, false, COUNT>(sc, mesh, r, a, b, c, nodeIdx, st, nullptr, cnt);
          if (NEE && didHit) break;
          if (__float_as_uint(c.z) & YC_TRI_LAST) break;
          ti++;
        }
        if (NEE && EARLY_OUT && didHit) return true;
        if (sp == 0) break;
        stack.pop(--sp, cur, d);
      } else {
        const float4* n = nodes + 8 * size_t(cur);
        const float4 nX = __ldg(n + w.nx), fX = __ldg(n + (w.nx ^ 1u)), nY = __ldg(n + w.ny), fY = __ldg(n + (w.ny ^ 1u)),
                     nZ = __ldg(n + w.nz), fZ = __ldg(n + (w.nz ^ 1u)), rf = __ldg(n + 6);
        const uint32_t ref[4] = {__float_as_uint(rf.x), __float_as_uint(rf.y), __float_as_uint(rf.z), __float_as_uint(rf.w)};
        float tn[4];
        bool hit[4];
        hit[0] = slabWide(w, V3(nX.x, nY.x, nZ.x), V3(fX.x, fY.x, fZ.x), kTMin, st.hit.t, tn[0]);
        hit[1] = slabWide(w, V3(nX.y, nY.y, nZ.y), V3(fX.y, fY.y, fZ.y), kTMin, st.hit.t, tn[1]);
        hit[2] = slabWide(w, V3(nX.z, nY.z, nZ.z), V3(fX.z, fY.z, fZ.z), kTMin, st.hit.t, tn[2]) && ref[2] != kWideEmpty;
        hit[3] = slabWide(w, V3(nX.w, nY.w, nZ.w), V3(fX.w, fY.w, fZ.w), kTMin, st.hit.t, tn[3]) && ref[3] != kWideEmpty;
        if (COUNT) cnt.box += 2u + (ref[2] != kWideEmpty) + (ref[3] != kWideEmpty);
        uint32_t best = kWideMiss;
        for (uint32_t k = 0; k < 4; k++) {
          const uint32_t key = wideKey(hit[k], tn[k], k);
          best = key < best ? key : best;
        }
        if (best == kWideMiss) {
          if (sp == 0) break;
          stack.pop(--sp, cur, d);
        } else {
          const uint32_t nearSlot = best & 3u;
          for (uint32_t k = 0; k < 4; k++)
            if (hit[k] && k != nearSlot) stack.push(sp++, ref[k], tn[k]);
          cur = ref[nearSlot];
          d = __uint_as_float(best & ~3u);
        }
      }
    } else {
      if (sp == 0) break;
      stack.pop(--sp, cur, d);
    }
  }
  return didHit;
}

// testNode over the whole scene graph with the wide walk per mesh (traceScene's counterpart, traverse.cuh).
template <bool NEE, bool COUNT, bool EARLY_OUT>
YB_DEV bool traceSceneWide(const DScene& sc, V3 origin, V3 dir, TraceState& st, TravStack& stack, TraceCounters& cnt) {
  V3 ro[YC_MAX_NODE_DEPTH + 1], rd[YC_MAX_NODE_DEPTH + 1];
  ro[0] = origin;
  rd[0] = dir;
  bool didHit = false;
  uint32_t i = 0;
  while (i < sc.nNodes) {
    const YcNode& nd = sc.nodes[i];
    const int k = nd.depth;
    LocalRay r;
    r.o = xformRows(nd.inv, ro[k], 1.0f), r.d = xformRows(nd.inv, rd[k], 0.0f);  // ray-integrator.cpp:26-30
    ro[k + 1] = r.o;
    rd[k + 1] = r.d;
    WideRay w;
    w.set(r.o, r.d);
    float d;
    if (COUNT) cnt.box++;
    if (!slabWideBox(w, V3(nd.bmin), V3(nd.bmax), kTMin, st.hit.t, d) || st.hit.t < d) {
      i = uint32_t(nd.skip);
      continue;
    }
    if (nd.mesh >= 0) {
      const bool h = testBVHWide<NEE, COUNT, EARLY_OUT>(sc, sc.meshes[nd.mesh], sc.wideMeshes[nd.mesh], r, w, int(i), st, stack, cnt);
      didHit |= h;
      if (NEE && EARLY_OUT && h) return true;
    }
    i++;
  }
  return didHit;
}

}  // namespace yb
