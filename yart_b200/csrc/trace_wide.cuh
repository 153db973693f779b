// trace_wide.cuh — persistent-warp traversal of the 4-wide quantised BVH (wide_bvh.cuh), CUDA only: the extend,
// shadow and ray-hook kernels of scenes without alpha-tested materials.
//
// Same warp scheduling as trace_kernels.cuh (lane state machine IDLE → NODE → TRAV, one atomic per warp to
// refill idle lanes, inner-node steps in a tight loop while enough lanes sit on inner nodes, speculative leaf
// parking, packed (ref, d) stack entries in shared memory with a sentinel at the bottom); what changes is the
// inner step: one 64-byte node = 2 x LDG.256 per lane (two L1 tag look-ups per visit instead of the BVH2 walk's
// two per HALF as much tree), four boxes decoded with 24 PRMT + 24 FFMA, 6 selects for near / far,
// 16 FMNMX(3) + 4 FSETP; 4 keys (entry distance | slot), unsigned min → nearest hit child; the other hit children
// are pushed with their entry distances (predicated STS.64), popped entries are culled by `d < hit.t`.
// Triangles are the reference's arithmetic (testTriangle).  Results versus the reference-order walk: see the
// header of wide_bvh.cuh.
#pragma once
#include "trace_kernels.cuh"

namespace yb {

struct Words8 {
  uint32_t w[8];
};
__device__ __forceinline__ Words8 loadHalfNode(const float4* p) {  // 32 bytes, one sector
  Words8 r;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
               : "l"(p));
  return r;
}

// IO policy as in tracePersistent; the sampler argument of load / store is a dummy (no alpha-tested materials).
template <bool NEE, bool COUNT, class IO>
__device__ __forceinline__ void traceWidePersistent(const DScene& sc, IO& io, uint32_t n, uint32_t* head, uint2* spillBase,
                                                    const TraceTuning tune, TraceCounters& cnt) {
  __shared__ uint2 shStack[kPsStack * kTraceBlock];
  __shared__ uint2* shSpill;
  WarpStack stack;
  stack.init(shStack, &shSpill, spillBase, tune.shEntries);
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u, ltMask = (1u << lane) - 1u;

  int state = kLaneIdle;
  bool exhausted = false;
  uint32_t item = 0, nextNode = 0, cur = 0;
  uint32_t sp = 0;
  int curNode = 0;
  float dcur = 0.0f;
  bool didHit = false, meshHit = false, rayIsWorld = false, worldFinite = false;
  V3 wo, wd;
  LocalRay r;  // o and d only (triangle tests); the box tests use `wr`
  WideRay wr;
  TraceState st;
  Sampler smp;
  const float4* __restrict__ nodes = nullptr;
  const float4* __restrict__ tris = nullptr;
  uint32_t meshIdx = 0;
  constexpr uint32_t kNoRef = 0xffffffffu;   // "no current node" (leaf bit set: never stepped as inner)
  constexpr uint32_t kPopRef = 0xfffffffeu;  // "pop at the top of the next inner step"
  uint32_t pend = 0u;                        // parked leaf (0 = none)
  float pendD = 0.0f;
  // bits of 1.0f in a register ptxas cannot fold (shEntries <= 25), so that planeUnit's PRMT carries its selector as the
  // immediate instead of fetching four selectors into registers on every visit
  const uint32_t one = 0x3f800000u | (uint32_t(tune.shEntries) >> 30);

  for (;;) {
    // ---- refill idle lanes -------------------------------------------------------------------
    const unsigned idle = __ballot_sync(FULL, state == kLaneIdle);
    if (idle) {
      if (!exhausted && (__popc(idle) >= tune.refillMin || idle == FULL)) {
        const uint32_t want = uint32_t(__popc(idle));
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(head, want);
        base = __shfl_sync(FULL, base, 0);
        if (state == kLaneIdle) {
          const uint32_t j = base + uint32_t(__popc(idle & ltMask));
          if (j < n) {
            float tMax;
            if (io.load(j, wo, wd, tMax, smp)) {
              item = j;
              rayIsWorld = false;
              worldFinite = isfinite(wo.x) && isfinite(wo.y) && isfinite(wo.z) && isfinite(wd.x) && isfinite(wd.y) &&
                            isfinite(wd.z);
              initTraceState(st, tMax);
              nextNode = 0;
              didHit = false;
              state = kLaneNode;
            }
          }
        }
        if (base + want >= n) exhausted = true;
      }
      if (exhausted && __ballot_sync(FULL, state != kLaneIdle) == 0) break;
    }

    // ---- scene-graph step (testNode): next node whose box the ray enters ---------------------------
    if (state == kLaneNode) {
      bool entered = false;
      while (nextNode < sc.nNodes) {
        const YcNode& nd = sc.nodes[nextNode];
        if (nd.identityChain && worldFinite) {
          if (!rayIsWorld) {
            r.o = wo + 0.0f, r.d = wd + 0.0f;  // what the identity matrix products leave (trace_kernels.cuh)
            wr.set(r.o, r.d);
            rayIsWorld = true;
          }
        } else {
          V3 o = wo, d = wd;
          nodeLocalRay(sc, nextNode, nd.depth, o, d);
          r.o = o, r.d = d;
          wr.set(o, d);
          rayIsWorld = false;
        }
        float dd;
        if (COUNT) cnt.box++;
        if (!slabWideBox(wr, V3(nd.bmin), V3(nd.bmax), kTMin, st.hit.t, dd) || st.hit.t < dd) {
          nextNode = uint32_t(nd.skip);
          continue;
        }
        const int mi = nd.mesh;
        curNode = int(nextNode);
        nextNode++;
        if (mi < 0) continue;
        const YcMesh& mesh = sc.meshes[mi];
        if (COUNT) cnt.box++;
        if (!slabWideBox(wr, V3(mesh.rootMin), V3(mesh.rootMax), kTMin, st.hit.t, dd)) continue;
        const WideMesh wm = sc.wideMeshes[mi];
        meshIdx = uint32_t(mi);
        nodes = sc.wideNodes + 4 * size_t(wm.nodeOffset);
        tris = sc.bvhTris + 3 * size_t(mesh.triOffset);
        cur = wm.rootRef;
        dcur = dd;
        stack.put(stack.base, kNoRef, 0.0f);  // sentinel
        sp = stack.base + kPsStride;
        pend = 0u;
        meshHit = false;
        entered = true;
        break;
      }
      if (entered) {
        state = kLaneTrav;
      } else {
        io.store(item, st, didHit, smp);
        state = kLaneIdle;
      }
    }

    // ---- inner-node steps --------------------------------------------------------------------------
    for (;;) {
      const bool trav = state == kLaneTrav;
      bool doPop = trav && cur == kPopRef;
      if (trav && pend == 0u && int32_t(cur) < -3) {  // a leaf (bit 31) other than kNoRef / kPopRef / kWideEmpty: park it
        pend = cur, pendD = dcur;
        doPop = true;
      }
      stack.popIf(doPop, sp, cur, dcur);  // the sentinel ends the mesh: cur = kNoRef
      const bool inner = trav && int32_t(cur) >= 0;
      const unsigned im = __ballot_sync(FULL, inner);
      if (im == 0) break;
      if (__popc(im) < tune.innerMin && __ballot_sync(FULL, trav && !inner)) break;
      if (inner) {
        const float4* np = nodes + 4 * size_t(cur);
        const Words8 a = loadHalfNode(np), b = loadHalfNode(np + 2);  // {p', S, q.x} and {q.y, q.z, refs}
        const uint4 rf = make_uint4(b.w[4], b.w[5], b.w[6], b.w[7]);
        // pop-time cull `d < hit.t`: a dead entry hits nothing
        const bool live = dcur < st.hit.t;
        const WideSlabs sl = slabWide4(wr, one, __uint_as_float(a.w[0]), __uint_as_float(a.w[1]), __uint_as_float(a.w[2]),
                                       __uint_as_float(a.w[3]), __uint_as_float(a.w[4]), __uint_as_float(a.w[5]), a.w[6], a.w[7], b.w[0],
                                       b.w[1], b.w[2], b.w[3], kTMin, st.hit.t, live);
        const float t0 = sl.tn[0], t1 = sl.tn[1], t2 = sl.tn[2], t3 = sl.tn[3];
        const bool h0 = sl.hit[0], h1 = sl.hit[1], h2 = sl.hit[2] && rf.z != kWideEmpty, h3 = sl.hit[3] && rf.w != kWideEmpty;
        if (COUNT && live) cnt.box += 2u + (rf.z != kWideEmpty) + (rf.w != kWideEmpty);
        // nearest hit child (lowest slot among equal entry distances) → next node; the others are pushed in slot order
        const float a0 = h0 ? t0 : INFINITY, a1 = h1 ? t1 : INFINITY, a2 = h2 ? t2 : INFINITY, a3 = h3 ? t3 : INFINITY;
        const float best = fminf(fminf(a0, a1), fminf(a2, a3));
        const bool any = h0 || h1 || h2 || h3;
        const bool p0 = h0 && a0 == best, p1 = !p0 && h1 && a1 == best, p2 = !p0 && !p1 && h2 && a2 == best;
        const uint32_t nearRef = p0 ? rf.x : (p1 ? rf.y : (p2 ? rf.z : rf.w));
        const bool q0 = h0 && !p0, q1 = h1 && !p1, q2 = h2 && !p2, q3 = h3 && (p0 || p1 || p2);
        if (sp + 3u * kPsStride < stack.limit) {  // the (at most three) pushes fit in shared memory
          uint32_t sa = sp;
          if (q0) stack.put(sa, rf.x, t0), sa += kPsStride;
          if (q1) stack.put(sa, rf.y, t1), sa += kPsStride;
          if (q2) stack.put(sa, rf.z, t2), sa += kPsStride;
          if (q3) stack.put(sa, rf.w, t3), sa += kPsStride;
          sp = sa;
        } else {
          stack.pushIf(q0, sp, rf.x, t0);
          stack.pushIf(q1, sp, rf.y, t1);
          stack.pushIf(q2, sp, rf.z, t2);
          stack.pushIf(q3, sp, rf.w, t3);
        }
        cur = any ? nearRef : kPopRef;
        dcur = best;
      }
    }

    // ---- leaf step ---------------------------------------------------------------------------------
    if (state == kLaneTrav && (int32_t(cur) < 0 || pend != 0u)) {
      {
        const uint32_t leaf = pend;
        const float leafD = pendD;
        if (int32_t(leaf) < -3 && leafD < st.hit.t && !(NEE && didHit)) {
          const YcMesh& mesh = sc.meshes[meshIdx];
          uint32_t ti = leaf & ~YC_REF_LEAF;
          while (true) {
            const float4 a = __ldg(tris + 3 * size_t(ti)), b = __ldg(tris + 3 * size_t(ti) + 1),
                         c = __ldg(tris + 3 * size_t(ti) + 2);
            meshHit |= testTriangle<NEE, false, COUNT>(sc, mesh, r, a, b, c, curNode, st, &smp, cnt);
            if (NEE && meshHit) break;
            if (__float_as_uint(c.z) & YC_TRI_LAST) break;
            ti++;
          }
          didHit |= meshHit;
        }
      }
      pend = 0u;
      if (NEE && didHit) {  // no alpha-tested materials: the first occluder decides (shadowStage, integrator.cuh)
        io.store(item, st, true, smp);
        state = kLaneIdle;
      } else if (cur == kNoRef) {
        state = kLaneNode;  // mesh walked and nothing parked any more
      }
    }
  }
}

}  // namespace yb
