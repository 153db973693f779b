// trace_kernels.cuh — persistent-warp traversal engine (CUDA only) shared by the extend, shadow and
// ray-hook kernels.
//
// Every ray performs exactly the reference's sequence of box and triangle tests
// (RayIntegrator::testNode / testBVH / testTriangle, src/cpu/ray-integrator.cpp:20-229, through the
// arithmetic in traverse.cuh); what changes is only how the 32 lanes of a warp are kept busy:
//   * each lane is a small state machine: IDLE → NODE (scene-graph node set-up) → TRAV (inside one
//     mesh's BVH: current ref is an inner node or a leaf) → … → IDLE;
//   * idle lanes are refilled from the work queue with ONE atomic per warp (ballot + popc), as soon
//     as enough lanes are idle, instead of waiting for the slowest ray of a batch of 32;
//   * inner-node steps run in a tight loop while enough lanes sit on inner nodes; lanes that reach a
//     leaf wait there (their own test order is unchanged) and all leaves are processed together, so
//     the two code paths do not serialise against each other on every iteration;
//   * the traversal stack lives in shared memory ([entry][thread], conflict-free) with a spill area
//     in global memory for trees deeper than kShStack; no local memory is used.
// Node and triangle records are fetched as 128-bit read-only loads (4 per inner node, 3 per triangle).
#pragma once
#include "integrator.cuh"

namespace yb {

constexpr int kTraceBlock = 128;       // 4 warps per CTA

// Scheduling knobs (they change only how lanes are grouped, never a ray's own test sequence).
struct TraceTuning {
  int refillMin;  // refill the warp from the queue when at least this many lanes are idle
  int innerMin;   // keep stepping inner nodes while at least this many lanes sit on one
};

struct WarpStack {
  uint32_t* shRef;  // + threadIdx.x
  float* shD;
  uint2* spill;     // this thread's kMaxStack - kShStack entries in global memory
  __device__ __forceinline__ void push(int sp, uint32_t ref, float d) {
    if (sp < kShStack) {
      shRef[sp * kTraceBlock] = ref;
      shD[sp * kTraceBlock] = d;
    } else if (sp < kMaxStack) {
      spill[sp - kShStack] = make_uint2(ref, __float_as_uint(d));
    }
  }
  __device__ __forceinline__ void pop(int sp, uint32_t& ref, float& d) const {
    if (sp < kShStack) {
      ref = shRef[sp * kTraceBlock];
      d = shD[sp * kTraceBlock];
    } else {
      const uint2 v = spill[sp - kShStack];
      ref = v.x;
      d = __uint_as_float(v.y);
    }
  }
};

// Ray in the object space of scene node `node`: the chain of inverse transforms root → node, applied
// in the recursion's order (ray-integrator.cpp:26-30).  The ancestor chain is a host-built table.
__device__ __forceinline__ void nodeLocalRay(const DScene& sc, uint32_t node, int depth, V3& o, V3& d) {
  const int32_t* __restrict__ path = sc.nodePath + size_t(node) * YC_MAX_NODE_DEPTH;
#pragma unroll 1
  for (int level = 0; level <= depth; level++) {
    const YcNode& nd = sc.nodes[path[level]];
    const V3 no = xformRows(nd.inv, o, 1.0f), ndir = xformRows(nd.inv, d, 0.0f);
    o = no;
    d = ndir;
  }
}

// One 64-byte inner-node record as two 256-bit read-only loads (LDG.E.256, new on sm_100): half the
// L1 wavefronts of four 128-bit loads — the traversal kernels are bound by the LSU data pipe.
struct NodeRec {
  float f[16];
};
__device__ __forceinline__ NodeRec loadNode(const float4* p) {
  NodeRec n;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(n.f[0]), "=f"(n.f[1]), "=f"(n.f[2]), "=f"(n.f[3]), "=f"(n.f[4]), "=f"(n.f[5]), "=f"(n.f[6]), "=f"(n.f[7])
               : "l"(p));
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(n.f[8]), "=f"(n.f[9]), "=f"(n.f[10]), "=f"(n.f[11]), "=f"(n.f[12]), "=f"(n.f[13]), "=f"(n.f[14]), "=f"(n.f[15])
               : "l"(p + 2));
  return n;
}

enum : int { kLaneIdle = 0, kLaneNode = 1, kLaneTrav = 2 };

// IO policy:  bool load(uint32_t j, V3& o, V3& d, float& tMax, Sampler& smp)   (false: nothing to trace)
//             void store(uint32_t j, const TraceState& st, bool didHit, const Sampler& smp)
template <bool NEE, bool ALPHA, bool COUNT, bool EARLY_OUT, class IO>
__device__ __forceinline__ void tracePersistent(const DScene& sc, IO& io, uint32_t n, uint32_t* head, uint2* spillBase,
                                                const TraceTuning tune, TraceCounters& cnt) {
  __shared__ uint32_t shRef[kShStack * kTraceBlock];
  __shared__ float shD[kShStack * kTraceBlock];
  WarpStack stack;
  stack.shRef = shRef + threadIdx.x;
  stack.shD = shD + threadIdx.x;
  stack.spill = spillBase + size_t(blockIdx.x * kTraceBlock + threadIdx.x) * (kMaxStack - kShStack);
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u, ltMask = (1u << lane) - 1u;

  int state = kLaneIdle;
  bool exhausted = false;
  uint32_t item = 0, nextNode = 0, cur = 0;
  int sp = 0, curNode = 0;
  float dcur = 0.0f;
  bool didHit = false, meshHit = false, rayIsWorld = false, worldFinite = false;
  V3 wo, wd;
  LocalRay r;
  TraceState st;
  Sampler smp;
  const float4* __restrict__ nodes = nullptr;
  const float4* __restrict__ tris = nullptr;
  uint32_t meshIdx = 0;
  constexpr bool SPEC = !COUNT;           // counting builds reproduce the reference's box-test counts
  constexpr uint32_t kNoRef = 0xffffffffu;  // "no current node" (has the leaf bit: never stepped as inner)
  uint32_t pend = 0u;                     // parked leaf (0 = none; leaf refs carry bit 31)
  float pendD = 0.0f;

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

    // ---- scene-graph step: find the next node whose box the ray enters (testNode) ------------------
    if (state == kLaneNode) {
      bool entered = false;
      while (nextNode < sc.nNodes) {
        const YcNode& nd = sc.nodes[nextNode];
        if (nd.identityChain && worldFinite) {
          // every transform from the root to this node is exactly the identity: the matrix products
          // ((0 + 1*x) + 0*y + 0*z) + 0*w reduce to x + 0 for finite inputs (-0 becomes +0), so the
          // local ray is the canonicalised world ray — computed once per ray and kept.
          if (!rayIsWorld) {
            r.set(wo + 0.0f, wd + 0.0f);
            rayIsWorld = true;
          }
        } else {
          V3 o = wo, d = wd;
          nodeLocalRay(sc, nextNode, nd.depth, o, d);
          r.set(o, d);
          rayIsWorld = false;
        }
        float dd;
        if (!slab<COUNT>(r, V3(nd.bmin), V3(nd.bmax), kTMin, st.hit.t, dd, cnt) || st.hit.t < dd) {
          nextNode = uint32_t(nd.skip);
          continue;
        }
        const int mi = nd.mesh;
        curNode = int(nextNode);
        nextNode++;
        if (mi < 0) continue;
        const YcMesh& mesh = sc.meshes[mi];
        if (!slab<COUNT>(r, V3(mesh.rootMin), V3(mesh.rootMax), kTMin, st.hit.t, dd, cnt)) continue;  // testBVH's root test
        meshIdx = uint32_t(mi);
        nodes = sc.bvhNodes + 4 * size_t(mesh.nodeOffset);
        tris = sc.bvhTris + 3 * size_t(mesh.triOffset);
        cur = mesh.rootRef;
        dcur = dd;
        sp = 0;
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

    // ---- inner-node steps ------------------------------------------------------------------------
    // SPEC (all non-counting builds): a lane that reaches a leaf parks it in `pend` and keeps walking;
    // it only blocks when it reaches a second leaf (or runs out of stack) with one still parked.
    // Parked leaves are tested in the order they were reached, before any later leaf, so the sequence
    // of triangle tests — and with it the accepted hit, exact ties and alpha-test draws — is the
    // reference's.  Only box culling sees a slightly older hit.t, i.e. a few extra box tests.
    for (;;) {
      if (SPEC && state == kLaneTrav && (cur & YC_REF_LEAF) && cur != kNoRef && pend == 0u) {
        pend = cur, pendD = dcur;
        if (sp == 0) cur = kNoRef;
        else stack.pop(--sp, cur, dcur);
      }
      const bool inner = state == kLaneTrav && !(cur & YC_REF_LEAF);
      const unsigned im = __ballot_sync(FULL, inner);
      if (im == 0) break;
      if (__popc(im) < tune.innerMin && __ballot_sync(FULL, state == kLaneTrav && (cur & YC_REF_LEAF))) break;
      if (inner) {
        const bool live = dcur < st.hit.t;
        bool hit1 = false, hit2 = false;
        float d1 = 0.0f, d2 = 0.0f;
        uint32_t c1 = 0, c2 = 0;
        if (live) {
          const NodeRec nr = loadNode(nodes + 4 * size_t(cur));
          hit1 = slabLive<COUNT>(r, V3(nr.f[0], nr.f[1], nr.f[2]), V3(nr.f[3], nr.f[4], nr.f[5]), kTMin, st.hit.t, d1, cnt);
          hit2 = slabLive<COUNT>(r, V3(nr.f[6], nr.f[7], nr.f[8]), V3(nr.f[9], nr.f[10], nr.f[11]), kTMin, st.hit.t, d2, cnt);
          c1 = __float_as_uint(nr.f[12]), c2 = __float_as_uint(nr.f[13]);
        }
        // testBVH's descend / push / pop decisions (ray-integrator.cpp:131-152) as selects
        const bool both = hit1 && hit2;
        const bool swapped = both && d1 > d2;
        const uint32_t farRef = swapped ? c1 : c2;
        const float farD = swapped ? d1 : d2;
        const uint32_t nearRef = (hit1 && !swapped) ? c1 : c2;
        const float nearD = (hit1 && !swapped) ? d1 : d2;
        if (both) stack.push(sp++, farRef, farD);
        if (hit1 || hit2) {
          cur = nearRef, dcur = nearD;
        } else if (sp == 0) {
          if (SPEC && pend != 0u) cur = kNoRef;  // mesh walked, one leaf still parked
          else state = kLaneNode;                // this mesh is done: on to the next scene node
        } else {
          stack.pop(--sp, cur, dcur);
        }
      }
    }

    // ---- leaf step -------------------------------------------------------------------------------
    // kNoRef carries the leaf bit, so "cur is a leaf" also covers "nothing left but a parked leaf".
    if (state == kLaneTrav && ((cur & YC_REF_LEAF) || (SPEC && pend != 0u))) {
      uint32_t leaf = cur;
      float leafD = dcur;
      if (SPEC) {
        if (pend != 0u) {
          leaf = pend, leafD = pendD;
          pend = 0u;
        } else {
          // only reachable for cur == kNoRef with nothing parked (cannot happen) or a fresh leaf
          if (cur != kNoRef) {
            if (sp == 0) cur = kNoRef;
            else stack.pop(--sp, cur, dcur);
          }
        }
      }
      if (leaf != kNoRef && leafD < st.hit.t) {
        const YcMesh& mesh = sc.meshes[meshIdx];
        uint32_t ti = leaf & ~YC_REF_LEAF;
        while (true) {
          const float4 a = __ldg(tris + 3 * size_t(ti)), b = __ldg(tris + 3 * size_t(ti) + 1),
                       c = __ldg(tris + 3 * size_t(ti) + 2);
          meshHit |= testTriangle<NEE, ALPHA, COUNT>(sc, mesh, r, a, b, c, curNode, st, &smp, cnt);
          if (NEE && meshHit) break;  // ray-integrator.cpp:121: leaves the leaf loop only
          if (__float_as_uint(c.z) & YC_TRI_LAST) break;
          ti++;
        }
        didHit |= meshHit;
      }
      if (NEE && EARLY_OUT && didHit) {
        io.store(item, st, true, smp);
        state = kLaneIdle;
        pend = 0u;
      } else if (SPEC) {
        if (cur == kNoRef) state = kLaneNode;  // stack empty and nothing parked any more
      } else if (sp == 0) {
        state = kLaneNode;
      } else {
        stack.pop(--sp, cur, dcur);
      }
    }
  }
}

}  // namespace yb
