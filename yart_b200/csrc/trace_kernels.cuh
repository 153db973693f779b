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
//   * inner-node steps run in a tight loop while enough lanes sit on inner nodes; a lane that reaches a
//     leaf parks it and keeps walking (its own test order is unchanged: parked leaves are tested in the
//     order they were reached, before any later leaf), and all leaves are processed together, so the two
//     code paths do not serialise against each other on every iteration;
//   * the inner step is branch-free apart from its `inner` guard: packed (ref, d) stack entries in shared
//     memory ([entry][thread], conflict-free, one LDS.64 / STS.64 per pop / push), the stack pointer is
//     the shared byte address, a sentinel at the bottom replaces the "empty" test, one predicated pop
//     site per iteration, descend / push as selects; trees deeper than kShStack spill to global memory
//     through an out-of-line path; no local memory is used.
// Node records are fetched as two 256-bit read-only loads, triangles as three 128-bit loads.
#pragma once
#include "integrator.cuh"

namespace yb {

constexpr int kTraceBlock = 128;       // 4 warps per CTA

// Scheduling knobs (they change only how lanes are grouped, never a ray's own test sequence).
struct TraceTuning {
  int refillMin;  // refill the warp from the queue when at least this many lanes are idle
  int innerMin;   // keep stepping inner nodes while at least this many lanes sit on one
  int shEntries;  // shared-memory stack entries in use (2..kPsStack; the rest spills): tests shrink it to exercise the spill path
};

// Per-lane traversal stack of the persistent kernels: packed (ref, d) entries, [entry][thread] in shared
// memory (one 64-bit access per push / pop, conflict-free), global spill area beyond kPsStack entries.
// The stack pointer is the shared-memory byte address of the next free slot, so push / pop are one
// add and one LDS.64 / STS.64; entry 0 holds a sentinel (kNoRef) written when a mesh is entered, so a
// pop never tests for "empty": popping the sentinel yields kNoRef = "this mesh is walked".
constexpr int kPsStack = kShStack + 1;
constexpr uint32_t kPsStride = kTraceBlock * 8;  // bytes between consecutive entries of one thread
constexpr int kSpillEntries = kMaxStack;         // global spill entries per thread (covers the smallest shared stack)
struct WarpStack {
  uint32_t base;   // shared address of this thread's entry 0
  uint32_t limit;  // base + kPsStack * kPsStride: first address that lives in the spill area
  uint2** spillBase;  // in shared memory: the spill area's base pointer (only the rare path reads it)
  __device__ __forceinline__ void init(uint2* shStack, uint2** shSpill, uint2* spill, int shEntries) {
    base = uint32_t(__cvta_generic_to_shared(shStack + threadIdx.x));  // rarely needed: may be recomputed
    const uint32_t l = base + uint32_t(shEntries) * kPsStride;
    asm volatile("mov.u32 %0, %1;" : "=r"(limit) : "r"(l));  // opaque: compared at every push / pop, kept in a register
    if (threadIdx.x == 0) *shSpill = spill;
    __syncthreads();
    spillBase = shSpill;
  }
  // rare path (trees deeper than kShStack): everything is derived here so nothing of it stays live
  __device__ __noinline__ static uint2* spillSlot(uint2** spillBase, uint32_t off) {
    return *spillBase + size_t(blockIdx.x * kTraceBlock + threadIdx.x) * kSpillEntries + off / kPsStride;
  }
  __device__ __forceinline__ void put(uint32_t sa, uint32_t ref, float d) const {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(sa), "r"(ref), "r"(__float_as_uint(d)));
  }
  // push / pop under a lane predicate: the shared-memory access is a single predicated instruction and
  // only the (rare) spill case branches.
  __device__ __forceinline__ void pushIf(bool on, uint32_t& sa, uint32_t ref, float d) const {
    if (on && sa < limit) put(sa, ref, d);
    if (on && sa >= limit) {
      if (sa < base + uint32_t(kMaxStack + 1) * kPsStride) *spillSlot(spillBase, sa - limit) = make_uint2(ref, __float_as_uint(d));
    }
    if (on) sa += kPsStride;
  }
  __device__ __forceinline__ void popIf(bool on, uint32_t& sa, uint32_t& ref, float& d) const {
    if (on) sa -= kPsStride;
    uint32_t x = ref, y = __float_as_uint(d);
    if (on && sa < limit) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(sa));
    if (on && sa >= limit) {
      const uint2 v = *spillSlot(spillBase, sa - limit);
      x = v.x, y = v.y;
    }
    ref = x;
    d = __uint_as_float(y);
  }
  __device__ __forceinline__ void pop(uint32_t& sa, uint32_t& ref, float& d) const { popIf(true, sa, ref, d); }
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
  __shared__ uint2 shStack[kPsStack * kTraceBlock];
  __shared__ uint2* shSpill;
  WarpStack stack;
  stack.init(shStack, &shSpill, spillBase, tune.shEntries);
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u, ltMask = (1u << lane) - 1u;

  int state = kLaneIdle;
  bool exhausted = false;
  uint32_t item = 0, nextNode = 0, cur = 0;
  uint32_t sp = 0;  // stack pointer: shared byte address of the next free slot (WarpStack)
  int curNode = 0;
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
  constexpr uint32_t kNoRef = 0xffffffffu;   // "no current node" (has the leaf bit: never stepped as inner)
  constexpr uint32_t kPopRef = 0xfffffffeu;  // "pop at the top of the next inner step" (leaf bit set as well)
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

    // ---- inner-node steps ------------------------------------------------------------------------
    // SPEC (all non-counting builds): a lane that reaches a leaf parks it in `pend` and keeps walking;
    // it only blocks when it reaches a second leaf (or has walked the mesh) with one still parked.
    // Parked leaves are tested in the order they were reached, before any later leaf, so the sequence
    // of triangle tests — and with it the accepted hit, exact ties and alpha-test draws — is the
    // reference's.  Only box culling sees a slightly older hit.t, i.e. a few extra box tests.
    // The loop body is branch-free apart from the `inner` guard: descend / push / pop are selects and
    // predicated 64-bit shared-memory accesses, and the stack sentinel stands in for the "empty" test.
    for (;;) {
      const bool trav = state == kLaneTrav;
      bool doPop = trav && cur == kPopRef;  // left by the previous step: one pop site per iteration
      if (SPEC && trav && pend == 0u && int32_t(cur) < -2) {  // a leaf (bit 31) other than kNoRef / kPopRef
        pend = cur, pendD = dcur;
        doPop = true;
      }
      stack.popIf(doPop, sp, cur, dcur);  // the sentinel ends the mesh: cur = kNoRef
      const bool inner = trav && int32_t(cur) >= 0;
      const unsigned im = __ballot_sync(FULL, inner);
      if (im == 0) break;
      if (__popc(im) < tune.innerMin && __ballot_sync(FULL, trav && !inner)) break;
      if (inner) {
        const NodeRec nr = loadNode(nodes + 4 * size_t(cur));
        // testBVH's pop-time cull `d < hit.t` (ray-integrator.cpp:100): an entry that fails it is
        // dropped without testing its children — here both tests are made to fail (tmax = -inf)
        const bool live = dcur < st.hit.t;
        const float tmx = live ? st.hit.t : -INFINITY;
        float d1, d2;
        const bool hit1 = slabLive(r, V3(nr.f[0], nr.f[1], nr.f[2]), V3(nr.f[3], nr.f[4], nr.f[5]), kTMin, tmx, d1);
        const bool hit2 = slabLive(r, V3(nr.f[6], nr.f[7], nr.f[8]), V3(nr.f[9], nr.f[10], nr.f[11]), kTMin, tmx, d2);
        if (COUNT && live) cnt.box += 2;
        const uint32_t c1 = __float_as_uint(nr.f[12]), c2 = __float_as_uint(nr.f[13]);
        // testBVH's descend / push / pop decisions (ray-integrator.cpp:131-152) as selects
        const bool both = hit1 && hit2;
        const bool swapped = both && d1 > d2;
        const bool first = hit1 && !swapped;
        const uint32_t farRef = swapped ? c1 : c2;
        const float farD = swapped ? d1 : d2;
        const uint32_t nearRef = first ? c1 : c2;
        const float nearD = first ? d1 : d2;
        stack.pushIf(both, sp, farRef, farD);
        cur = (hit1 || hit2) ? nearRef : kPopRef;
        dcur = nearD;
      }
    }

    // ---- leaf step -------------------------------------------------------------------------------
    // kNoRef carries the leaf bit, so "cur is a leaf" also covers "mesh walked" (with or without a
    // parked leaf); a lane still on an inner node uses the break to test its parked leaf.
    if (state == kLaneTrav && (int32_t(cur) < 0 || (SPEC && pend != 0u))) {
      uint32_t leaf = cur;
      float leafD = dcur;
      if (SPEC) {
        // a real leaf in `cur` with nothing parked cannot leave the loop above (it is parked at once)
        leaf = pend, leafD = pendD;
        pend = 0u;
      }
      if (int32_t(leaf) < -1 && leafD < st.hit.t) {
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
      } else {
        if (!SPEC && cur != kNoRef) stack.pop(sp, cur, dcur);
        if (cur == kNoRef) state = kLaneNode;  // mesh walked and nothing parked any more
      }
    }
  }
}

}  // namespace yb
