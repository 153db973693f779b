// bvh_build.hpp — host SAH BVH builder that reproduces the reference's tree exactly.
//
// Restates yart's binned-SAH builder (reference src/core/bvh.hpp:41-67 init, :101-113
// updateBounds, :121-133 getCentroidBounds, :140-184 subdivide, :273-347 SahBVH::getSplit) so
// that the GPU traverses the SAME tree (same boxes, same left/right order, same triangle order
// inside leaves) and therefore tests candidates in the reference's order.  Differences are
// purely mechanical: triangle bounds/centroids are computed once instead of per visit, and
// independent subtrees are built on worker threads; node numbering is re-derived afterwards
// in the reference's allocation order (children adjacent, left subtree numbered first).
#pragma once
#include <atomic>
#include <cstdint>
#include <memory>
#include <thread>
#include <vector>

#include "hmath.hpp"

namespace yartb {

// Reference node layout (src/core/bvh.hpp:21-33), used for parity tests and as the builder's output.
struct RefBvhNode {
  float mn[3], mx[3];
  uint32_t leftFirst;
  uint32_t span;  // 0 → inner
};

struct BvhBuildResult {
  std::vector<RefBvhNode> nodes;   // reference numbering
  std::vector<uint32_t> indices;   // BVH::m_indices
};

class SahBvhBuilder {
 public:
  static constexpr uint32_t kMaxLeafSize = 20;  // MAX_LEAF_SIZE, bvh.hpp:14
  static constexpr uint32_t kBins = 20;         // nBins, bvh.hpp:283
  // Mesh::BVHType (mesh.hpp:17): SahBVH is what the reference instantiates; MedianSplitBVH (bvh.hpp:237-264)
  // is its baseline alternative — same init / subdivide, a different getSplit.
  enum Kind : uint32_t { kSah = 0, kMedianSplit = 1 };
  Kind kind = kSah;

  BvhBuildResult build(const float* positions, size_t nVerts, const uint32_t* faces /*stride 4*/, size_t nTris,
                       unsigned threads = 0) {
    nTris_ = nTris;
    triBounds_.resize(nTris);
    centroids_.resize(nTris);
    for (size_t i = 0; i < nTris; i++) {
      f3 v0(positions + 3 * size_t(faces[4 * i + 0])), v1(positions + 3 * size_t(faces[4 * i + 1])),
        v2(positions + 3 * size_t(faces[4 * i + 2]));
      triBounds_[i] = Bounds3::fromTriangle(v0, v1, v2);  // bvh.hpp:105-109
      centroids_[i] = (v0 + v1 + v2) / 3.0f;              // primitives.hpp:46
    }
    BvhBuildResult out;
    out.indices.resize(nTris);
    for (size_t i = 0; i < nTris; i++) out.indices[i] = uint32_t(i);
    idx_ = out.indices.data();
    if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
    budget_.store(int(threads) - 1);

    auto root = std::make_unique<TNode>();
    root->first = 0;
    root->span = uint32_t(nTris);
    root->bounds = rangeBounds(0, uint32_t(nTris));
    subdivide(root.get());

    // Number nodes in the reference's allocation order (bvh.hpp:165-166, 180-183).
    out.nodes.reserve(2 * nTris);
    out.nodes.resize(1);
    number(root.get(), 0, out.nodes);
    return out;
  }

 private:
  struct TNode {
    Bounds3 bounds;
    uint32_t first = 0, span = 0;
    std::unique_ptr<TNode> left, right;
  };
  struct Bin {
    uint32_t count = 0;
    Bounds3 bounds;
  };

  size_t nTris_ = 0;
  std::vector<Bounds3> triBounds_;
  std::vector<f3> centroids_;
  uint32_t* idx_ = nullptr;
  std::atomic<int> budget_{0};

  // updateBounds: min/max fold over padded triangle boxes (order-independent, no rounding)
  Bounds3 rangeBounds(uint32_t first, uint32_t span) const {
    Bounds3 b;
    for (uint32_t i = 0; i < span; i++) b = Bounds3::join(b, triBounds_[idx_[first + i]]);
    return b;
  }

  // uint32_t(float) as x86-64 g++ evaluates it (cvttss2si to 64 bit, low word): NaN → 0
  static uint32_t toU32(float x) {
    if (x != x) return 0u;
    return uint32_t(int64_t(x));
  }

  // MedianSplitBVH::getSplit, bvh.hpp:239-263: longest axis of the centroid bounds, split in the middle
  bool getSplitMedian(const TNode& node, uint8_t& axis, float& splitPos) const {
    if (node.span <= 2) return false;
    Bounds3 cb;
    for (uint32_t i = node.first; i < node.first + node.span; i++) cb.expandToInclude(centroids_[idx_[i]]);
    axis = 0;
    const f3 size = cb.mx - cb.mn;
    if (size[1] > size[0]) axis = 1;
    if (size[2] > size[axis]) axis = 2;
    splitPos = cb.mn[axis] + size[axis] * 0.5f;
    return true;
  }

  bool getSplit(const TNode& node, uint8_t& axis, float& splitPos) const {
    if (kind == kMedianSplit) return getSplitMedian(node, axis, splitPos);
    float minCost = std::numeric_limits<float>::infinity();
    Bounds3 cb;
    for (uint32_t i = node.first; i < node.first + node.span; i++) cb.expandToInclude(centroids_[idx_[i]]);
    constexpr uint32_t nSplits = kBins - 1;
    for (uint8_t a = 0; a < 3; a++) {
      float bmin = cb.mn[a], bsize = (cb.mx - cb.mn)[a];
      Bin bins[kBins];
      float scale = float(kBins) / bsize;
      for (uint32_t i = 0; i < node.span; i++) {
        uint32_t t = idx_[node.first + i];
        uint32_t b = std::min(kBins - 1, toU32(scale * (centroids_[t][a] - bmin)));
        bins[b].count++;
        bins[b].bounds = Bounds3::join(bins[b].bounds, triBounds_[t]);
      }
      float costs[nSplits] = {0.0f};
      uint32_t countBelow = 0;
      Bounds3 below;
      for (uint32_t i = 0; i < nSplits; i++) {
        below = Bounds3::join(below, bins[i].bounds);
        countBelow += bins[i].count;
        costs[i] += float(countBelow) * below.area();  // empty box → 0 * inf = NaN, never the minimum
      }
      uint32_t countAbove = 0;
      Bounds3 above;
      for (uint32_t i = nSplits; i > 0; i--) {
        above = Bounds3::join(above, bins[i].bounds);
        countAbove += bins[i].count;
        costs[i - 1] += float(countAbove) * above.area();
      }
      for (uint32_t i = 0; i < nSplits; i++) {
        if (costs[i] < minCost) {
          minCost = costs[i];
          axis = a;
          splitPos = bmin + bsize * (float(i + 1) / float(kBins));
        }
      }
    }
    float leafCost = (float(node.span) - 0.5f) * node.bounds.area();
    if (node.span <= kMaxLeafSize && leafCost < minCost) return false;
    return true;
  }

  void subdivide(TNode* node) {
    uint8_t axis = 0;
    float splitPos = 0;
    if (!getSplit(*node, axis, splitPos)) return;
    int64_t i = node->first;
    int64_t j = i + int64_t(node->span) - 1;
    while (i <= j) {
      float c = centroids_[idx_[i]][axis];
      if (c < splitPos) i++;
      else std::swap(idx_[i], idx_[j--]);
    }
    size_t leftCount = size_t(i - node->first);
    if (leftCount == 0 || leftCount == node->span) return;
    node->left = std::make_unique<TNode>();
    node->right = std::make_unique<TNode>();
    node->left->first = node->first;
    node->left->span = uint32_t(leftCount);
    node->right->first = uint32_t(i);
    node->right->span = node->span - uint32_t(leftCount);
    node->left->bounds = rangeBounds(node->left->first, node->left->span);
    node->right->bounds = rangeBounds(node->right->first, node->right->span);

    // Subtrees touch disjoint index ranges → build the left one on another thread when worthwhile.
    bool forked = false;
    std::thread worker;
    if (node->left->span > 20000 && node->right->span > 20000) {
      int b = budget_.fetch_sub(1);
      if (b > 0) {
        forked = true;
        TNode* l = node->left.get();
        worker = std::thread([this, l] { subdivide(l); });
      } else {
        budget_.fetch_add(1);
      }
    }
    if (!forked) subdivide(node->left.get());
    subdivide(node->right.get());
    if (forked) {
      worker.join();
      budget_.fetch_add(1);
    }
  }

  void number(const TNode* t, uint32_t self, std::vector<RefBvhNode>& out) const {
    RefBvhNode& n = out[self];
    for (int k = 0; k < 3; k++) {
      n.mn[k] = t->bounds.mn[k];
      n.mx[k] = t->bounds.mx[k];
    }
    if (!t->left) {
      n.leftFirst = t->first;
      n.span = t->span;
      return;
    }
    uint32_t l = uint32_t(out.size());
    out.resize(out.size() + 2);
    out[self].leftFirst = l;  // `n` may dangle after resize
    out[self].span = 0;
    number(t->left.get(), l, out);
    number(t->right.get(), l + 1, out);
  }
};

}  // namespace yartb
