// bvh_build.cpp — binned-SAH BVH2 builder (host).  See bvh_build.h.
#include "bvh_build.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace nrb {

namespace {
#ifndef NRB_BINS
#define NRB_BINS 16
#endif
constexpr int kBins = NRB_BINS;
float kCostNode = 1.2f;  // relative cost of one two-box node visit vs one triangle test (env NRB_BVH_CNODE)
int kLeafMax = kMaxLeafTris;  // env NRB_BVH_LEAF (<= kMaxLeafTris)
constexpr int kForceMedianDepth = 36;

inline float centroid(const Box &b, int axis) { return 0.5f * (b.lo[axis] + b.hi[axis]); }

void set_child_box(BvhNode &n, int which, const Box &b) {
  if (which == 0) {
    n.n0 = make_float4(b.lo[0], b.hi[0], b.lo[1], b.hi[1]);
    n.n2.x = b.lo[2];
    n.n2.y = b.hi[2];
  } else {
    n.n1 = make_float4(b.lo[0], b.hi[0], b.lo[1], b.hi[1]);
    n.n2.z = b.lo[2];
    n.n2.w = b.hi[2];
  }
}
}  // namespace

void pad_box(Box &b, float scene_extent) {
  for (int i = 0; i < 3; ++i) {
    float a = std::max(std::fabs(b.lo[i]), std::fabs(b.hi[i]));
    float eps = a * 4.76837158e-7f /* 2^-21 */ + scene_extent * 1e-7f + 1e-30f;
    b.lo[i] -= eps;
    b.hi[i] += eps;
  }
}

// Placeholder child codes for deferred subtrees: inner-node indices never get this large.
static constexpr int kPlaceholderBase = 0x40000000;

int BvhBuilder::build_triangles(std::vector<BuildItem> &items, Box *root_box) {
  if (const char *e = getenv("NRB_BVH_CNODE")) kCostNode = (float)atof(e);
  if (const char *e = getenv("NRB_BVH_LEAF")) kLeafMax = std::max(1, std::min(kMaxLeafTris, atoi(e)));
  unsigned hw = std::thread::hardware_concurrency();
  if (const char *e = getenv("NRB_BVH_THREADS")) hw = (unsigned)std::max(1, atoi(e));
  if (items.size() < 200000 || hw < 2) return build_rec(items.data(), items.size(), false, 0, root_box);

  // Large mesh: split the top of the tree here, build the subtrees (<= n / (8 * threads) triangles each)
  // on worker threads into private pools, then splice them into the shared pool.
  std::vector<Task> tasks;
  tasks_ = &tasks;
  task_size_ = std::max<size_t>(16384, items.size() / (8 * (size_t)hw));
  int root = build_rec(items.data(), items.size(), false, 0, root_box);
  tasks_ = nullptr;
  std::vector<BvhBuilder> subs(tasks.size());
  std::vector<int> sub_root(tasks.size());
  std::vector<int> sub_depth(tasks.size(), 0);
  {
    std::atomic<size_t> next(0);
    auto worker = [&]() {
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= tasks.size()) return;
        Box b;
        sub_root[i] = subs[i].build_rec(tasks[i].items, tasks[i].n, false, tasks[i].depth, &b);
        sub_depth[i] = subs[i].max_depth_seen;
      }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < std::min<unsigned>(hw, (unsigned)tasks.size()); ++t) pool.emplace_back(worker);
    for (auto &t : pool) t.join();
  }
  // splice: shift node indices and leaf triangle offsets of each private pool
  std::vector<int> resolved(tasks.size());
  for (size_t i = 0; i < tasks.size(); ++i) {
    const int node_off = (int)nodes.size();
    const uint32_t tri_off = (uint32_t)tri_order.size();
    auto shift = [&](int c) -> int {
      if (c >= 0) return c + node_off;
      uint32_t code = (uint32_t)~c;
      uint32_t first = (code >> 3) + tri_off;
      return ~(int)((first << 3) | (code & 7u));
    };
    for (BvhNode n : subs[i].nodes) {
      n.n3.x = shift(n.n3.x);
      n.n3.y = shift(n.n3.y);
      nodes.push_back(n);
    }
    tri_order.insert(tri_order.end(), subs[i].tri_order.begin(), subs[i].tri_order.end());
    resolved[i] = shift(sub_root[i]);
    max_depth_seen = std::max(max_depth_seen, sub_depth[i]);
    subs[i] = BvhBuilder();
  }
  auto fix = [&](int &c) {
    if (c >= kPlaceholderBase && c != kEmpty) c = resolved[c - kPlaceholderBase];
  };
  for (auto &n : nodes) {
    fix(n.n3.x);
    fix(n.n3.y);
  }
  fix(root);
  return root;
}

int BvhBuilder::build_payloads(std::vector<BuildItem> &items, Box *root_box) {
  return build_rec(items.data(), items.size(), true, 0, root_box);
}

int BvhBuilder::build_rec(BuildItem *items, size_t n, bool payload, int depth, Box *out_box) {
  if (depth > max_depth_seen) max_depth_seen = depth;
  Box bounds, cbounds;
  bounds.reset();
  cbounds.reset();
  for (size_t i = 0; i < n; ++i) {
    bounds.grow(items[i].box);
    float c[3] = {centroid(items[i].box, 0), centroid(items[i].box, 1), centroid(items[i].box, 2)};
    cbounds.grow(c);
  }
  *out_box = bounds;

  auto make_leaf_node = [&]() -> int {
    if (payload) return items[0].payload;  // n == 1
    uint32_t first = (uint32_t)tri_order.size();
    for (size_t i = 0; i < n; ++i) tri_order.push_back((uint32_t)items[i].payload);
    return make_leaf(first, (uint32_t)n, false);
  };

  size_t max_leaf = payload ? 1 : (size_t)kLeafMax;
  if (n == 1) return make_leaf_node();
  if (tasks_ && n <= task_size_ && depth > 0) {
    // small enough: build this subtree on a worker thread
    int ph = kPlaceholderBase + (int)tasks_->size();
    tasks_->push_back(Task{items, n, depth, ph});
    return ph;
  }

  // ---- choose a split ---------------------------------------------------------------------
  int best_axis = -1, best_bin = -1;
  float best_cost = 3.402823466e+38f;
  if (depth < kForceMedianDepth) {
    for (int axis = 0; axis < 3; ++axis) {
      float lo = cbounds.lo[axis], ext = cbounds.hi[axis] - cbounds.lo[axis];
      if (!(ext > 0.0f)) continue;
      Box bb[kBins];
      size_t cnt[kBins];
      for (int b = 0; b < kBins; ++b) bb[b].reset(), cnt[b] = 0;
      float scale = (float)kBins * (1.0f - 1e-6f) / ext;
      for (size_t i = 0; i < n; ++i) {
        int b = (int)((centroid(items[i].box, axis) - lo) * scale);
        b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
        bb[b].grow(items[i].box);
        cnt[b]++;
      }
      float right_area[kBins];
      size_t right_cnt[kBins];
      Box acc;
      acc.reset();
      size_t c = 0;
      for (int b = kBins - 1; b > 0; --b) {
        if (cnt[b]) acc.grow(bb[b]);
        c += cnt[b];
        right_area[b] = c ? acc.half_area() : 0.0f;
        right_cnt[b] = c;
      }
      acc.reset();
      c = 0;
      for (int b = 0; b < kBins - 1; ++b) {
        if (cnt[b]) acc.grow(bb[b]);
        c += cnt[b];
        if (c == 0 || right_cnt[b + 1] == 0) continue;
        float cost = acc.half_area() * (float)c + right_area[b + 1] * (float)right_cnt[b + 1];
        if (cost < best_cost) best_cost = cost, best_axis = axis, best_bin = b;
      }
    }
  }
  if (n <= max_leaf && !payload) {
    float leaf_cost = (float)n * bounds.half_area();
    float split_cost = best_axis >= 0 ? kCostNode * bounds.half_area() + best_cost : 3.402823466e+38f;
    if (n <= 2 || leaf_cost <= split_cost) return make_leaf_node();
  }

  size_t mid;
  if (best_axis >= 0) {
    float lo = cbounds.lo[best_axis], ext = cbounds.hi[best_axis] - cbounds.lo[best_axis];
    float scale = (float)kBins * (1.0f - 1e-6f) / ext;
    BuildItem *m = std::partition(items, items + n, [&](const BuildItem &it) {
      int b = (int)((centroid(it.box, best_axis) - lo) * scale);
      b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
      return b <= best_bin;
    });
    mid = (size_t)(m - items);
  } else {
    mid = 0;
  }
  if (mid == 0 || mid == n) {
    // degenerate (coincident centroids) or forced: balanced median split on the widest centroid axis
    int axis = 0;
    float e0 = cbounds.hi[0] - cbounds.lo[0], e1 = cbounds.hi[1] - cbounds.lo[1], e2 = cbounds.hi[2] - cbounds.lo[2];
    if (e1 > e0 && e1 >= e2) axis = 1;
    if (e2 > e0 && e2 > e1) axis = 2;
    mid = n / 2;
    std::nth_element(items, items + mid, items + n, [&](const BuildItem &a, const BuildItem &b) {
      return centroid(a.box, axis) < centroid(b.box, axis);
    });
  }

  int idx = (int)nodes.size();
  nodes.emplace_back();
  std::memset(&nodes[idx], 0, sizeof(BvhNode));
  Box b0, b1;
  int c0 = build_rec(items, mid, payload, depth + 1, &b0);
  int c1 = build_rec(items + mid, n - mid, payload, depth + 1, &b1);
  BvhNode &nd = nodes[idx];
  set_child_box(nd, 0, b0);
  set_child_box(nd, 1, b1);
  nd.n3 = make_int4(c0, c1, 0, 0);
  return idx;
}

}  // namespace nrb
