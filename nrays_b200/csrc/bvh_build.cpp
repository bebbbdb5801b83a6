// bvh_build.cpp — binned-SAH BVH2 builder (host).  See bvh_build.h.
#include "bvh_build.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <cstdio>
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

// The top levels of a large tree run on the calling thread while the worker threads have nothing to do yet: their passes over
// the items (bounds, binning) are split into one contiguous chunk per hardware thread.
static constexpr size_t kParallelPass = 1u << 18;
template <class F>
static void chunked(size_t n, unsigned hw, F &&fn) {  // fn(thread, lo, hi)
  const size_t chunk = (n + hw - 1) / hw;
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < hw; ++t) {
    const size_t lo = (size_t)t * chunk, hi = std::min(n, lo + chunk);
    if (lo >= hi) break;
    pool.emplace_back([&fn, t, lo, hi]() { fn(t, lo, hi); });
  }
  for (auto &th : pool) th.join();
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
  // subtrees per thread: the levels between the threaded passes (>= kParallelPass items) and the task size run on THIS thread
  // alone, so few, large tasks win (C4 on 16 cores: top of the tree 423 / 379 / 329 / 279 ms at 8 / 4 / 2 / 1, subtrees 107 /
  // 116 / 126 / 160 ms)
  size_t task_div = 2;
  if (const char *e = getenv("NRB_BVH_TASK_DIV")) task_div = (size_t)std::max(1, atoi(e));
  task_size_ = std::max<size_t>(16384, items.size() / (task_div * (size_t)hw));
  auto T0 = std::chrono::steady_clock::now();
  int root = build_rec(items.data(), items.size(), false, 0, root_box);
  tasks_ = nullptr;
  auto T1 = std::chrono::steady_clock::now();
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
  auto T2 = std::chrono::steady_clock::now();
  // splice: shift node indices and leaf triangle offsets of each private pool (every pool copied by a worker thread)
  std::vector<int> resolved(tasks.size());
  std::vector<size_t> node_at(tasks.size() + 1), tri_at(tasks.size() + 1);
  node_at[0] = nodes.size(), tri_at[0] = tri_order.size();
  for (size_t i = 0; i < tasks.size(); ++i) {
    node_at[i + 1] = node_at[i] + subs[i].nodes.size();
    tri_at[i + 1] = tri_at[i] + subs[i].tri_order.size();
  }
  nodes.resize(node_at.back());
  tri_order.resize(tri_at.back());
  {
    std::atomic<size_t> next(0);
    auto worker = [&]() {
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= tasks.size()) return;
        const int node_off = (int)node_at[i];
        const uint32_t tri_off = (uint32_t)tri_at[i];
        auto shift = [&](int c) -> int {
          if (c >= 0) return c + node_off;
          uint32_t code = (uint32_t)~c;
          uint32_t first = (code >> 3) + tri_off;
          return ~(int)((first << 3) | (code & 7u));
        };
        BvhNode *dst = nodes.data() + node_at[i];
        for (size_t k = 0; k < subs[i].nodes.size(); ++k) {
          BvhNode n = subs[i].nodes[k];
          n.n3.x = shift(n.n3.x);
          n.n3.y = shift(n.n3.y);
          dst[k] = n;
        }
        std::copy(subs[i].tri_order.begin(), subs[i].tri_order.end(), tri_order.begin() + (ptrdiff_t)tri_at[i]);
        resolved[i] = shift(sub_root[i]);
        subs[i] = BvhBuilder();
      }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < std::min<unsigned>(hw, (unsigned)tasks.size()); ++t) pool.emplace_back(worker);
    for (auto &t : pool) t.join();
  }
  for (size_t i = 0; i < tasks.size(); ++i) max_depth_seen = std::max(max_depth_seen, sub_depth[i]);
  auto fix = [&](int &c) {
    if (c >= kPlaceholderBase && c != kEmpty) c = resolved[c - kPlaceholderBase];
  };
  for (auto &n : nodes) {
    fix(n.n3.x);
    fix(n.n3.y);
  }
  fix(root);
  if (getenv("NRB_BUILD_TIMES")) {
    auto T3 = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "[nrb] build:   SAH: top of the tree %.1f ms (%zu subtrees), subtrees %.1f ms, splice %.1f ms\n", ms(T0, T1), tasks.size(),
            ms(T1, T2), ms(T2, T3));
  }
  return root;
}

int BvhBuilder::build_payloads(std::vector<BuildItem> &items, Box *root_box) {
  return build_rec(items.data(), items.size(), true, 0, root_box);
}

int BvhBuilder::build_rec(BuildItem *items, size_t n, bool payload, int depth, Box *out_box, const Box *known_bounds,
                          const Box *known_cbounds) {
  if (depth > max_depth_seen) max_depth_seen = depth;
  Box bounds, cbounds;
  bounds.reset();
  cbounds.reset();
  unsigned hw = 1;
  if (tasks_ && n >= kParallelPass) {
    hw = std::max(1u, std::thread::hardware_concurrency());
    if (const char *e = getenv("NRB_BVH_THREADS")) hw = (unsigned)std::max(1, atoi(e));
  }
  auto grow_bounds = [&](size_t lo_i, size_t hi_i, Box &bd, Box &cb) {
    for (size_t i = lo_i; i < hi_i; ++i) {
      bd.grow(items[i].box);
      float c[3] = {centroid(items[i].box, 0), centroid(items[i].box, 1), centroid(items[i].box, 2)};
      cb.grow(c);
    }
  };
  if (known_bounds) {
    bounds = *known_bounds, cbounds = *known_cbounds;  // the parent's bins already hold them
  } else if (hw > 1) {
    std::vector<Box> pb(hw), pc(hw);
    for (unsigned t = 0; t < hw; ++t) pb[t].reset(), pc[t].reset();
    chunked(n, hw, [&](unsigned t, size_t lo_i, size_t hi_i) { grow_bounds(lo_i, hi_i, pb[t], pc[t]); });
    for (unsigned t = 0; t < hw; ++t)
      if (pb[t].valid()) bounds.grow(pb[t]), cbounds.grow(pc[t]);
  } else {
    grow_bounds(0, n, bounds, cbounds);
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
  // one sweep over the items fills the bins of all three axes (the items of the large nodes stream from DRAM)
  Box bb3[3][kBins];
  size_t cnt3[3][kBins];
  float lo3[3], scale3[3];
  bool use3[3];
  // large nodes also bin the centroids and keep the per-chunk counts: the children's boxes and the partition's chunk
  // offsets then come out of the bins instead of two more sweeps over the items
  const bool big = n >= kParallelPass;
  Box cb3[3][kBins];
  std::vector<size_t> chunk_cnt;  // [chunk][axis][bin] of the threaded sweep
  if (depth < kForceMedianDepth) {
    if (big)
      for (int axis = 0; axis < 3; ++axis)
        for (int b = 0; b < kBins; ++b) cb3[axis][b].reset();
    for (int axis = 0; axis < 3; ++axis) {
      const float ext = cbounds.hi[axis] - cbounds.lo[axis];
      use3[axis] = ext > 0.0f;
      lo3[axis] = cbounds.lo[axis];
      scale3[axis] = use3[axis] ? (float)kBins * (1.0f - 1e-6f) / ext : 0.0f;
      for (int b = 0; b < kBins; ++b) bb3[axis][b].reset(), cnt3[axis][b] = 0;
    }
    auto bin_items = [&](size_t lo_i, size_t hi_i, Box (*tb)[kBins], size_t (*tc)[kBins], Box (*tcb)[kBins]) {
      for (size_t i = lo_i; i < hi_i; ++i) {
        const Box &bx = items[i].box;
        const float c[3] = {centroid(bx, 0), centroid(bx, 1), centroid(bx, 2)};
        for (int axis = 0; axis < 3; ++axis) {
          if (!use3[axis]) continue;
          int b = (int)((c[axis] - lo3[axis]) * scale3[axis]);
          b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
          tb[axis][b].grow(bx);
          tc[axis][b]++;
          if (tcb) tcb[axis][b].grow(c);
        }
      }
    };
    if (hw > 1) {
      struct ThreadBins {
        Box bb[3][kBins];
        Box cb[3][kBins];
        size_t cnt[3][kBins];
      };
      std::vector<ThreadBins> tb(hw);
      for (auto &t : tb)
        for (int axis = 0; axis < 3; ++axis)
          for (int b = 0; b < kBins; ++b) t.bb[axis][b].reset(), t.cb[axis][b].reset(), t.cnt[axis][b] = 0;
      chunked(n, hw, [&](unsigned t, size_t lo_i, size_t hi_i) { bin_items(lo_i, hi_i, tb[t].bb, tb[t].cnt, big ? tb[t].cb : nullptr); });
      chunk_cnt.assign((size_t)hw * 3 * kBins, 0);
      for (unsigned ti = 0; ti < hw; ++ti) {
        const ThreadBins &t = tb[ti];
        for (int axis = 0; axis < 3; ++axis)
          for (int b = 0; b < kBins; ++b) {
            chunk_cnt[((size_t)ti * 3 + axis) * kBins + b] = t.cnt[axis][b];
            if (!t.cnt[axis][b]) continue;
            bb3[axis][b].grow(t.bb[axis][b]), cnt3[axis][b] += t.cnt[axis][b];
            if (big) cb3[axis][b].grow(t.cb[axis][b]);
          }
      }
    } else {
      bin_items(0, n, bb3, cnt3, big ? cb3 : nullptr);
    }
    for (int axis = 0; axis < 3; ++axis) {
      if (!use3[axis]) continue;
      const Box *bb = bb3[axis];
      const size_t *cnt = cnt3[axis];
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
    auto left = [&](const BuildItem &it) {
      int b = (int)((centroid(it.box, best_axis) - lo) * scale);
      b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
      return b <= best_bin;
    };
    if (n >= kParallelPass) {
      // Large node: a STABLE partition through a scratch array — counts per chunk, prefix, scatter, copy back — so the
      // result does not depend on how many threads share the work (std::partition's order would).  The size test alone
      // picks this path: a single-threaded build takes it too and produces the same tree.
      const unsigned pw = std::max(1u, hw);
      if (scratch_n_ < n) {
        scratch_.reset(new BuildItem[n]);
        scratch_n_ = n;
      }
      BuildItem *tmp = scratch_.get();
      // left items per chunk: the bins of the sweep above were filled chunk by chunk, with the same chunks
      std::vector<size_t> n_left(pw + 1, 0);
      for (unsigned t = 0; t < pw; ++t) {
        size_t c = 0;
        for (int b = 0; b <= best_bin; ++b)
          c += pw > 1 ? chunk_cnt[((size_t)t * 3 + best_axis) * kBins + b] : cnt3[best_axis][b];
        n_left[t + 1] = n_left[t] + c;  // left items in front of chunk t + 1
      }
      const size_t total_left = n_left[pw];
      auto scatter_chunk = [&](unsigned t, size_t lo_i, size_t hi_i) {
        size_t l = n_left[t], r = total_left + (lo_i - n_left[t]);
        for (size_t i = lo_i; i < hi_i; ++i) {
          if (left(items[i])) tmp[l++] = items[i];
          else tmp[r++] = items[i];
        }
      };
      auto copy_chunk = [&](unsigned, size_t lo_i, size_t hi_i) { std::memcpy(items + lo_i, tmp + lo_i, (hi_i - lo_i) * sizeof(BuildItem)); };
      if (pw > 1) {
        chunked(n, pw, scatter_chunk);
        chunked(n, pw, copy_chunk);
      } else {
        scatter_chunk(0, 0, n);
        copy_chunk(0, 0, n);
      }
      mid = total_left;
    } else {
      BuildItem *m = std::partition(items, items + n, left);
      mid = (size_t)(m - items);
    }
  } else {
    mid = 0;
  }
  bool from_bins = best_axis >= 0;
  if (mid == 0 || mid == n) {
    from_bins = false;
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
  Box kb[2], kc[2];
  const bool known = big && from_bins;
  if (known) {
    for (int side = 0; side < 2; ++side) kb[side].reset(), kc[side].reset();
    for (int b = 0; b < kBins; ++b) {
      if (!cnt3[best_axis][b]) continue;
      const int side = b <= best_bin ? 0 : 1;
      kb[side].grow(bb3[best_axis][b]), kc[side].grow(cb3[best_axis][b]);
    }
  }
  int c0 = build_rec(items, mid, payload, depth + 1, &b0, known ? &kb[0] : nullptr, known ? &kc[0] : nullptr);
  int c1 = build_rec(items + mid, n - mid, payload, depth + 1, &b1, known ? &kb[1] : nullptr, known ? &kc[1] : nullptr);
  BvhNode &nd = nodes[idx];
  set_child_box(nd, 0, b0);
  set_child_box(nd, 1, b1);
  nd.n3 = make_int4(c0, c1, 0, 0);
  return idx;
}

}  // namespace nrb
