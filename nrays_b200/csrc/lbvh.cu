// lbvh.cu — GPU BVH build (replaces BVT::new_balanced, src/scene.rs:126, and the inner
// BVT of TriMesh::new, examples/loader3d.rs:695, on the device).
//
// Linear BVH after Karras, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees"
// (HPG 2012): 63-bit Morton codes of the (padded) triangle boxes' centres -> radix sort -> one thread per
// internal node finds its range and split by longest-common-prefix searches -> bottom-up box fit with one
// atomic flag per node.  The binary radix tree is then collapsed into this library's layout: every subtree
// with <= 4 triangles (a contiguous range of the sorted order) becomes one leaf, the remaining internal nodes
// are compacted (exclusive scan) and written as 64-byte two-box nodes.
//
// Tree shape never changes render results, only traversal cost: the LBVH builds ~100x faster
// than the host binned-SAH build and traverses slower, so SAH stays the default (NrbBuildOptions.builder).
// The sort is cub::DeviceRadixSort (a library call, outside the render hot path).
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include <cub/cub.cuh>

#include "lbvh.h"

namespace nrb {

namespace {

#define LB_CU(call)                    \
  do {                                 \
    cudaError_t e__ = (call);          \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

struct DBox {
  float lo[3], hi[3];
};

__device__ __forceinline__ unsigned long long expand21(unsigned long long v) {
  v &= 0x1FFFFFull;
  v = (v | (v << 32)) & 0x1F00000000FFFFull;
  v = (v | (v << 16)) & 0x1F0000FF0000FFull;
  v = (v | (v << 8)) & 0x100F00F00F00F00Full;
  v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}

__global__ void k_morton(const DBox *boxes, uint32_t n, DBox scene, unsigned long long *keys, uint32_t *vals) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  DBox b = boxes[i];
  unsigned long long code = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float ext = scene.hi[a] - scene.lo[a];
    float c = 0.5f * (b.lo[a] + b.hi[a]);
    float t = ext > 0.0f ? (c - scene.lo[a]) / ext : 0.0f;
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    unsigned long long q = (unsigned long long)(t * 2097151.0f);
    code |= expand21(q) << (2 - a);
  }
  keys[i] = code;
  vals[i] = i;
}

// longest common prefix of the keys at sorted positions i and j (ties broken by the positions themselves)
__device__ __forceinline__ int delta(const unsigned long long *keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  unsigned long long a = keys[i], b = keys[j];
  if (a != b) return __clzll(a ^ b);
  return 64 + __clz((unsigned)i ^ (unsigned)j);
}

// Karras 2012, Algorithm "binary radix tree": internal node i in [0, n-2]
__global__ void k_tree(const unsigned long long *keys, int n, int *left, int *right, int *range_lo, int *range_hi, int *parent_int,
                       int *parent_leaf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = delta(keys, n, i, j);
  int s = 0;
  int t = l;
  do {
    t = (t + 1) >> 1;
    if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  int gamma = i + s * d + min(d, 0);
  int lo = min(i, j), hi = max(i, j);
  // child codes: >= 0 internal node, < 0 leaf ~position
  int lc = (lo == gamma) ? ~gamma : gamma;
  int rc = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
  left[i] = lc, right[i] = rc;
  range_lo[i] = lo, range_hi[i] = hi;
  if (lc >= 0) parent_int[lc] = i; else parent_leaf[~lc] = i;
  if (rc >= 0) parent_int[rc] = i; else parent_leaf[~rc] = i;
  if (i == 0) parent_int[0] = -1;
}

// node boxes / heights are produced by other SMs during the same kernel: read them through L2 (ld.cg), a
// neighbouring entry of the same 128-byte line may sit stale in this SM's L1
__device__ __forceinline__ DBox load_box_cg(const DBox *p) {
  DBox b;
  const float *f = reinterpret_cast<const float *>(p);
#pragma unroll
  for (int k = 0; k < 3; ++k) b.lo[k] = __ldcg(f + k), b.hi[k] = __ldcg(f + 3 + k);
  return b;
}

// bottom-up: the second thread to reach a node merges its children's boxes; `height` = levels of EMITTED
// nodes below (subtrees of <= 4 triangles collapse into a leaf and count 0)
__global__ void k_fit(const DBox *boxes, const uint32_t *vals, int n, const int *left, const int *right, const int *range_lo,
                      const int *range_hi, const int *parent_int, const int *parent_leaf, DBox *node_box, int *height, int *flags) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int node = parent_leaf[j];
  while (node >= 0) {
    if (atomicAdd(&flags[node], 1) == 0) return;  // first arrival: the sibling subtree is not ready yet
    __threadfence();
    int lc = left[node], rc = right[node];
    DBox a = lc >= 0 ? load_box_cg(node_box + lc) : boxes[vals[~lc]];
    DBox b = rc >= 0 ? load_box_cg(node_box + rc) : boxes[vals[~rc]];
    DBox m;
#pragma unroll
    for (int k = 0; k < 3; ++k) m.lo[k] = fminf(a.lo[k], b.lo[k]), m.hi[k] = fmaxf(a.hi[k], b.hi[k]);
    node_box[node] = m;
    int size = range_hi[node] - range_lo[node] + 1;
    int hl = lc >= 0 ? __ldcg(height + lc) : 0, hr = rc >= 0 ? __ldcg(height + rc) : 0;
    height[node] = size <= kMaxLeafTris ? 0 : 1 + max(hl, hr);
    __threadfence();
    node = parent_int[node];
  }
}

__global__ void k_flag(const int *range_lo, const int *range_hi, int n, int *emit) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  emit[i] = (range_hi[i] - range_lo[i] + 1) > kMaxLeafTris ? 1 : 0;
}

__device__ __forceinline__ int child_code(int c, const int *range_lo, const int *range_hi, const int *new_index) {
  if (c < 0) return make_leaf((uint32_t)~c, 1, false);
  int size = range_hi[c] - range_lo[c] + 1;
  if (size <= kMaxLeafTris) return make_leaf((uint32_t)range_lo[c], (uint32_t)size, false);
  return new_index[c];
}

__global__ void k_emit(const DBox *boxes, const uint32_t *vals, int n, const int *left, const int *right, const int *range_lo,
                       const int *range_hi, const DBox *node_box, const int *emit, const int *new_index, BvhNode *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1 || !emit[i]) return;
  int lc = left[i], rc = right[i];
  DBox a = lc >= 0 ? node_box[lc] : boxes[vals[~lc]];
  DBox b = rc >= 0 ? node_box[rc] : boxes[vals[~rc]];
  BvhNode nd;
  nd.n0 = make_float4(a.lo[0], a.hi[0], a.lo[1], a.hi[1]);
  nd.n1 = make_float4(b.lo[0], b.hi[0], b.lo[1], b.hi[1]);
  nd.n2 = make_float4(a.lo[2], a.hi[2], b.lo[2], b.hi[2]);
  nd.n3 = make_int4(child_code(lc, range_lo, range_hi, new_index), child_code(rc, range_lo, range_hi, new_index), 0, 0);
  out[new_index[i]] = nd;
}

struct BuildClock {  // NRB_BUILD_TIMES=1: host-side phases of the device builders on stderr
  bool on = getenv("NRB_BUILD_TIMES") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void lap(const char *what) {
    if (!on) return;
    cudaDeviceSynchronize();
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[nrb] build:   device builder: %-22s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

// Device scratch for one build: ONE allocation carved into aligned pieces.  (Two dozen separate cudaMalloc / cudaFree calls
// cost 50-80 ms on this part — several times the build kernels themselves.)  Usage: reserve() every piece, commit(), then take().
struct Scratch {
  struct Piece {
    void **slot;
    size_t bytes;
  };
  std::vector<Piece> pieces;
  char *base = nullptr;
  ~Scratch() {
    release_extra();
    if (base) cudaFree(base);
  }
  template <class T>
  void reserve(T **p, size_t n) {
    *p = nullptr;
    pieces.push_back(Piece{reinterpret_cast<void **>(p), (std::max<size_t>(n, 1) * sizeof(T) + 255) & ~(size_t)255});
  }
  cudaError_t commit() {
    size_t total = 0;
    for (const Piece &pc : pieces) total += pc.bytes;
    cudaError_t e = cudaMalloc((void **)&base, total);
    if (e != cudaSuccess) return e;
    size_t off = 0;
    for (const Piece &pc : pieces) {
      *pc.slot = base + off;
      off += pc.bytes;
    }
    return cudaSuccess;
  }
  // late, data-dependent pieces (output nodes): their own allocation, freed with the rest
  std::vector<void *> extra;
  template <class T>
  cudaError_t alloc(T **p, size_t n) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) extra.push_back(q);
    *p = reinterpret_cast<T *>(q);
    return e;
  }
  void release_extra() {
    for (void *q : extra) cudaFree(q);
    extra.clear();
  }
};

}  // namespace

cudaError_t lbvh_build(const Box *h_boxes, uint32_t n, std::vector<BvhNode> &nodes_out, std::vector<uint32_t> &order_out,
                       int *root_code, Box *root_box, int *depth, float *gpu_ms) {
  nodes_out.clear();
  order_out.resize(n);
  Box rb;
  rb.reset();
  for (uint32_t i = 0; i < n; ++i) rb.grow(h_boxes[i]);
  *root_box = rb;
  *depth = 0;
  if (gpu_ms) *gpu_ms = 0.0f;
  if (n <= (uint32_t)kMaxLeafTris) {
    for (uint32_t i = 0; i < n; ++i) order_out[i] = i;
    *root_code = make_leaf(0, n, false);
    return cudaSuccess;
  }
  static_assert(sizeof(Box) == sizeof(DBox), "Box layout");
  Scratch sc;
  DBox *d_boxes, *d_node_box;
  unsigned long long *d_keys, *d_keys2;
  uint32_t *d_vals, *d_vals2;
  int *d_left, *d_right, *d_lo, *d_hi, *d_pi, *d_pl, *d_height, *d_flags, *d_emit, *d_new;
  BvhNode *d_out;
  sc.reserve(&d_boxes, n);
  sc.reserve(&d_node_box, n);
  sc.reserve(&d_keys, n);
  sc.reserve(&d_keys2, n);
  sc.reserve(&d_vals, n);
  sc.reserve(&d_vals2, n);
  sc.reserve(&d_left, n);
  sc.reserve(&d_right, n);
  sc.reserve(&d_lo, n);
  sc.reserve(&d_hi, n);
  sc.reserve(&d_pi, n);
  sc.reserve(&d_pl, n);
  sc.reserve(&d_height, n);
  sc.reserve(&d_flags, n);
  sc.reserve(&d_emit, n);
  sc.reserve(&d_new, n);
  size_t sort_bytes = 0, scan_bytes = 0;
  LB_CU(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n, 0, 63));
  LB_CU(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_emit, d_new, (int)(n - 1)));
  char *d_tmp;
  sc.reserve(&d_tmp, std::max(sort_bytes, scan_bytes));
  LB_CU(sc.commit());
  LB_CU(cudaMemcpy(d_boxes, h_boxes, sizeof(DBox) * n, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  LB_CU(cudaEventCreate(&e0));
  LB_CU(cudaEventCreate(&e1));
  LB_CU(cudaEventRecord(e0));
  DBox scene;
  std::memcpy(&scene, &rb, sizeof(scene));
  const int T = 256;
  k_morton<<<(n + T - 1) / T, T>>>(d_boxes, n, scene, d_keys, d_vals);
  LB_CU(cub::DeviceRadixSort::SortPairs(d_tmp, sort_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n, 0, 63));
  LB_CU(cudaMemset(d_flags, 0, sizeof(int) * n));
  LB_CU(cudaMemset(d_height, 0, sizeof(int) * n));
  k_tree<<<(n - 1 + T - 1) / T, T>>>(d_keys2, (int)n, d_left, d_right, d_lo, d_hi, d_pi, d_pl);
  k_fit<<<(n + T - 1) / T, T>>>(d_boxes, d_vals2, (int)n, d_left, d_right, d_lo, d_hi, d_pi, d_pl, d_node_box, d_height, d_flags);
  k_flag<<<(n - 1 + T - 1) / T, T>>>(d_lo, d_hi, (int)n, d_emit);
  LB_CU(cub::DeviceScan::ExclusiveSum(d_tmp, scan_bytes, d_emit, d_new, (int)(n - 1)));
  int last_new = 0, last_emit = 0, h_height = 0;
  LB_CU(cudaMemcpy(&last_new, d_new + (n - 2), sizeof(int), cudaMemcpyDeviceToHost));
  LB_CU(cudaMemcpy(&last_emit, d_emit + (n - 2), sizeof(int), cudaMemcpyDeviceToHost));
  LB_CU(cudaMemcpy(&h_height, d_height, sizeof(int), cudaMemcpyDeviceToHost));
  const int n_out = last_new + last_emit;
  LB_CU(sc.alloc(&d_out, (size_t)n_out));
  k_emit<<<(n - 1 + T - 1) / T, T>>>(d_boxes, d_vals2, (int)n, d_left, d_right, d_lo, d_hi, d_node_box, d_emit, d_new, d_out);
  LB_CU(cudaEventRecord(e1));
  LB_CU(cudaDeviceSynchronize());
  LB_CU(cudaGetLastError());
  if (gpu_ms) cudaEventElapsedTime(gpu_ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  nodes_out.reserve((size_t)n_out + n_out / 8 + 131072);  // room for the trees the caller appends (SAH top, top-level joins): no 100 MB reallocation
  nodes_out.resize(n_out);
  LB_CU(cudaMemcpy(nodes_out.data(), d_out, sizeof(BvhNode) * n_out, cudaMemcpyDeviceToHost));
  LB_CU(cudaMemcpy(order_out.data(), d_vals2, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
  *root_code = 0;  // Karras' root is internal node 0 and, with n > 4, it is emitted first (new index 0)
  *depth = h_height;
  return cudaSuccess;
}


// =====================================================================================================================
// PLOC — parallel locally-ordered clustering (Meister & Bittner, "Parallel Locally-Ordered Clustering for Bounding
// Volume Hierarchy Construction", TVCG 2018).  Bottom-up agglomerative build on the Morton-sorted triangles: every
// cluster looks r positions to the left and right for the neighbour whose merged box has the smallest surface area;
// mutual nearest neighbours merge into a new node; the cluster array is compacted (one scan) and the loop repeats until one
// cluster is left (~log_1.6 n rounds).  Unlike the radix tree of the LBVH — whose topology is dictated by Morton
// prefixes — every merge is chosen by the SAH's own measure, so the tree traverses close to the host binned-SAH tree while
// still building in a few milliseconds on the device.
//
// Nodes: ids [0, n) are the sorted triangles, [n, 2n-1) internal nodes in creation order (children always precede parents).
// Post-pass: subtrees of <= 4 triangles collapse into one leaf; leaf positions come from a top-down offset pass run round
// by round in reverse creation order (a parent's round is always later than its children's).
// =====================================================================================================================
namespace {

__device__ __forceinline__ float union_half_area(const DBox &a, const DBox &b) {
  float dx = fmaxf(a.hi[0], b.hi[0]) - fminf(a.lo[0], b.lo[0]);
  float dy = fmaxf(a.hi[1], b.hi[1]) - fminf(a.lo[1], b.lo[1]);
  float dz = fmaxf(a.hi[2], b.hi[2]) - fminf(a.lo[2], b.lo[2]);
  return dx * dy + dy * dz + dz * dx;
}

__device__ __forceinline__ float half_area(const DBox &a) {
  float dx = a.hi[0] - a.lo[0], dy = a.hi[1] - a.lo[1], dz = a.hi[2] - a.lo[2];
  return dx * dy + dy * dz + dz * dx;
}

constexpr float kPlocCostNode = 1.2f;  // cost of one two-box node visit relative to one triangle test (as the host SAH builder)

__global__ void k_ploc_init(const DBox *boxes, const uint32_t *vals, int n, DBox *cbox, int *cid, DBox *node_box, int *node_count,
                            int *node_height, float *node_cost) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  DBox b = boxes[vals[i]];
  cbox[i] = b, cid[i] = i;
  node_box[i] = b, node_count[i] = 1, node_height[i] = 0, node_cost[i] = half_area(b);
}

// nearest neighbour within +-r positions by merged surface area; ties go to the smaller index (deterministic)
__global__ void k_ploc_nn(const DBox *cbox, int m, int r, int *nn) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const DBox me = cbox[i];
  float best = 3.402823466e+38f;
  int bj = -1;
  const int lo = max(0, i - r), hi = min(m - 1, i + r);
  for (int j = lo; j <= hi; ++j) {
    if (j == i) continue;
    float a = union_half_area(me, cbox[j]);
    if (a < best) best = a, bj = j;
  }
  nn[i] = bj;
}

// keys for the compaction scan: low word = cluster survives (1) or is absorbed by its partner (0); high word = cluster is
// the LEFT member of a merging pair, i.e. creates a node
__global__ void k_ploc_flags(const int *nn, int m, unsigned long long *key) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int j = nn[i];
  const bool mutual = j >= 0 && nn[j] == i;
  const unsigned long long keep = (mutual && j < i) ? 0ull : 1ull;
  const unsigned long long make = (mutual && i < j) ? 1ull : 0ull;
  key[i] = keep | (make << 32);
}

__global__ void k_ploc_apply(const DBox *cbox, const int *cid, const int *nn, const unsigned long long *key,
                             const unsigned long long *scan, int m, int node_base, DBox *cbox_out, int *cid_out, DBox *node_box,
                             int *node_left, int *node_right, int *node_count, int *node_height, float *node_cost, int *node_leaf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const unsigned long long k = key[i];
  if (!(k & 1ull)) return;  // absorbed
  const unsigned long long sc = scan[i];
  const int pos = (int)(uint32_t)sc;
  if (k >> 32) {
    const int j = nn[i];
    const int nid = node_base + (int)(uint32_t)(sc >> 32);
    const DBox a = cbox[i], b = cbox[j];
    DBox u;
#pragma unroll
    for (int q = 0; q < 3; ++q) u.lo[q] = fminf(a.lo[q], b.lo[q]), u.hi[q] = fmaxf(a.hi[q], b.hi[q]);
    const int l = cid[i], rr = cid[j];
    const int cnt = node_count[l] + node_count[rr];
    node_box[nid] = u, node_left[nid] = l, node_right[nid] = rr, node_count[nid] = cnt;
    // SAH decision, bottom-up: a subtree of <= 4 triangles becomes ONE leaf only if testing all its triangles is cheaper than
    // descending (the host builder decides the same way, bvh_build.cpp)
    const float area = half_area(u);
    const float split = kPlocCostNode * area + node_cost[l] + node_cost[rr];
    const float as_leaf = (float)cnt * area;
    const bool leaf = cnt <= kMaxLeafTris && as_leaf <= split;
    node_cost[nid] = leaf ? as_leaf : split;
    node_leaf[nid] = leaf ? 1 : 0;
    node_height[nid] = leaf ? 0 : 1 + max(node_height[l], node_height[rr]);
    cbox_out[pos] = u, cid_out[pos] = nid;
  } else {
    cbox_out[pos] = cbox[i], cid_out[pos] = cid[i];
  }
}

__global__ void k_ploc_emit_flag(const int *node_leaf, const int *node_inside, int n, int n_internal, int *emit) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_internal) return;
  emit[i] = (!node_leaf[n + i] && !node_inside[n + i]) ? 1 : 0;
}

// top-down leaf offsets for the internal nodes [first, last) of one round (parents of later rounds are done)
__global__ void k_ploc_offsets(int n, int first, int last, const int *node_left, const int *node_right, const int *node_count,
                               const int *node_leaf, int *node_offset, int *node_inside, const uint32_t *vals, uint32_t *order_out) {
  int t = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= last) return;
  const int off = node_offset[t];
  const int l = node_left[t], r = node_right[t];
  const int cl = node_count[l];
  const int inside = node_inside[t] | node_leaf[t];  // everything below a leaf-subtree root is part of that leaf
  if (l < n) order_out[off] = vals[l]; else node_offset[l] = off, node_inside[l] = inside;
  if (r < n) order_out[off + cl] = vals[r]; else node_offset[r] = off + cl, node_inside[r] = inside;
}

__device__ __forceinline__ int ploc_child_code(int c, int n, int pos, const int *node_count, const int *node_leaf, const int *new_index) {
  if (c < n) return make_leaf((uint32_t)pos, 1, false);
  if (node_leaf[c]) return make_leaf((uint32_t)pos, (uint32_t)node_count[c], false);
  return new_index[c - n];
}

__global__ void k_ploc_emit(int n, int n_internal, const int *node_left, const int *node_right, const int *node_count, const int *node_leaf,
                            const int *node_offset, const DBox *node_box, const int *emit, const int *new_index, BvhNode *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_internal || !emit[i]) return;
  const int t = n + i;
  const int l = node_left[t], r = node_right[t];
  const DBox a = node_box[l], b = node_box[r];
  const int off = node_offset[t];
  BvhNode nd;
  nd.n0 = make_float4(a.lo[0], a.hi[0], a.lo[1], a.hi[1]);
  nd.n1 = make_float4(b.lo[0], b.hi[0], b.lo[1], b.hi[1]);
  nd.n2 = make_float4(a.lo[2], a.hi[2], b.lo[2], b.hi[2]);
  nd.n3 = make_int4(ploc_child_code(l, n, off, node_count, node_leaf, new_index),
                    ploc_child_code(r, n, off + node_count[l], node_count, node_leaf, new_index), 0, 0);
  out[new_index[i]] = nd;
}

}  // namespace

cudaError_t ploc_build(const Box *h_boxes, uint32_t n, int radius, std::vector<BvhNode> &nodes_out, std::vector<uint32_t> &order_out,
                       int *root_code, Box *root_box, int *depth, float *gpu_ms) {
  nodes_out.clear();
  order_out.resize(n);
  Box rb;
  rb.reset();
  for (uint32_t i = 0; i < n; ++i) rb.grow(h_boxes[i]);
  *root_box = rb;
  *depth = 0;
  if (gpu_ms) *gpu_ms = 0.0f;
  if (n <= (uint32_t)kMaxLeafTris) {
    for (uint32_t i = 0; i < n; ++i) order_out[i] = i;
    *root_code = make_leaf(0, n, false);
    return cudaSuccess;
  }
  radius = radius < 1 ? 1 : (radius > 64 ? 64 : radius);
  BuildClock clk;
  Scratch sc;
  DBox *d_boxes, *d_cbox[2], *d_node_box;
  unsigned long long *d_keys, *d_keys2, *d_flag, *d_scan;
  uint32_t *d_vals, *d_vals2, *d_order;
  int *d_cid[2], *d_nn, *d_left, *d_right, *d_count, *d_height, *d_offset, *d_emit, *d_new, *d_leaf, *d_inside;
  float *d_cost;
  const size_t n_nodes = 2 * (size_t)n;
  sc.reserve(&d_boxes, n);
  sc.reserve(&d_cbox[0], n);
  sc.reserve(&d_cbox[1], n);
  sc.reserve(&d_node_box, n_nodes);
  sc.reserve(&d_keys, n);
  sc.reserve(&d_keys2, n);
  sc.reserve(&d_flag, n);
  sc.reserve(&d_scan, n);
  sc.reserve(&d_vals, n);
  sc.reserve(&d_vals2, n);
  sc.reserve(&d_order, n);
  sc.reserve(&d_cid[0], n);
  sc.reserve(&d_cid[1], n);
  sc.reserve(&d_nn, n);
  sc.reserve(&d_left, n_nodes);
  sc.reserve(&d_right, n_nodes);
  sc.reserve(&d_count, n_nodes);
  sc.reserve(&d_height, n_nodes);
  sc.reserve(&d_offset, n_nodes);
  sc.reserve(&d_leaf, n_nodes);
  sc.reserve(&d_inside, n_nodes);
  sc.reserve(&d_cost, n_nodes);
  sc.reserve(&d_emit, n);
  sc.reserve(&d_new, n);
  size_t scan_bytes = 0, sort_bytes = 0, scan2_bytes = 0;
  LB_CU(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_flag, d_scan, (int)n));
  LB_CU(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n, 0, 63));
  LB_CU(cub::DeviceScan::ExclusiveSum(nullptr, scan2_bytes, d_emit, d_new, (int)n));
  char *d_tmp;
  sc.reserve(&d_tmp, std::max(std::max(scan_bytes, sort_bytes), scan2_bytes));
  LB_CU(sc.commit());
  LB_CU(cudaMemset(d_leaf, 0, sizeof(int) * n_nodes));
  LB_CU(cudaMemset(d_inside, 0, sizeof(int) * n_nodes));
  unsigned long long *h_last = nullptr;  // pinned: (scan, key) of the last cluster of the round
  LB_CU(cudaHostAlloc((void **)&h_last, 2 * sizeof(unsigned long long), cudaHostAllocDefault));
  struct Unpin {
    void *p;
    ~Unpin() { cudaFreeHost(p); }
  } unpin{h_last};

  clk.lap("cudaMalloc");
  LB_CU(cudaMemcpy(d_boxes, h_boxes, sizeof(DBox) * n, cudaMemcpyHostToDevice));
  clk.lap("boxes to device");
  cudaEvent_t e0, e1;
  LB_CU(cudaEventCreate(&e0));
  LB_CU(cudaEventCreate(&e1));
  LB_CU(cudaEventRecord(e0));
  DBox scene;
  std::memcpy(&scene, &rb, sizeof(scene));
  const int T = 256;
  k_morton<<<(n + T - 1) / T, T>>>(d_boxes, n, scene, d_keys, d_vals);
  LB_CU(cub::DeviceRadixSort::SortPairs(d_tmp, sort_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n, 0, 63));
  clk.lap("morton + sort");
  k_ploc_init<<<(n + T - 1) / T, T>>>(d_boxes, d_vals2, (int)n, d_cbox[0], d_cid[0], d_node_box, d_count, d_height, d_cost);

  std::vector<int> round_first;  // first internal node id of every round (+ end)
  int m = (int)n, node_base = (int)n, cur = 0;
  for (int round = 0; m > 1; ++round) {
    if (round > 4096) return cudaErrorUnknown;  // cannot happen: every round merges at least the globally closest pair
    round_first.push_back(node_base);
    k_ploc_nn<<<(m + T - 1) / T, T>>>(d_cbox[cur], m, radius, d_nn);
    k_ploc_flags<<<(m + T - 1) / T, T>>>(d_nn, m, d_flag);
    LB_CU(cub::DeviceScan::ExclusiveSum(d_tmp, scan_bytes, d_flag, d_scan, m));
    k_ploc_apply<<<(m + T - 1) / T, T>>>(d_cbox[cur], d_cid[cur], d_nn, d_flag, d_scan, m, node_base, d_cbox[1 - cur], d_cid[1 - cur],
                                         d_node_box, d_left, d_right, d_count, d_height, d_cost, d_leaf);
    LB_CU(cudaMemcpyAsync(&h_last[0], d_scan + (m - 1), sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    LB_CU(cudaMemcpyAsync(&h_last[1], d_flag + (m - 1), sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    LB_CU(cudaStreamSynchronize(0));
    const unsigned long long tot = h_last[0] + h_last[1];
    const int kept = (int)(uint32_t)tot, made = (int)(uint32_t)(tot >> 32);
    if (made <= 0 || kept != m - made) return cudaErrorUnknown;
    m = kept, node_base += made, cur = 1 - cur;
  }
  round_first.push_back(node_base);
  clk.lap("clustering rounds");
  const int n_internal = node_base - (int)n;  // == n - 1
  int root = 0, h_height = 0;
  LB_CU(cudaMemcpy(&root, d_cid[cur], sizeof(int), cudaMemcpyDeviceToHost));
  LB_CU(cudaMemcpy(&h_height, d_height + root, sizeof(int), cudaMemcpyDeviceToHost));
  // leaf positions, top-down round by round
  {
    const int zero = 0;
    LB_CU(cudaMemcpy(d_offset + root, &zero, sizeof(int), cudaMemcpyHostToDevice));
    for (int g = (int)round_first.size() - 2; g >= 0; --g) {
      const int first = round_first[g], last = round_first[g + 1];
      if (last > first)
        k_ploc_offsets<<<(last - first + T - 1) / T, T>>>((int)n, first, last, d_left, d_right, d_count, d_leaf, d_offset, d_inside, d_vals2, d_order);
    }
  }
  k_ploc_emit_flag<<<(n_internal + T - 1) / T, T>>>(d_leaf, d_inside, (int)n, n_internal, d_emit);
  LB_CU(cub::DeviceScan::ExclusiveSum(d_tmp, scan2_bytes, d_emit, d_new, n_internal));
  int last_new = 0, last_emit = 0;
  LB_CU(cudaMemcpy(&last_new, d_new + (n_internal - 1), sizeof(int), cudaMemcpyDeviceToHost));
  LB_CU(cudaMemcpy(&last_emit, d_emit + (n_internal - 1), sizeof(int), cudaMemcpyDeviceToHost));
  const int n_out = last_new + last_emit;
  BvhNode *d_out;
  LB_CU(sc.alloc(&d_out, (size_t)std::max(n_out, 1)));
  k_ploc_emit<<<(n_internal + T - 1) / T, T>>>((int)n, n_internal, d_left, d_right, d_count, d_leaf, d_offset, d_node_box, d_emit, d_new, d_out);
  int root_new = 0;
  LB_CU(cudaMemcpy(&root_new, d_new + (root - (int)n), sizeof(int), cudaMemcpyDeviceToHost));
  LB_CU(cudaEventRecord(e1));
  LB_CU(cudaDeviceSynchronize());
  LB_CU(cudaGetLastError());
  if (gpu_ms) cudaEventElapsedTime(gpu_ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  clk.lap("offsets + emit");
  nodes_out.reserve((size_t)n_out + n_out / 8 + 131072);  // room for the trees the caller appends (SAH top, top-level joins): no 100 MB reallocation
  nodes_out.resize(n_out);
  LB_CU(cudaMemcpy(nodes_out.data(), d_out, sizeof(BvhNode) * n_out, cudaMemcpyDeviceToHost));
  LB_CU(cudaMemcpy(order_out.data(), d_order, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
  clk.lap("tree to host");
  *root_code = root_new;  // n > 4: the root holds more than 4 triangles, cannot be a leaf and is emitted
  *depth = h_height;
  return cudaSuccess;
}

}  // namespace nrb
