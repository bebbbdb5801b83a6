// api.cu — C-ABI of include/nrays_b200.h: scene upload (Scene::new, src/scene.rs:119-133) and the
// wavefront driver that replaces scene::render (src/scene.rs:29-116).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nrays_b200.h"
#include "bvh_build.h"
#include "kernels.h"
#include "lbvh.h"

using namespace nrb;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define CU(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return fail(NRB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));                \
  } while (0)

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  cudaError_t ensure(size_t n) {
    if (n <= bytes) return cudaSuccess;
    release();
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  template <class T>
  T *as() const {
    return reinterpret_cast<T *>(p);
  }
};

// Large POD host array WITHOUT value-initialisation (a std::vector zero-fills on one thread; these arrays are filled by
// parallel loops right after allocation).
template <class T>
struct HostArray {
  std::unique_ptr<T[]> p;
  size_t n = 0;
  void resize_uninit(size_t count) {
    p.reset(count ? new T[count] : nullptr);
    n = count;
  }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  T *data() { return p.get(); }
  const T *data() const { return p.get(); }
  T &operator[](size_t i) { return p[i]; }
  const T &operator[](size_t i) const { return p[i]; }
};

template <class T>
cudaError_t upload(DevBuf &b, const HostArray<T> &v) {
  size_t n = std::max<size_t>(v.size() * sizeof(T), 16);
  cudaError_t e = b.ensure(n);
  if (e != cudaSuccess) return e;
  if (!v.empty()) e = cudaMemcpy(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return e;
}

template <class T>
cudaError_t upload(DevBuf &b, const std::vector<T> &v) {
  size_t n = std::max<size_t>(v.size() * sizeof(T), 16);
  cudaError_t e = b.ensure(n);
  if (e != cudaSuccess) return e;
  if (!v.empty()) e = cudaMemcpy(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return e;
}

// Host-side data-parallel loop for the per-triangle phases of Scene::new (transform, gathers): contiguous chunks, one per
// hardware thread (NRB_BVH_THREADS overrides); small ranges run inline.
template <class F>
void parallel_for(size_t n, F &&fn) {
  unsigned hw = std::thread::hardware_concurrency();
  if (const char *e = getenv("NRB_BVH_THREADS")) hw = (unsigned)std::max(1, atoi(e));
  if (n < 65536 || hw < 2) {
    fn(0, n);
    return;
  }
  const size_t chunk = (n + hw - 1) / hw;
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < hw; ++t) {
    const size_t lo = (size_t)t * chunk, hi = std::min(n, lo + chunk);
    if (lo >= hi) break;
    pool.emplace_back([&fn, lo, hi]() { fn(lo, hi); });
  }
  for (auto &t : pool) t.join();
}

// A few thousand uneven tasks (subtrees): worker threads take them one at a time from a shared counter.
template <class F>
void parallel_tasks(size_t n, F &&fn) {
  unsigned hw = std::thread::hardware_concurrency();
  if (const char *e = getenv("NRB_BVH_THREADS")) hw = (unsigned)std::max(1, atoi(e));
  if (n < 2 || hw < 2) {
    for (size_t i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<size_t> next(0);
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < std::min<size_t>(hw, n); ++t)
    pool.emplace_back([&fn, &next, n]() {
      for (size_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
    });
  for (auto &t : pool) t.join();
}

// Sample slots per batch and shadow-queue entries per wave.  Every batch pays its own launches and the tail of each persistent
// kernel (its slowest warps), so frames larger than one batch want big batches: C5 (132.7 M slots) 30.3 ms at 8 Mi slots,
// 28.9 at 16 Mi, 28.4 at 32 Mi, 28.2 at 64 Mi; C4 (16.6 M slots) 8.27 -> 7.90 ms once it is one batch.  32 Mi slots cost
// 0.5 GB of hit records + 1.5 GB of shadow queue + up to 1.5 GB per ray queue — small change on a 180 GB part.
constexpr size_t kBatchSlots = 32u << 20;
constexpr size_t kShadowCap = 32u << 20;

size_t env_size(const char *name, size_t dflt) {
  const char *s = getenv(name);
  if (!s || !*s) return dflt;
  return (size_t)strtoull(s, nullptr, 10);
}

}  // namespace

struct NrbScene {
  int device = 0;
  cudaStream_t stream = nullptr;      // stream in use
  cudaStream_t own_stream = nullptr;  // created by the library
  int sm_count = 148;
  // scene tables
  DevBuf d_nodes, d_tris, d_tri_uvs, d_shapes, d_node_info, d_materials, d_textures, d_texels, d_lights, d_planes,
      d_candidates, d_nmaps;
  SceneView view{};
  bool has_shapes = false;
  bool has_nmap = false;  // some node carries a depth-shift texture: general trace kernel (kernels.cu: trace_general_kernel)
  int child_factor = 0;  // max secondary rays per ray (reflection + refraction possible in this scene)
  uint32_t refl_chain_max = 0;  // most reflections one ray chain can make before its energy is spent (0xFFFFFFFF: unbounded)
  uint64_t n_bvh_nodes = 0, n_tris = 0, scene_bytes = 0, node_bytes = 64;
  int grid_trace = 148, grid_tail = 148;
  NrbBuildInfo build_info{};
  // frame state
  DevBuf d_q[2][3], d_hits, d_sq[3], d_accum, d_counters, d_wave, d_out, d_out8;
  uint32_t q_cap[2] = {0, 0}, sq_cap = 0, hits_cap = 0;
  Counters *h_counters = nullptr;  // pinned mirror
  uint32_t *h_wave_counts = nullptr;  // pinned: exact ray count of every wave of the current frame
  size_t h_wave_cap = 0;
  int grid_shade = 148;
  std::vector<cudaEvent_t> events;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  // nrb_render: the image's device->host copy runs on its own stream while the tail phase finishes, then the pixels the
  // tail changed follow as an ordered patch
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_early = nullptr, ev_copied = nullptr;
  // wave ray counts travel to the host on their own stream: a 4-byte device->host copy in the render stream would sit between
  // two dependent kernels for ~10 us (DMA set-up), once per wave
  cudaStream_t count_stream = nullptr;

  ~NrbScene() {
    cudaSetDevice(device);
    if (h_counters) cudaFreeHost(h_counters);
    if (h_wave_counts) cudaFreeHost(h_wave_counts);
    for (auto e : events) cudaEventDestroy(e);
    if (ev_begin) cudaEventDestroy(ev_begin);
    if (ev_end) cudaEventDestroy(ev_end);
    if (own_stream) cudaStreamDestroy(own_stream);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (count_stream) cudaStreamDestroy(count_stream);
    if (ev_early) cudaEventDestroy(ev_early);
    if (ev_copied) cudaEventDestroy(ev_copied);
  }
};

namespace {

// ---------------------------------------------------------------------------------------------
// Scene::new — flatten + BVH build + upload
// ---------------------------------------------------------------------------------------------
bool shape_has_uv(int kind) { return kind == NRB_SHAPE_BALL || kind == NRB_SHAPE_CUBOID || kind == NRB_SHAPE_TRIMESH; }

int relayout_bfs(std::vector<BvhNode> &nodes, int &root_all, int &root_opaque, std::vector<Candidate> &cands,
                 std::vector<Candidate> &nmaps, const std::vector<char> &dead) {
  // Level-order relabel from root_all so the top of the tree is one contiguous prefix of the array
  // (what the closest-hit kernel touches for every ray; also the part worth staging on chip).
  if (nodes.empty()) return 0;
  std::vector<int> remap(nodes.size(), -1), order;
  order.reserve(nodes.size());
  auto add_root = [&](int r) {
    if (r >= 0 && r != kEmpty && remap[r] < 0) {
      remap[r] = (int)order.size();
      order.push_back(r);
    }
  };
  add_root(root_all);
  add_root(root_opaque);                   // separate from root_all when the closest-hit tree is unified (flatten_scene)
  for (auto &c : cands) add_root(c.root);
  for (auto &c : nmaps) add_root(c.root);  // depth-shift nodes keep their own sub-trees outside root_all
  // Level order for the TOP of the tree (what every ray touches: kept contiguous, a few MB), builder order below it: the
  // builders emit a parent before its subtrees (depth-first), which keeps a deep path's nodes close together and costs one
  // linear sweep here instead of a random-access BFS over millions of nodes.  NRB_RELAYOUT_TOP=0 restores the full BFS.
  const size_t top_cap = env_size("NRB_RELAYOUT_TOP", 1u << 16);
  size_t head = 0;
  for (; head < order.size() && (top_cap == 0 || order.size() < top_cap); ++head) {
    const BvhNode &n = nodes[order[head]];
    int ch[2] = {n.n3.x, n.n3.y};
    for (int c : ch)
      if (c >= 0 && remap[c] < 0) {
        remap[c] = (int)order.size();
        order.push_back(c);
      }
  }
  if (head < order.size() || order.size() < nodes.size()) {
    // everything not placed yet that is reachable: sweep in builder order (children of placed nodes are reachable by construction;
    // unreachable nodes do not exist — every builder node hangs off root_all or an nmap root)
    for (size_t i = 0; i < nodes.size(); ++i)
      if (remap[i] < 0 && !(i < dead.size() && dead[i])) {  // dead: tops of device-built trees replaced by a SAH top
        remap[i] = (int)order.size();
        order.push_back((int)i);
      }
  }
  std::vector<BvhNode> out(order.size());
  parallel_for(order.size(), [&](size_t lo_i, size_t hi_i) {
    for (size_t i = lo_i; i < hi_i; ++i) {
      BvhNode n = nodes[order[i]];
      if (n.n3.x >= 0) n.n3.x = remap[n.n3.x];
      if (n.n3.y >= 0) n.n3.y = remap[n.n3.y];
      out[i] = n;
    }
  });
  auto fix = [&](int &c) {
    if (c >= 0 && c != kEmpty) c = remap[c];
  };
  fix(root_all);
  fix(root_opaque);
  for (auto &c : cands) fix(c.root);
  for (auto &c : nmaps) fix(c.root);
  nodes.swap(out);
  return 0;
}

// Host-side result of Scene::new: validated tables + BVH, ready to upload.
struct HostScene {
  std::vector<BvhNode> nodes;
  HostArray<Tri> tris;
  HostArray<TriUV> tri_uvs;
  std::vector<Shape> shapes;
  std::vector<NodeInfo> node_info;
  std::vector<Material> materials;
  std::vector<Texture> textures;
  std::vector<Light> lights;
  std::vector<int> planes;
  std::vector<Candidate> candidates;
  std::vector<Candidate> nmaps;  // nodes with a depth-shift texture (SceneNode.nmap)
  int root_all = kEmpty, root_opaque = kEmpty;
  int shadow_samples = 0;
  bool any_refl = false, any_refr = false;
  uint64_t n_source_tris = 0;  // triangles of the scene (H.tris holds one entry per LEAF reference: twice that with a unified tree)
  bool unified = false;  // root_all is a separate tree over all triangles (the shadow structure hangs off root_opaque + candidates)
  uint32_t refl_chain_max = 0;
  int depth_tri = 0, depth_mid = 0, depth_top = 0;
  float gpu_build_ms = 0.0f;  // device time of the LBVH kernels (NRB_BUILDER_LBVH)
};

struct PhaseTimer {  // NRB_BUILD_TIMES=1: per-phase host times of Scene::new on stderr
  bool on = getenv("NRB_BUILD_TIMES") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void lap(const char *what) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[nrb] build: %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

int flatten_scene(const NrbSceneDesc &d, HostScene &H, uint32_t builder = NRB_BUILDER_SAH) {
  PhaseTimer pt;
  if (d.n_nodes && !d.nodes) return fail(NRB_ERR_INVALID_ARG, "nodes is NULL");
  if (d.n_lights && !d.lights) return fail(NRB_ERR_INVALID_ARG, "lights is NULL");
  if (d.n_materials && !d.materials) return fail(NRB_ERR_INVALID_ARG, "materials is NULL");
  if (d.n_textures && !d.textures) return fail(NRB_ERR_INVALID_ARG, "textures is NULL");
  if (d.n_texels && !d.texels) return fail(NRB_ERR_INVALID_ARG, "texels is NULL");
  if (d.n_texels >= 0xFFFFFFFFull) return fail(NRB_ERR_INVALID_ARG, "texel pool too large (>= 2^32 texels)");

  std::vector<Texture> textures(d.n_textures);
  for (uint32_t i = 0; i < d.n_textures; ++i) {
    const NrbTextureDesc &t = d.textures[i];
    // ImageData::new asserts dims >= 1 and len == w*h (src/texture2d.rs:17-19)
    if (t.width < 1 || t.height < 1 || t.texel_offset + (uint64_t)t.width * t.height > d.n_texels)
      return fail(NRB_ERR_INVALID_ARG, "texture " + std::to_string(i) + " outside the texel pool");
    if ((t.interpolation != NRB_INTERP_BILINEAR && t.interpolation != NRB_INTERP_NEAREST) ||
        (t.overflow != NRB_OVERFLOW_CLAMP && t.overflow != NRB_OVERFLOW_WRAP))
      return fail(NRB_ERR_INVALID_ARG, "texture " + std::to_string(i) + " has an unknown sampling mode");
    textures[i] = Texture{t.width, t.height, t.interpolation, t.overflow, (uint32_t)t.texel_offset,
                          (uint32_t)(d.n_texels - t.texel_offset)};
  }
  std::vector<Material> materials(d.n_materials);
  for (uint32_t i = 0; i < d.n_materials; ++i) {
    const NrbMaterialDesc &m = d.materials[i];
    if (m.kind != NRB_MAT_PHONG && m.kind != NRB_MAT_NORMAL && m.kind != NRB_MAT_UV)
      return fail(NRB_ERR_UNSUPPORTED, "material " + std::to_string(i) + ": only Phong / Normal / UV cross the C-ABI");
    if (m.texture >= (int)d.n_textures || m.alpha_texture >= (int)d.n_textures)
      return fail(NRB_ERR_INVALID_ARG, "material " + std::to_string(i) + " texture index out of range");
    Material mm{};
    mm.kind = m.kind;
    for (int k = 0; k < 3; ++k) mm.ambient[k] = m.ambient[k], mm.diffuse[k] = m.diffuse[k], mm.specular[k] = m.specular[k];
    mm.shininess = m.shininess;
    mm.tex = m.kind == NRB_MAT_PHONG ? m.texture : -1;
    mm.alpha_tex = m.kind == NRB_MAT_PHONG ? m.alpha_texture : -1;
    materials[i] = mm;
  }
  std::vector<Light> lights(d.n_lights);
  int shadow_samples = 0;
  for (uint32_t i = 0; i < d.n_lights; ++i) {
    const NrbLightDesc &l = d.lights[i];
    if (l.racsample == 0 || l.racsample > 255)
      return fail(NRB_ERR_INVALID_ARG, "light " + std::to_string(i) + ": racsample must be in [1, 255]");
    Light L{};
    for (int k = 0; k < 3; ++k) L.pos[k] = (float)l.pos[k], L.color[k] = l.color[k];
    L.radius = (float)l.radius;
    L.racsample = l.racsample;
    lights[i] = L;
    shadow_samples += (int)(l.racsample * l.racsample);
  }
  if (d.n_lights > 65535) return fail(NRB_ERR_INVALID_ARG, "too many lights (> 65535)");

  // ---- nodes -------------------------------------------------------------------------------
  std::vector<NodeInfo> node_info(d.n_nodes);
  std::vector<Shape> shapes;
  std::vector<int> planes;
  std::vector<int> shape_of_node(d.n_nodes, -1);
  std::vector<char> is_cand(d.n_nodes, 0), is_nmap(d.n_nodes, 0);
  bool any_refl = false, any_refr = false;
  uint64_t total_tris = 0;
  for (uint32_t i = 0; i < d.n_nodes; ++i) {
    const NrbNodeDesc &n = d.nodes[i];
    if (n.material < 0 || n.material >= (int)d.n_materials)
      return fail(NRB_ERR_INVALID_ARG, "node " + std::to_string(i) + " material index out of range");
    if (n.shape < NRB_SHAPE_BALL || n.shape > NRB_SHAPE_TRIMESH)
      return fail(NRB_ERR_INVALID_ARG, "node " + std::to_string(i) + " has an unknown shape kind");
    if (n.nmap_texture >= (int)d.n_textures)
      return fail(NRB_ERR_INVALID_ARG, "node " + std::to_string(i) + " nmap texture index out of range");
    // the depth shift applies only where the cast returns uvs (src/scene_node.rs:63); elsewhere the texture is inert
    is_nmap[i] = (n.nmap_texture >= 0 && shape_has_uv(n.shape)) ? 1 : 0;
    const Material &m = materials[n.material];
    bool cand = (n.alpha < 1.0f) || (m.kind == NRB_MAT_PHONG && m.alpha_tex >= 0) ||
                (m.kind == NRB_MAT_UV && !shape_has_uv(n.shape));
    bool may_refract = (n.alpha != 1.0f) || (m.kind == NRB_MAT_PHONG && m.alpha_tex >= 0) ||
                       (m.kind == NRB_MAT_UV && !shape_has_uv(n.shape));
    is_cand[i] = cand ? 1 : 0;
    any_refl = any_refl || (n.refl_mix != 0.0f);
    any_refr = any_refr || may_refract;
    if (n.refl_mix != 0.0f) {
      // trace_reflection reflects while energy > 0.1 and pays refl_atenuation per bounce (src/scene.rs:204-208): a chain
      // that only meets this node reflects at most ceil(0.9 / att) times (+1 for f32 rounding of the energy sequence)
      const float att = n.refl_atenuation;
      const uint32_t r = (att > 0.0f && 0.9f / att < 1.0e6f) ? (uint32_t)std::ceil(0.9f / att) + 1u : 0xFFFFFFFFu;
      H.refl_chain_max = std::max(H.refl_chain_max, r);
    }
    NodeInfo ni{};
    ni.material = n.material;
    ni.refl_mix = n.refl_mix;
    ni.refl_att = n.refl_atenuation;
    ni.alpha = n.alpha;
    ni.refr_coeff = (float)n.refr_coeff;
    ni.flags = cand ? 1 : 0;
    ni.nmap_tex = is_nmap[i] ? n.nmap_texture : -1;
    node_info[i] = ni;
    if (n.shape == NRB_SHAPE_TRIMESH) {
      if (n.first_index % 3 != 0 || n.tri_count == 0 || n.first_index + 3 * n.tri_count > d.n_indices || !d.indices ||
          !d.positions)
        return fail(NRB_ERR_INVALID_ARG, "node " + std::to_string(i) + ": trimesh index range invalid");
      total_tris += n.tri_count;
    } else {
      Shape s{};
      s.kind = n.shape;
      s.node = (int)i;
      s.solid = n.solid ? 1 : 0;
      for (int k = 0; k < 3; ++k) s.p[k] = (float)n.param[k], s.trans[k] = (float)n.trans[k];
      for (int k = 0; k < 9; ++k) s.rot[k] = (float)n.rot[k];
      shape_of_node[i] = (int)shapes.size();
      if (n.shape == NRB_SHAPE_PLANE) planes.push_back((int)shapes.size());
      shapes.push_back(s);
    }
  }
  if (total_tris >= (1ull << 28)) return fail(NRB_ERR_INVALID_ARG, "too many triangles (>= 2^28)");

  // ---- triangles to world space (f64 transform, then f32) -------------------------------------
  // uninitialised on purpose: a std::vector would zero 280 MB on ONE thread before the parallel transform touches it
  std::unique_ptr<Tri[]> tris_in(new Tri[total_tris]);
  std::unique_ptr<TriUV[]> uvs_in(new TriUV[total_tris]);
  std::unique_ptr<Box[]> tri_box(new Box[total_tris]);
  std::vector<uint64_t> node_tri_begin(d.n_nodes, 0);
  Box scene_box;
  scene_box.reset();
  {
    uint64_t t_out = 0;
    std::atomic<int> bad_node(-1);
    for (uint32_t i = 0; i < d.n_nodes; ++i) {
      const NrbNodeDesc &n = d.nodes[i];
      if (n.shape != NRB_SHAPE_TRIMESH) continue;
      node_tri_begin[i] = t_out;
      const uint64_t base = t_out;
      std::vector<Box> part_box;  // per-chunk scene boxes, merged below
      std::mutex part_mu;
      parallel_for((size_t)n.tri_count, [&](size_t lo_t, size_t hi_t) {
        Box local;
        local.reset();
        for (uint64_t t = lo_t; t < hi_t; ++t) {
          double w[3][3];
          TriUV uv{};
          for (int k = 0; k < 3; ++k) {
            uint64_t vi = (uint64_t)d.indices[n.first_index + 3 * t + k] + n.vertex_base;
            if (vi >= d.n_vertices) {
              bad_node.store((int)i);
              return;
            }
            const float *p = d.positions + 3 * vi;
            for (int r = 0; r < 3; ++r)
              w[k][r] = n.rot[3 * r] * (double)p[0] + n.rot[3 * r + 1] * (double)p[1] + n.rot[3 * r + 2] * (double)p[2] + n.trans[r];
            float uu = d.uvs ? d.uvs[2 * vi] : 0.0f, vv = d.uvs ? d.uvs[2 * vi + 1] : 0.0f;
            if (k == 0) uv.u0 = uu, uv.v0 = vv;
            if (k == 1) uv.u1 = uu, uv.v1 = vv;
            if (k == 2) uv.u2 = uu, uv.v2 = vv;
          }
          Tri tr;
          tr.t0 = make_float4((float)w[0][0], (float)w[0][1], (float)w[0][2], 0.0f);
          int node_id = (int)i;
          std::memcpy(&tr.t0.w, &node_id, 4);
          tr.t1 = make_float4((float)(w[1][0] - w[0][0]), (float)(w[1][1] - w[0][1]), (float)(w[1][2] - w[0][2]), 0.0f);
          int material_id = (int)n.material;  // copy of NodeInfo.material: shade fetches the material without waiting for NodeInfo
          std::memcpy(&tr.t1.w, &material_id, 4);
          tr.t2 = make_float4((float)(w[2][0] - w[0][0]), (float)(w[2][1] - w[0][1]), (float)(w[2][2] - w[0][2]), 0.0f);
          tris_in[base + t] = tr;
          uvs_in[base + t] = uv;
          Box b;
          b.reset();
          float v0[3] = {tr.t0.x, tr.t0.y, tr.t0.z};
          float v1[3] = {tr.t0.x + tr.t1.x, tr.t0.y + tr.t1.y, tr.t0.z + tr.t1.z};
          float v2[3] = {tr.t0.x + tr.t2.x, tr.t0.y + tr.t2.y, tr.t0.z + tr.t2.z};
          b.grow(v0), b.grow(v1), b.grow(v2);
          for (int k = 0; k < 3; ++k) {
            float wk[3] = {(float)w[k][0], (float)w[k][1], (float)w[k][2]};
            b.grow(wk);
          }
          tri_box[base + t] = b;
          local.grow(b);
        }
        std::lock_guard<std::mutex> g(part_mu);
        part_box.push_back(local);
      });
      if (bad_node.load() >= 0) return fail(NRB_ERR_INVALID_ARG, "node " + std::to_string(bad_node.load()) + ": vertex index out of range");
      for (const Box &b : part_box)
        if (b.valid()) scene_box.grow(b);
      t_out += n.tri_count;
    }
  }
  pt.lap("triangles to world space");
  // shape boxes (bounding_volume(&transform), src/scene_node.rs:41), conservative
  std::vector<Box> shape_box(shapes.size());
  for (size_t si = 0; si < shapes.size(); ++si) {
    const NrbNodeDesc &n = d.nodes[shapes[si].node];
    Box b;
    b.reset();
    if (n.shape == NRB_SHAPE_PLANE) {
      shape_box[si] = b;
      continue;
    }
    double he[3];
    switch (n.shape) {
      case NRB_SHAPE_BALL:
        he[0] = he[1] = he[2] = n.param[0];
        break;
      case NRB_SHAPE_CUBOID:
        he[0] = n.param[0], he[1] = n.param[1], he[2] = n.param[2];
        break;
      case NRB_SHAPE_CAPSULE:
        he[0] = n.param[1], he[1] = n.param[0] + n.param[1], he[2] = n.param[1];
        break;
      default:  // cylinder, cone
        he[0] = n.param[1], he[1] = n.param[0], he[2] = n.param[1];
        break;
    }
    for (int r = 0; r < 3; ++r) {
      double e = n.shape == NRB_SHAPE_BALL
                     ? he[0]
                     : std::fabs(n.rot[3 * r]) * he[0] + std::fabs(n.rot[3 * r + 1]) * he[1] + std::fabs(n.rot[3 * r + 2]) * he[2];
      b.lo[r] = (float)(n.trans[r] - e);
      b.hi[r] = (float)(n.trans[r] + e);
    }
    shape_box[si] = b;
    scene_box.grow(b);
  }
  float extent = 0.0f;
  if (scene_box.valid())
    for (int k = 0; k < 3; ++k) extent = std::max(extent, scene_box.hi[k] - scene_box.lo[k]);
  parallel_for((size_t)total_tris, [&](size_t lo_t, size_t hi_t) {
    for (size_t k = lo_t; k < hi_t; ++k) pad_box(tri_box[k], extent);
  });
  for (size_t si = 0; si < shapes.size(); ++si)
    if (shape_box[si].valid()) pad_box(shape_box[si], extent);

  // ---- BVH: opaque triangles | opaque shapes | one sub-root per transparent candidate | top level ----
  BvhBuilder bb;
  bb.nodes.reserve(total_tris / 2 + 16);
  bb.tri_order.reserve(total_tris);
  std::vector<BuildItem> opaque_items, top_items;
  // builds one tree over a set of triangles with the selected builder and appends it to the shared pools
  auto build_set = [&](std::vector<BuildItem> &items, Box *rb, int *code) -> int {
    if (builder == NRB_BUILDER_LBVH || builder == NRB_BUILDER_PLOC) {
      PhaseTimer sub_pt;
      std::vector<Box> boxes(items.size());
      parallel_for(items.size(), [&](size_t lo_t, size_t hi_t) {
        for (size_t k = lo_t; k < hi_t; ++k) boxes[k] = items[k].box;
      });
      std::vector<BvhNode> sub;
      std::vector<uint32_t> order;
      int depth = 0, root = kEmpty;
      float ms = 0.0f;
      cudaError_t e = builder == NRB_BUILDER_PLOC
                          ? ploc_build(boxes.data(), (uint32_t)boxes.size(), (int)env_size("NRB_PLOC_RADIUS", 16), sub, order, &root, rb, &depth, &ms)
                          : lbvh_build(boxes.data(), (uint32_t)boxes.size(), sub, order, &root, rb, &depth, &ms);
      if (e != cudaSuccess) return fail(NRB_ERR_CUDA, std::string("device BVH build: ") + cudaGetErrorString(e));
      sub_pt.lap("  boxes + device build call");
      H.gpu_build_ms += ms;
      const int node_off = (int)bb.nodes.size();
      const uint32_t tri_off = (uint32_t)bb.tri_order.size();
      auto shift = [&](int c) -> int {
        if (c >= 0) return c + node_off;
        uint32_t lc = (uint32_t)~c;
        return ~(int)((((lc >> 3) + tri_off) << 3) | (lc & 7u));
      };
      const size_t n_sub = sub.size();
      if (node_off == 0 && tri_off == 0) {
        bb.nodes = std::move(sub);  // first tree of the pool: the downloaded nodes ARE the pool (no 100 MB copy, no zero fill)
      } else {
        bb.nodes.resize((size_t)node_off + n_sub);
        parallel_for(n_sub, [&](size_t lo_t, size_t hi_t) {
          for (size_t k = lo_t; k < hi_t; ++k) {
            BvhNode nd = sub[k];
            nd.n3.x = shift(nd.n3.x);
            nd.n3.y = shift(nd.n3.y);
            bb.nodes[(size_t)node_off + k] = nd;
          }
        });
      }
      bb.tri_order.resize((size_t)tri_off + order.size());
      parallel_for(order.size(), [&](size_t lo_t, size_t hi_t) {
        for (size_t k = lo_t; k < hi_t; ++k) bb.tri_order[(size_t)tri_off + k] = (uint32_t)items[order[k]].payload;
      });
      *code = shift(root);
      sub_pt.lap("  splice into the node pool");
      const int prev_depth = bb.max_depth_seen;
      bb.max_depth_seen = std::max(prev_depth, depth + 1);
      // ---- SAH top ----------------------------------------------------------------------------------------------------
      // Morton-ordered builders place the BOTTOM of the tree well (neighbouring triangles end up together) and the TOP badly
      // (LBVH: splits dictated by Morton prefixes; PLOC: agglomeration knows nothing about the rays' long walks through an
      // atrium): 48 / 40 node visits per primary ray on C3 against 28 for the host SAH tree.  So the device tree is cut where
      // its subtrees hold <= T triangles and the host's binned-SAH builder rebuilds everything above the cut over those few
      // thousand subtree boxes (milliseconds).  NRB_RESAH_TOP=0 keeps the device tree as built.
      if (env_size("NRB_RESAH_TOP", 1) != 0 && *code >= 0) {
        // triangles below every spliced node and the levels of inner nodes under it.  The top of the tree is walked breadth-first
        // until ~2 k subtrees are open; those are finished in parallel (iterative post-order each), then the top bottom-up.
        std::vector<uint32_t> cnt(n_sub, 0);
        std::vector<uint16_t> dep(n_sub, 0);
        auto leaf_count = [](int c) -> uint32_t { return (((uint32_t)~c >> 1) & 3u) + 1u; };
        auto finish = [&](int li) {  // li's children are done (or leaves)
          const BvhNode &nd = bb.nodes[(size_t)node_off + li];
          const int a = nd.n3.x, b = nd.n3.y;
          cnt[li] = (a >= 0 ? cnt[a - node_off] : leaf_count(a)) + (b >= 0 ? cnt[b - node_off] : leaf_count(b));
          dep[li] = (uint16_t)(1 + std::max<int>(a >= 0 ? dep[a - node_off] : 0, b >= 0 ? dep[b - node_off] : 0));
        };
        {
          std::vector<int> top{*code - node_off};  // breadth-first prefix: children always behind their parent
          size_t head = 0;
          while (head < top.size() && top.size() - head < 2048) {
            const BvhNode &nd = bb.nodes[(size_t)node_off + top[head++]];
            if (nd.n3.x >= 0) top.push_back(nd.n3.x - node_off);
            if (nd.n3.y >= 0) top.push_back(nd.n3.y - node_off);
          }
          parallel_tasks(top.size() - head, [&](size_t r) {
            std::vector<std::pair<int, int>> st;  // (local node, state)
            st.emplace_back(top[head + r], 0);
            while (!st.empty()) {
              auto &t = st.back();
              if (t.second == 0) {
                t.second = 1;
                const BvhNode &nd = bb.nodes[(size_t)node_off + t.first];
                const int a = nd.n3.x, b = nd.n3.y;  // (t is invalidated by the pushes)
                if (a >= 0) st.emplace_back(a - node_off, 0);
                if (b >= 0) st.emplace_back(b - node_off, 0);
              } else {
                finish(t.first);
                st.pop_back();
              }
            }
          });
          for (size_t k = head; k-- > 0;) finish(top[k]);
        }
        const uint32_t total = cnt[*code - node_off];
        sub_pt.lap("  subtree triangle counts");
        // cut size: ~65 k subtrees for meshes up to 1 M triangles, ~33 k beyond (C3: T = 4 -> 2.39 ms against 2.23 for the host
        // SAH tree, 2.94 as built; C4: T = 88 -> 7.75 ms against 7.28, 8.02 as built; the SAH top costs 30-150 ms of host time)
        const uint32_t T = (uint32_t)env_size("NRB_RESAH_LEAF", std::max<size_t>(4, total / (total <= (1u << 20) ? 65536u : 32768u)));
        if (total > 2 * T) {
          std::vector<BuildItem> cut;
          std::vector<int> walk{*code - node_off};
          while (!walk.empty()) {
            const int li = walk.back();
            walk.pop_back();
            const BvhNode nd = bb.nodes[(size_t)node_off + li];
            bb.dead.resize(bb.nodes.size(), 0);
            bb.dead[(size_t)node_off + li] = 1;  // every node above the cut is replaced by the SAH top
            const Box b0{{nd.n0.x, nd.n0.z, nd.n2.x}, {nd.n0.y, nd.n0.w, nd.n2.y}}, b1{{nd.n1.x, nd.n1.z, nd.n2.z}, {nd.n1.y, nd.n1.w, nd.n2.w}};
            const int ch[2] = {nd.n3.x, nd.n3.y};
            const Box *bx[2] = {&b0, &b1};
            for (int k = 0; k < 2; ++k) {
              if (ch[k] >= 0 && cnt[ch[k] - node_off] > T) walk.push_back(ch[k] - node_off);
              else cut.push_back(BuildItem{*bx[k], ch[k]});
            }
          }
          sub_pt.lap("  cut");
          *code = bb.build_payloads(cut, rb);
          sub_pt.lap("  SAH top");
          // true depth of the rebuilt tree (levels of inner nodes on the longest path): SAH top + the device subtree below it
          int deepest = 0;
          std::vector<std::pair<int, int>> st{{*code, 1}};
          while (!st.empty()) {
            const auto [c, lvl] = st.back();
            st.pop_back();
            if (c < 0) continue;
            if ((size_t)c < (size_t)node_off + n_sub) {  // a cut subtree (device-built, untouched)
              deepest = std::max(deepest, lvl - 1 + (int)dep[c - node_off]);
              continue;
            }
            deepest = std::max(deepest, lvl);
            const BvhNode &nd = bb.nodes[(size_t)c];
            st.emplace_back(nd.n3.x, lvl + 1);
            st.emplace_back(nd.n3.y, lvl + 1);
          }
          bb.max_depth_seen = std::max(prev_depth, deepest);
          sub_pt.lap("  depth of the joined tree");
        }
      }
      return NRB_OK;
    }
    *code = bb.build_triangles(items, rb);
    return NRB_OK;
  };
  {
    std::vector<BuildItem> items;
    for (uint32_t i = 0; i < d.n_nodes; ++i) {
      const NrbNodeDesc &n = d.nodes[i];
      if (n.shape != NRB_SHAPE_TRIMESH || is_cand[i] || is_nmap[i]) continue;
      const size_t at = items.size();
      items.resize(at + n.tri_count);
      const uint64_t b0 = node_tri_begin[i];
      parallel_for((size_t)n.tri_count, [&](size_t lo_t, size_t hi_t) {
        for (size_t t = lo_t; t < hi_t; ++t) items[at + t] = BuildItem{tri_box[b0 + t], (int)(b0 + t)};
      });
    }
    PhaseTimer items_pt;
    items_pt.on = items_pt.on && !items.empty();
    items_pt.lap("  build items");
    if (!items.empty()) {
      Box rb;
      int code = kEmpty;
      int brc = build_set(items, &rb, &code);
      if (brc) return brc;
      opaque_items.push_back(BuildItem{rb, code});
    }
  }
  pt.lap("opaque triangle tree");
  int depth_tri = bb.max_depth_seen;
  std::vector<Candidate> candidates, nmaps;
  for (uint32_t i = 0; i < d.n_nodes; ++i) {
    const NrbNodeDesc &n = d.nodes[i];
    if (n.shape == NRB_SHAPE_PLANE) continue;
    if (is_nmap[i]) {
      // own sub-root, outside every flat tree; the box is the REFERENCE's node AABB (HasBoundingVolume::bounding_volume:
      // local box, centre transformed, half extents times |R|; ball: centre -+ r) because it decides the search order
      Candidate c{};
      c.node = (int)i;
      double lc[3] = {0, 0, 0}, lhe[3] = {0, 0, 0};
      if (n.shape == NRB_SHAPE_TRIMESH) {
        std::vector<BuildItem> items;
        for (uint64_t t = 0; t < n.tri_count; ++t) items.push_back(BuildItem{tri_box[node_tri_begin[i] + t], (int)(node_tri_begin[i] + t)});
        Box rb;
        bb.max_depth_seen = 0;
        int code = kEmpty;
        int brc = build_set(items, &rb, &code);
        if (brc) return brc;
        depth_tri = std::max(depth_tri, bb.max_depth_seen);
        c.root = code;
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (uint64_t k = 0; k < 3 * n.tri_count; ++k) {
          const float *p = d.positions + 3 * (n.vertex_base + d.indices[n.first_index + k]);
          for (int a = 0; a < 3; ++a) lo[a] = std::min(lo[a], (double)p[a]), hi[a] = std::max(hi[a], (double)p[a]);
        }
        for (int a = 0; a < 3; ++a) lc[a] = 0.5 * (lo[a] + hi[a]), lhe[a] = 0.5 * (hi[a] - lo[a]);
      } else {
        c.root = make_leaf((uint32_t)shape_of_node[i], 1, true);
        if (n.shape == NRB_SHAPE_CUBOID) lhe[0] = n.param[0], lhe[1] = n.param[1], lhe[2] = n.param[2];
      }
      for (int r = 0; r < 3; ++r) {
        double cw = n.rot[3 * r] * lc[0] + n.rot[3 * r + 1] * lc[1] + n.rot[3 * r + 2] * lc[2] + n.trans[r];
        double he = std::fabs(n.rot[3 * r]) * lhe[0] + std::fabs(n.rot[3 * r + 1]) * lhe[1] + std::fabs(n.rot[3 * r + 2]) * lhe[2];
        if (n.shape == NRB_SHAPE_BALL) cw = n.trans[r], he = n.param[0];
        c.lo[r] = (float)(cw - he), c.hi[r] = (float)(cw + he);
      }
      nmaps.push_back(c);
      continue;
    }
    if (n.shape == NRB_SHAPE_TRIMESH) {
      if (!is_cand[i]) continue;
      std::vector<BuildItem> items;
      for (uint64_t t = 0; t < n.tri_count; ++t) items.push_back(BuildItem{tri_box[node_tri_begin[i] + t], (int)(node_tri_begin[i] + t)});
      Box rb;
      bb.max_depth_seen = 0;
      int code = kEmpty;
      int brc = build_set(items, &rb, &code);
      if (brc) return brc;
      depth_tri = std::max(depth_tri, bb.max_depth_seen);
      Candidate c{};
      for (int k = 0; k < 3; ++k) c.lo[k] = rb.lo[k], c.hi[k] = rb.hi[k];
      c.root = code, c.node = (int)i;
      candidates.push_back(c);
      top_items.push_back(BuildItem{rb, code});
    } else {
      int si = shape_of_node[i];
      int code = make_leaf((uint32_t)si, 1, true);
      if (is_cand[i]) {
        Candidate c{};
        for (int k = 0; k < 3; ++k) c.lo[k] = shape_box[si].lo[k], c.hi[k] = shape_box[si].hi[k];
        c.root = code, c.node = (int)i;
        candidates.push_back(c);
        top_items.push_back(BuildItem{shape_box[si], code});
      } else {
        opaque_items.push_back(BuildItem{shape_box[si], code});
      }
    }
  }
  int root_opaque = kEmpty, root_all = kEmpty;
  int depth_mid = 0, depth_top = 0;
  if (!opaque_items.empty()) {
    Box rb;
    bb.max_depth_seen = 0;
    root_opaque = bb.build_payloads(opaque_items, &rb);
    depth_mid = bb.max_depth_seen;
    top_items.push_back(BuildItem{rb, root_opaque});
  }
  // ---- closest-hit entry (root_all) ------------------------------------------------------------------------------------
  // Shadow queries need the per-SceneNode structure above (opaque tree + one sub-root per transparent candidate).  Closest-hit
  // queries do not: when candidate MESHES exist, a separate tree over ALL triangles serves them.  Joining the candidates'
  // sub-trees under a top node would make every ray walk several overlapping trees (foliage quads are spread through the
  // whole atrium): the unified tree needs 15 % fewer node visits per primary ray on C3 (27.9 vs 33.0, scripts/bvh_sim.cpp)
  // at the price of a second copy of the nodes and leaf-ordered triangles.  NRB_UNIFIED_TREE=0 restores the joined form.
  bool any_cand_mesh = false;
  for (uint32_t i = 0; i < d.n_nodes; ++i)
    any_cand_mesh = any_cand_mesh || (d.nodes[i].shape == NRB_SHAPE_TRIMESH && is_cand[i] && !is_nmap[i]);
  int depth_all = 0;
  if (any_cand_mesh && env_size("NRB_UNIFIED_TREE", 1) != 0) {
    std::vector<BuildItem> items;
    for (uint32_t i = 0; i < d.n_nodes; ++i) {
      const NrbNodeDesc &n = d.nodes[i];
      if (n.shape != NRB_SHAPE_TRIMESH || is_nmap[i]) continue;
      const size_t at = items.size();
      items.resize(at + n.tri_count);
      const uint64_t b0 = node_tri_begin[i];
      parallel_for((size_t)n.tri_count, [&](size_t lo_t, size_t hi_t) {
        for (size_t t = lo_t; t < hi_t; ++t) items[at + t] = BuildItem{tri_box[b0 + t], (int)(b0 + t)};
      });
    }
    std::vector<BuildItem> all_items;
    Box rb;
    bb.max_depth_seen = 0;
    int code = kEmpty;
    int brc = build_set(items, &rb, &code);
    if (brc) return brc;
    depth_all = bb.max_depth_seen;
    all_items.push_back(BuildItem{rb, code});
    for (uint32_t i = 0; i < d.n_nodes; ++i) {
      const NrbNodeDesc &n = d.nodes[i];
      if (n.shape == NRB_SHAPE_TRIMESH || n.shape == NRB_SHAPE_PLANE || is_nmap[i]) continue;
      const int si = shape_of_node[i];
      all_items.push_back(BuildItem{shape_box[si], make_leaf((uint32_t)si, 1, true)});
    }
    bb.max_depth_seen = 0;
    root_all = bb.build_payloads(all_items, &rb);
    depth_top = bb.max_depth_seen;
    H.unified = true;
  } else if (!top_items.empty()) {
    Box rb;
    bb.max_depth_seen = 0;
    root_all = bb.build_payloads(top_items, &rb);
    depth_top = bb.max_depth_seen;
  }
  if (H.unified && depth_all + depth_top + 4 > kStackSize) return fail(NRB_ERR_UNSUPPORTED, "BVH deeper than the traversal stack");
  // one stack slot per level at most (the far child of each two-hit node) + the sentinel
  if (depth_tri + depth_mid + depth_top + 4 > kStackSize)
    return fail(NRB_ERR_UNSUPPORTED, "BVH deeper than the traversal stack");
  if (nmaps.size() > 32) return fail(NRB_ERR_UNSUPPORTED, "more than 32 nodes with a depth-shift (nmap) texture");
  pt.lap("candidate / top trees");
  relayout_bfs(bb.nodes, root_all, root_opaque, candidates, nmaps, bb.dead);
  pt.lap("level-order relayout");

  // leaf-ordered triangle arrays
  H.tris.resize_uninit(bb.tri_order.size());
  H.tri_uvs.resize_uninit(bb.tri_order.size());
  parallel_for(bb.tri_order.size(), [&](size_t lo_t, size_t hi_t) {
    for (size_t k = lo_t; k < hi_t; ++k) H.tris[k] = tris_in[bb.tri_order[k]], H.tri_uvs[k] = uvs_in[bb.tri_order[k]];
  });
  pt.lap("leaf-ordered triangle arrays");
  H.nodes.swap(bb.nodes);
  H.shapes.swap(shapes);
  H.node_info.swap(node_info);
  H.materials.swap(materials);
  H.textures.swap(textures);
  H.lights.swap(lights);
  H.planes.swap(planes);
  H.candidates.swap(candidates);
  H.nmaps.swap(nmaps);
  H.root_all = root_all, H.root_opaque = root_opaque;
  H.n_source_tris = total_tris;
  H.shadow_samples = shadow_samples;
  H.any_refl = any_refl, H.any_refr = any_refr;
  H.depth_tri = depth_tri, H.depth_mid = depth_mid, H.depth_top = depth_top;
  return NRB_OK;
}

// Structural self-check of the built BVH (used by nrb_scene_validate): every triangle sits in exactly one
// leaf, every child box contains what is below it, every node is reachable exactly once.
int check_bvh(const HostScene &H, std::string &why) {
  std::vector<char> tri_seen(H.tris.size(), 0), node_seen(H.nodes.size(), 0);
  struct Item {
    int code;
    Box box;
    bool has_box;
  };
  std::vector<Item> stack;
  if (H.root_all != kEmpty) stack.push_back(Item{H.root_all, Box{}, false});
  if (H.unified) {  // the shadow structure is a second set of trees over the same triangles
    if (H.root_opaque != kEmpty) stack.push_back(Item{H.root_opaque, Box{}, false});
    for (const Candidate &c : H.candidates) stack.push_back(Item{c.root, Box{}, false});
  }
  for (const Candidate &nm : H.nmaps) stack.push_back(Item{nm.root, Box{}, false});
  auto inside = [](const Box &outer, const Box &inner) {
    for (int k = 0; k < 3; ++k)
      if (inner.lo[k] < outer.lo[k] || inner.hi[k] > outer.hi[k]) return false;
    return true;
  };
  while (!stack.empty()) {
    Item it = stack.back();
    stack.pop_back();
    if (it.code >= 0) {
      if ((size_t)it.code >= H.nodes.size()) return why = "child index out of range", 1;
      if (node_seen[it.code]++) return why = "node reachable twice", 1;
      const BvhNode &n = H.nodes[it.code];
      Box b0{{n.n0.x, n.n0.z, n.n2.x}, {n.n0.y, n.n0.w, n.n2.y}}, b1{{n.n1.x, n.n1.z, n.n2.z}, {n.n1.y, n.n1.w, n.n2.w}};
      if (it.has_box && (!inside(it.box, b0) || !inside(it.box, b1))) return why = "child box outside its parent box", 1;
      stack.push_back(Item{n.n3.x, b0, true});
      stack.push_back(Item{n.n3.y, b1, true});
    } else {
      uint32_t code = (uint32_t)~it.code, first = code >> 3, cnt = ((code >> 1) & 3u) + 1u;
      if (code & 1u) {
        if (first >= H.shapes.size()) return why = "shape leaf out of range", 1;
        continue;
      }
      for (uint32_t k = 0; k < cnt; ++k) {
        if (first + k >= H.tris.size()) return why = "triangle leaf out of range", 1;
        if (tri_seen[first + k]++) return why = "triangle in two leaves", 1;
        const Tri &t = H.tris[first + k];
        float v[3][3] = {{t.t0.x, t.t0.y, t.t0.z}, {t.t0.x + t.t1.x, t.t0.y + t.t1.y, t.t0.z + t.t1.z}, {t.t0.x + t.t2.x, t.t0.y + t.t2.y, t.t0.z + t.t2.z}};
        if (it.has_box)
          for (auto &p : v)
            for (int a = 0; a < 3; ++a)
              if (p[a] < it.box.lo[a] || p[a] > it.box.hi[a]) return why = "triangle vertex outside its leaf box", 1;
      }
    }
  }
  for (char c : tri_seen)
    if (!c) return why = "triangle not referenced by any leaf", 1;
  for (char c : node_seen)
    if (!c) return why = "unreachable node", 1;
  return 0;
}

// Device node formats: box centres + half extents (device_types.cuh "node formats").  The half extent is rounded up so
// the device box contains the builder's [lo, hi] box — in fp32 for format 0, to bf16 for format 2.
struct DevNodes {
  HostArray<float> a;  // 16 words per node for formats 0 / 2 (format 2 uses the first 12), 8 words for the grid format 3
};

// Format 3: ONE 16-bit grid over every box of the scene.  A plane at cell q sits at lo + q * cell; boxes are snapped outward
// and widened by one more cell, which covers the kernel's arithmetic (kernels.cu: ray_pre_for<3> folds 2^23 * cell / d into the
// ray constant: at most 0.52 cell of rounding, see check_device_nodes).  Unbounded or absurdly large boxes have no grid.
struct NodeGrid {
  float lo[3], cell[3];
  bool ok;
};
constexpr int kGridMargin = 16;  // cells left free at both ends of the grid
constexpr double kGridCells = 65535.0 - 2.0 * kGridMargin;

NodeGrid make_node_grid(const std::vector<BvhNode> &in) {
  NodeGrid g;
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (const BvhNode &n : in) {
    const float l[3] = {std::min(n.n0.x, n.n1.x), std::min(n.n0.z, n.n1.z), std::min(n.n2.x, n.n2.z)};
    const float h[3] = {std::max(n.n0.y, n.n1.y), std::max(n.n0.w, n.n1.w), std::max(n.n2.y, n.n2.w)};
    for (int k = 0; k < 3; ++k) lo[k] = std::min(lo[k], l[k]), hi[k] = std::max(hi[k], h[k]);
  }
  g.ok = !in.empty();
  for (int k = 0; k < 3; ++k) {
    if (!(lo[k] <= hi[k]) || !(std::fabs(lo[k]) < 1.0e9f) || !(std::fabs(hi[k]) < 1.0e9f)) g.ok = false;
    // A cell is never finer than four float steps of the coordinates themselves (flat axes, scenes far from the origin): the one
    // cell of margin then covers at least what format 0's ulp-scaled padding covers.  (The "- 4": float rounding of cell / lo
    // must not push a box off the grid.)
    const double mag = std::max(std::fabs((double)lo[k]), std::fabs((double)hi[k]));
    g.cell[k] = (float)std::max({((double)hi[k] - (double)lo[k]) / (kGridCells - 4.0), mag * 4.8e-7, 1.0e-20});
    g.lo[k] = (float)((double)lo[k] - (kGridMargin + 2.0) * (double)g.cell[k]);
  }
  return g;
}

// (lo, hi) -> cells, outward plus one; false if the box leaves the grid (cannot happen for boxes make_node_grid has seen)
bool grid_cells(const NodeGrid &g, int k, float lo, float hi, uint32_t &qlo, uint32_t &qhi) {
  const double a = std::floor(((double)lo - (double)g.lo[k]) / (double)g.cell[k]) - 1.0;
  const double b = std::ceil(((double)hi - (double)g.lo[k]) / (double)g.cell[k]) + 1.0;
  if (!(a >= 0.0) || !(b <= 65535.0) || !(a <= b)) return false;
  qlo = (uint32_t)a, qhi = (uint32_t)b;
  return true;
}

struct NodeBoxes {  // one node's two child boxes as (centre, half extent), plus the child codes
  float c[2][3], h[2][3];
  int ch[2];
};

float bf16_round_up(float v) {  // smallest bf16 >= v for v >= 0
  uint32_t u;
  std::memcpy(&u, &v, 4);
  if (u & 0xFFFFu) u = (u | 0xFFFFu) + 1u;
  u &= 0xFFFF0000u;
  float r;
  std::memcpy(&r, &u, 4);
  return r;
}
uint32_t bf16_pair(float lo, float hi) {  // two bf16-exact floats -> one word (lo in the low half)
  uint32_t a, b;
  std::memcpy(&a, &lo, 4);
  std::memcpy(&b, &hi, 4);
  return (a >> 16) | (b & 0xFFFF0000u);
}

NodeBoxes centre_half(const BvhNode &n, int format) {
  const float lo[2][3] = {{n.n0.x, n.n0.z, n.n2.x}, {n.n1.x, n.n1.z, n.n2.z}};
  const float hi[2][3] = {{n.n0.y, n.n0.w, n.n2.y}, {n.n1.y, n.n1.w, n.n2.w}};
  NodeBoxes o;
  for (int b = 0; b < 2; ++b)
    for (int k = 0; k < 3; ++k) {
      float l = std::max(lo[b][k], -1.0e30f), u = std::min(hi[b][k], 1.0e30f);
      float cc = 0.5f * l + 0.5f * u;
      float hh = std::max(u - cc, cc - l);
      hh = hh * 1.0000005f + std::fabs(cc) * 1.2e-7f + 1e-30f;  // covers the rounding of cc and of u - cc
      if (format == 2) hh = bf16_round_up(hh);
      o.c[b][k] = cc;
      o.h[b][k] = hh;
    }
  o.ch[0] = n.n3.x, o.ch[1] = n.n3.y;
  return o;
}

// How much the grid inflates the tree: mean over all child boxes of sqrt(snapped half area / builder's half area), each capped at
// 100.  One far-away triangle stretches the grid over a huge extent, a cell becomes larger than the geometry that matters and every
// snapped box overlaps its neighbours — still correct, but the traversal degenerates; such scenes keep float boxes (upload_scene).
// (Per box, not a ratio of sums: the sums are dominated by the few boxes at the top of the tree, which no grid inflates.)
double grid_inflation(const std::vector<BvhNode> &in, const NodeGrid &g) {
  std::mutex mu;
  double sum_r = 0.0;
  uint64_t n_boxes = 0;
  parallel_for(in.size(), [&](size_t lo_i, size_t hi_i) {
    double r = 0.0;
    uint64_t nb = 0;
    for (size_t i = lo_i; i < hi_i; ++i) {
      const BvhNode &n = in[i];
      const float lo[2][3] = {{n.n0.x, n.n0.z, n.n2.x}, {n.n1.x, n.n1.z, n.n2.z}};
      const float hi[2][3] = {{n.n0.y, n.n0.w, n.n2.y}, {n.n1.y, n.n1.w, n.n2.w}};
      for (int b = 0; b < 2; ++b) {
        double e[3], eq[3];
        for (int k = 0; k < 3; ++k) {
          uint32_t ql = 0, qh = 65535;
          grid_cells(g, k, lo[b][k], hi[b][k], ql, qh);
          e[k] = (double)hi[b][k] - (double)lo[b][k];
          eq[k] = (double)(qh - ql) * (double)g.cell[k];
        }
        const double o = e[0] * e[1] + e[1] * e[2] + e[2] * e[0], q = eq[0] * eq[1] + eq[1] * eq[2] + eq[2] * eq[0];
        if (!(o > 0.0)) continue;  // a point or an axis-parallel segment
        r += std::min(100.0, std::sqrt(q / o));
        ++nb;
      }
    }
    std::lock_guard<std::mutex> lk(mu);
    sum_r += r, n_boxes += nb;
  });
  return n_boxes ? sum_r / (double)n_boxes : 1.0;
}

DevNodes to_device_nodes(const std::vector<BvhNode> &in, int format, const NodeGrid *grid = nullptr, bool *grid_failed = nullptr) {
  DevNodes d;
  const size_t wa = format >= 3 ? 8 : 16;
  d.a.resize_uninit(in.size() * wa);
  auto put_int = [](float *dst, int v) { std::memcpy(dst, &v, 4); };
  auto put_u32 = [](float *dst, uint32_t v) { std::memcpy(dst, &v, 4); };
  std::atomic<bool> off_grid{false};
  parallel_for(in.size(), [&](size_t lo_i, size_t hi_i) {
    for (size_t i = lo_i; i < hi_i; ++i) {
      float *o = &d.a[i * wa];
      if (format >= 3) {
        const BvhNode &n = in[i];
        const float lo[2][3] = {{n.n0.x, n.n0.z, n.n2.x}, {n.n1.x, n.n1.z, n.n2.z}};
        const float hi[2][3] = {{n.n0.y, n.n0.w, n.n2.y}, {n.n1.y, n.n1.w, n.n2.w}};
        for (int b = 0; b < 2; ++b)
          for (int k = 0; k < 3; ++k) {
            uint32_t ql = 65535u, qh = 0u;
            if (!grid_cells(*grid, k, lo[b][k], hi[b][k], ql, qh)) off_grid = true;
            put_u32(o + 3 * b + k, (qh << 16) | ql);
          }
        put_int(o + 6, n.n3.x), put_int(o + 7, n.n3.y);
        continue;
      }
      const NodeBoxes nb = centre_half(in[i], format);
      std::memset(o, 0, wa * sizeof(float));
      if (format == 0) {
        const float rec[12] = {nb.c[0][0], nb.c[0][1], nb.c[0][2], nb.c[1][0], nb.c[1][1], nb.c[1][2],
                               nb.h[0][0], nb.h[0][1], nb.h[0][2], nb.h[1][0], nb.h[1][1], nb.h[1][2]};
        std::memcpy(o, rec, sizeof(rec));
        put_int(o + 12, nb.ch[0]), put_int(o + 13, nb.ch[1]);
      } else {
        const float rec[6] = {nb.c[0][0], nb.c[0][1], nb.c[1][0], nb.c[1][1], nb.c[0][2], nb.c[1][2]};
        std::memcpy(o, rec, sizeof(rec));
        put_u32(o + 6, bf16_pair(nb.h[0][0], nb.h[0][1]));
        put_u32(o + 7, bf16_pair(nb.h[1][0], nb.h[1][1]));
        float *ob = o + 8;
        put_u32(ob, bf16_pair(nb.h[0][2], nb.h[1][2]));
        put_int(ob + 1, nb.ch[0]), put_int(ob + 2, nb.ch[1]);
      }
    }
  });
  if (grid_failed) *grid_failed = off_grid.load();
  return d;
}

// Decodes node i of a device array back into (centre, half extent) exactly as the kernel reads it.
NodeBoxes decode_device_node(const DevNodes &d, size_t i, int format) {
  NodeBoxes o;
  auto get_int = [](const float *src) { int v; std::memcpy(&v, src, 4); return v; };
  auto bf_lo = [](const float *src) { uint32_t u; std::memcpy(&u, src, 4); u <<= 16; float r; std::memcpy(&r, &u, 4); return r; };
  auto bf_hi = [](const float *src) { float r; std::memcpy(&r, src, 4); return r; };  // the kernel does not mask the low half away
  if (format == 0) {
    const float *p = &d.a[i * 16];
    for (int k = 0; k < 3; ++k) o.c[0][k] = p[k], o.c[1][k] = p[3 + k], o.h[0][k] = p[6 + k], o.h[1][k] = p[9 + k];
    o.ch[0] = get_int(p + 12), o.ch[1] = get_int(p + 13);
  } else {
    const float *p = &d.a[i * 16], *q = p + 8;
    o.c[0][0] = p[0], o.c[0][1] = p[1], o.c[1][0] = p[2], o.c[1][1] = p[3], o.c[0][2] = p[4], o.c[1][2] = p[5];
    o.h[0][0] = bf_lo(p + 6), o.h[0][1] = bf_hi(p + 6), o.h[1][0] = bf_lo(p + 7), o.h[1][1] = bf_hi(p + 7);
    o.h[0][2] = bf_lo(q), o.h[1][2] = bf_hi(q);
    o.ch[0] = get_int(q + 1), o.ch[1] = get_int(q + 2);
  }
  return o;
}

// The device boxes (centre -+ half extent, evaluated in f32 as the kernel sees them) must contain the builder's boxes —
// for EVERY node format, whichever one the kernels were compiled for.
int check_device_nodes(const HostScene &H, std::string &why) {
  for (int format = 0; format <= 2; format += 2) {
    const DevNodes dev = to_device_nodes(H.nodes, format);
    for (size_t i = 0; i < H.nodes.size(); ++i) {
      const BvhNode &n = H.nodes[i];
      const NodeBoxes d = decode_device_node(dev, i, format);
      const float lo[2][3] = {{n.n0.x, n.n0.z, n.n2.x}, {n.n1.x, n.n1.z, n.n2.z}};
      const float hi[2][3] = {{n.n0.y, n.n0.w, n.n2.y}, {n.n1.y, n.n1.w, n.n2.w}};
      for (int b = 0; b < 2; ++b)
        for (int k = 0; k < 3; ++k) {
          const float l = std::max(lo[b][k], -1.0e30f), u = std::min(hi[b][k], 1.0e30f);
          if (!(d.c[b][k] - d.h[b][k] <= l) || !(d.c[b][k] + d.h[b][k] >= u) || !(d.h[b][k] >= 0.0f))
            return why = "device box (centre / half extent, node format " + std::to_string(format) + ") does not contain the builder's box", 1;
        }
      if (d.ch[0] != n.n3.x || d.ch[1] != n.n3.y) return why = "device node lost its child codes", 1;
    }
  }
  // format 3: every plane at least half a cell outside the builder's box (the kernel's rounding budget), on the grid
  const NodeGrid g = make_node_grid(H.nodes);
  if (g.ok) {
    bool failed = false;
    const DevNodes dev = to_device_nodes(H.nodes, 3, &g, &failed);
    if (failed) return why = "node box off the 16-bit grid", 1;
    for (size_t i = 0; i < H.nodes.size(); ++i) {
      const BvhNode &n = H.nodes[i];
      const float lo[2][3] = {{n.n0.x, n.n0.z, n.n2.x}, {n.n1.x, n.n1.z, n.n2.z}};
      const float hi[2][3] = {{n.n0.y, n.n0.w, n.n2.y}, {n.n1.y, n.n1.w, n.n2.w}};
      uint32_t w[8];
      std::memcpy(w, &dev.a[i * 8], 32);
      for (int b = 0; b < 2; ++b)
        for (int k = 0; k < 3; ++k) {
          const double pl = (double)g.lo[k] + (double)(w[3 * b + k] & 0xFFFFu) * (double)g.cell[k];
          const double ph = (double)g.lo[k] + (double)(w[3 * b + k] >> 16) * (double)g.cell[k];
          if (!(pl + 0.75 * (double)g.cell[k] <= (double)lo[b][k]) || !(ph - 0.75 * (double)g.cell[k] >= (double)hi[b][k]))
            return why = "grid box (node format 3) does not contain the builder's box with its margin", 1;
        }
      if ((int)w[6] != n.n3.x || (int)w[7] != n.n3.y) return why = "device node (format 3) lost its child codes", 1;
    }
  }
  return 0;
}

// NRB_DUMP_BVH=<file>: builder experiments (scripts/bvh_sim.cpp) — nodes + leaf-ordered triangles + candidates as built
void dump_bvh(const HostScene &H) {
  const char *path = getenv("NRB_DUMP_BVH");
  if (!path) return;
  if (FILE *f = fopen(path, "wb")) {
    uint64_t hdr[4] = {H.nodes.size(), H.tris.size(), (uint64_t)(uint32_t)H.root_all, (uint64_t)(uint32_t)H.root_opaque};
    fwrite(hdr, sizeof(hdr), 1, f);
    fwrite(H.nodes.data(), sizeof(BvhNode), H.nodes.size(), f);
    fwrite(H.tris.data(), sizeof(Tri), H.tris.size(), f);
    uint64_t nc = H.candidates.size();
    fwrite(&nc, sizeof(nc), 1, f);
    fwrite(H.candidates.data(), sizeof(Candidate), H.candidates.size(), f);
    fclose(f);
  }
}

int upload_scene(const NrbSceneDesc &d, const HostScene &H, NrbScene &S) {
  // Node format (device_types.cuh): scenes that fit L2 keep full fp32 boxes; scenes larger than L2 — where rays diverge and
  // every lane fetches its own cache line — use the 48-byte-per-visit bf16 form.  Depth-shift (nmap) scenes run the general
  // kernel, compiled for format 0 only.  NRB_NODE_FORMAT=0|2 overrides (experiments).
  const uint64_t geom_bytes = H.nodes.size() * 64ull + H.tris.size() * sizeof(Tri);
  const bool grid_capable = H.shapes.empty() && H.nmaps.empty() && !H.nodes.empty();  // formats 3 / 4: mesh-only scenes
  int nfmt = (geom_bytes > (96ull << 20) && H.nmaps.empty()) ? (grid_capable ? 4 : 2) : 0;
  if (const char *e = getenv("NRB_NODE_FORMAT")) {
    const int f = atoi(e);
    if ((f == 0 || f == 2) && (H.nmaps.empty() || f == 0)) nfmt = f;
    if ((f == 3 || f == 4) && grid_capable) nfmt = f;
  }
  NodeGrid grid = {};
  if (nfmt >= 3) {
    grid = make_node_grid(H.nodes);
    // the grid must resolve the geometry: snapped boxes on average at most 1.15x the builder's, linearly (C3 1.011, C4 1.007;
    // NRB_GRID_MAX_INFLATION)
    const char *lim = getenv("NRB_GRID_MAX_INFLATION");
    const double infl = grid.ok ? grid_inflation(H.nodes, grid) : 0.0;
    if (getenv("NRB_BUILD_TIMES")) fprintf(stderr, "[nrb] 16-bit grid: mean linear inflation of the child boxes %.3fx\n", infl);
    if (!grid.ok || infl > (lim ? atof(lim) : 1.15)) nfmt = nfmt == 4 ? 2 : 0;
  }
  {
    bool off_grid = false;
    const DevNodes dn = to_device_nodes(H.nodes, nfmt, &grid, &off_grid);
    if (off_grid) return fail(NRB_ERR_INTERNAL, "internal BVH invariant violated: node box off the 16-bit grid");
    CU(upload(S.d_nodes, dn.a));
  }
  CU(upload(S.d_tris, H.tris));
  CU(upload(S.d_tri_uvs, H.tri_uvs));
  CU(upload(S.d_shapes, H.shapes));
  CU(upload(S.d_node_info, H.node_info));
  CU(upload(S.d_materials, H.materials));
  CU(upload(S.d_textures, H.textures));
  CU(upload(S.d_lights, H.lights));
  CU(upload(S.d_planes, H.planes));
  CU(upload(S.d_candidates, H.candidates));
  CU(upload(S.d_nmaps, H.nmaps));
  {
    size_t tb = std::max<size_t>(d.n_texels * 16, 16);
    CU(S.d_texels.ensure(tb));
    if (d.n_texels) CU(cudaMemcpy(S.d_texels.p, d.texels, d.n_texels * 16, cudaMemcpyHostToDevice));
  }
  SceneView &v = S.view;
  v.nodes = S.d_nodes.p;
  v.node_format = nfmt;
  for (int k = 0; k < 3; ++k) v.grid_lo[k] = grid.lo[k], v.grid_cell[k] = grid.cell[k];
  v.tris = S.d_tris.as<Tri>();
  v.tri_uvs = S.d_tri_uvs.as<TriUV>();
  v.shapes = S.d_shapes.as<Shape>();
  v.node_info = S.d_node_info.as<NodeInfo>();
  v.materials = S.d_materials.as<Material>();
  v.textures = S.d_textures.as<Texture>();
  v.texels = S.d_texels.as<float4>();
  v.lights = S.d_lights.as<Light>();
  v.planes = S.d_planes.as<int>();
  v.candidates = S.d_candidates.as<Candidate>();
  v.nmaps = S.d_nmaps.as<Candidate>();
  v.n_nmap = (int)H.nmaps.size();
  v.root_all = H.root_all;
  v.root_opaque = H.root_opaque;
  v.n_planes = (int)H.planes.size();
  v.n_candidates = (int)H.candidates.size();
  v.n_lights = (int)H.lights.size();
  v.shadow_samples = H.shadow_samples;
  for (int k = 0; k < 3; ++k) v.background[k] = d.background[k];
  S.has_nmap = !H.nmaps.empty();
  S.has_shapes = !H.shapes.empty() || S.has_nmap;  // the general (HAS_SHAPES) kernel variants carry the nmap code
  S.child_factor = (H.any_refl ? 1 : 0) + (H.any_refr ? 1 : 0);
  S.refl_chain_max = H.refl_chain_max;
  S.n_bvh_nodes = H.nodes.size();
  S.n_tris = H.n_source_tris;
  S.node_bytes = nfmt >= 3 ? 32 : 64;  // stride; format 2 reads 48 of its 64 per visit
  S.scene_bytes = H.nodes.size() * S.node_bytes + H.tris.size() * (sizeof(Tri) + sizeof(TriUV)) +
                  H.shapes.size() * sizeof(Shape) + d.n_texels * 16;
  return NRB_OK;
}

// ---------------------------------------------------------------------------------------------
// scene::render — wavefront driver
// ---------------------------------------------------------------------------------------------
int make_frame_params(const NrbCamera &cam, const NrbTileSet *tiles, FrameParams &fp) {
  if (cam.ray_per_pixel == 0) return fail(NRB_ERR_INVALID_ARG, "ray_per_pixel must be > 0 (assert at src/scene.rs:37)");
  if (cam.width == 0 || cam.height == 0) return fail(NRB_ERR_INVALID_ARG, "resolution must be non-zero");
  uint64_t samples = (uint64_t)cam.width * cam.height * cam.ray_per_pixel;
  if (samples >= (1ull << 32)) return fail(NRB_ERR_INVALID_ARG, "width*height*ray_per_pixel must be < 2^32");
  // sample slots count PADDED 16x16 tiles (ragged frames have more slots than samples) and are 32-bit on the device
  if ((uint64_t)((cam.width + NRB_TILE - 1) / NRB_TILE) * ((cam.height + NRB_TILE - 1) / NRB_TILE) * (NRB_TILE * NRB_TILE) * cam.ray_per_pixel >= (1ull << 32))
    return fail(NRB_ERR_INVALID_ARG, "tiles_x*tiles_y*256*ray_per_pixel must be < 2^32");
  std::memset(&fp, 0, sizeof(fp));
  fp.width = cam.width, fp.height = cam.height, fp.spp = cam.ray_per_pixel;
  fp.max_depth = cam.max_depth ? cam.max_depth : 64u;
  fp.tiles_x = (cam.width + NRB_TILE - 1) / NRB_TILE;
  fp.tiles_y = (cam.height + NRB_TILE - 1) / NRB_TILE;
  uint32_t n_tiles = fp.tiles_x * fp.tiles_y;
  if (tiles) {
    if (tiles->stride == 0 || tiles->first >= tiles->stride) return fail(NRB_ERR_INVALID_ARG, "tile set: need first < stride, stride > 0");
    fp.tile_first = tiles->first, fp.tile_stride = tiles->stride;
    fp.n_local_tiles = tiles->first < n_tiles ? (n_tiles - tiles->first + tiles->stride - 1) / tiles->stride : 0;
    fp.packed = 1;
  } else {
    fp.tile_first = 0, fp.tile_stride = 1, fp.n_local_tiles = n_tiles, fp.packed = 0;
  }
  fp.window = (float)cam.window_width;
  fp.inv_w = 1.0f / (float)cam.width;
  fp.inv_h = 1.0f / (float)cam.height;
  // h = M * (x, y, -1, 1) = A x + B y + C (column-major M); direction ~ h.xyz - eye * h.w, evaluated
  // in f64 here so the device adds three f32 vectors and never cancels against eye.
  const double *M = cam.projection;
  double A[4], B[4], Cc[4];
  for (int r = 0; r < 4; ++r) A[r] = M[r], B[r] = M[4 + r], Cc[r] = -M[8 + r] + M[12 + r];
  if (A[3] == 0.0 && B[3] == 0.0 && Cc[3] == 0.0)
    return fail(NRB_ERR_INVALID_ARG, "projection maps the near plane to w == 0 (from_homogeneous().unwrap() panics: src/scene.rs:85)");
  // scale so the vectors are O(1) in f32
  double sc = 0.0;
  for (int k = 0; k < 3; ++k) sc = std::max(sc, std::fabs(Cc[k] - cam.eye[k] * Cc[3]));
  for (int k = 0; k < 3; ++k) sc = std::max(sc, std::max(std::fabs(A[k] - cam.eye[k] * A[3]), std::fabs(B[k] - cam.eye[k] * B[3])));
  if (!(sc > 0.0) || !std::isfinite(sc)) return fail(NRB_ERR_INVALID_ARG, "degenerate projection matrix");
  for (int k = 0; k < 3; ++k) {
    fp.eye[k] = (float)cam.eye[k];
    fp.dx[k] = (float)((A[k] - cam.eye[k] * A[3]) / sc);
    fp.dy[k] = (float)((B[k] - cam.eye[k] * B[3]) / sc);
    fp.d0[k] = (float)((Cc[k] - cam.eye[k] * Cc[3]) / sc);
  }
  double ws = std::max(std::fabs(A[3]), std::max(std::fabs(B[3]), std::fabs(Cc[3])));
  fp.wx = (float)(A[3] / ws), fp.wy = (float)(B[3] / ws), fp.w0 = (float)(Cc[3] / ws);
  fp.div_spp = make_fastdiv(fp.spp), fp.div_tiles_x = make_fastdiv(fp.tiles_x), fp.div_width = make_fastdiv(fp.width);
  fp.div_tile_stride = make_fastdiv(fp.tile_stride);
  fp.seed_lo = (uint32_t)cam.seed;
  fp.seed_hi = (uint32_t)(cam.seed >> 32);
  return NRB_OK;
}

cudaError_t ensure_ray_queue(NrbScene &S, int qi, uint32_t cap) {
  if (cap <= S.q_cap[qi]) return cudaSuccess;
  cap = std::max<uint32_t>(cap, 1024);
  for (int c = 0; c < 3; ++c) {
    cudaError_t e = S.d_q[qi][c].ensure((size_t)cap * 16);
    if (e != cudaSuccess) return e;
  }
  S.q_cap[qi] = cap;
  return cudaSuccess;
}

RayQueue ray_queue(NrbScene &S, int qi) {
  return RayQueue{S.d_q[qi][0].as<float4>(), S.d_q[qi][1].as<float4>(), S.d_q[qi][2].as<float4>(), S.q_cap[qi]};
}

cudaEvent_t get_event(NrbScene &S, size_t &used) {
  if (used == S.events.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    S.events.push_back(e);
  }
  return S.events[used++];
}

// Renders into S.d_accum and resolves to `d_out` (device; float rgb or u8 rgb).
// `to_image`: with a tile set, resolve this rank's tiles into the row-major image `d_out` (possibly peer memory)
// instead of the packed tile buffer.
// `early_h_out` (nrb_render with a pinned, device-mapped destination; `early_d_out` is the same memory as the device
// sees it): when the frame reaches its tail phase, the image as it stands is resolved and its device->host copy starts on
// S.copy_stream, overlapping the tail; the final resolve is replaced by a kernel that stores the pixels the tail
// changed straight into the host image once that copy has landed.  *early_used tells the caller which happened.
int render_device(NrbScene &S, const NrbCamera &cam, const NrbTileSet *tiles, float *d_out, uint8_t *d_out8,
                  uint32_t *n_local_tiles, NrbStats *stats, bool to_image = false, float *early_h_out = nullptr,
                  float *early_d_out = nullptr, bool *early_used = nullptr, void *segments_h_out = nullptr) {
  CU(cudaSetDevice(S.device));
  FrameParams fp;
  int rc = make_frame_params(cam, tiles, fp);
  if (rc) return rc;
  if (n_local_tiles) *n_local_tiles = fp.n_local_tiles;
  cudaStream_t st = S.stream;
  const uint32_t n_acc = fp.packed ? fp.n_local_tiles * NRB_TILE * NRB_TILE : fp.width * fp.height;
  CU(S.d_accum.ensure(std::max<size_t>((size_t)n_acc * 16, 16)));
  float4 *accum = S.d_accum.as<float4>();
  Counters *dc = S.d_counters.as<Counters>();
  uint32_t launches = 0, waves = 0;
  size_t wave_counts_used = 0;
  size_t ev_used = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> trace_spans, shade_spans, tail_spans;
  const bool dump = getenv("NRB_DUMP_WAVES") != nullptr;

  {
    uint64_t n_batches = (fp.n_local_tiles + std::max<uint64_t>(1, env_size("NRB_BATCH_SLOTS", kBatchSlots) / (NRB_TILE * NRB_TILE * fp.spp)) - 1) /
                         std::max<uint64_t>(1, env_size("NRB_BATCH_SLOTS", kBatchSlots) / (NRB_TILE * NRB_TILE * fp.spp));
    size_t need = (size_t)(n_batches + 1) * ((size_t)fp.max_depth + 2);
    if (need > S.h_wave_cap) {
      CU(cudaStreamSynchronize(st));
      if (S.h_wave_counts) cudaFreeHost(S.h_wave_counts);
      S.h_wave_counts = nullptr;
      CU(cudaHostAlloc((void **)&S.h_wave_counts, need * sizeof(uint32_t), cudaHostAllocDefault));
      S.h_wave_cap = need;
    }
  }
  CU(cudaEventRecord(S.ev_begin, st));
  CU(cudaMemsetAsync(accum, 0, (size_t)n_acc * 16, st));
  CU(cudaMemsetAsync(dc, 0, sizeof(Counters), st));

  const uint32_t per_tile = NRB_TILE * NRB_TILE * fp.spp;
  const uint64_t batch_slots_req = env_size("NRB_BATCH_SLOTS", kBatchSlots);
  const uint32_t tiles_per_batch = (uint32_t)std::max<uint64_t>(1, batch_slots_req / per_tile);
  const uint64_t shadow_cap_req = env_size("NRB_SHADOW_CAP", kShadowCap);
  const uint64_t mem_ceiling = env_size("NRB_QUEUE_BYTES", 64ull << 30);
  const uint32_t S_total = (uint32_t)S.view.shadow_samples;
  const uint64_t tail_threshold = env_size("NRB_TAIL_RAYS", 4u << 20);
  const uint32_t tail_min_wave = (uint32_t)env_size("NRB_TAIL_MIN_WAVE", 2);
  // Trace-kernel knobs (profiles/README.md has the measurements).  Dynamic fetch — a warp pauses to refill its idle lanes
  // once fewer than N lanes are still traversing — loses 8-25 % on coherent frames (a refill round costs more than
  // the idle lanes) and gains ~7 % where neighbouring rays diverge; single-packet fetches balance such frames better
  // (-4 %).  "Fine geometry" = more triangles than half the pixels, i.e. triangles smaller than the pixel grid
  // (hairball: 2.88 M triangles behind 2.07 M pixels): there the secondary / shadow queues use both.
  // opaque any-hit phase of shadow rays walked from the light end (rays of one point light leave together)
  const int reverse_shadow = (int)env_size("NRB_REVERSE_SHADOW", 1);
  const bool fine_geometry = S.n_tris > (uint64_t)fp.width * fp.height / 2;
  const int refill_primary = (int)env_size("NRB_REFILL_PRIMARY", 0);
  const int refill_rays = (int)env_size("NRB_REFILL_RAYS", fine_geometry ? 20 : 0);
  const int refill_shadow = (int)env_size("NRB_REFILL_SHADOW", fine_geometry ? 20 : 0);
  const uint32_t small_queue = (uint32_t)env_size("NRB_SMALL_QUEUE", fine_geometry ? 0xFFFFFFFFu : kSmallQueue);
  uint64_t primary = 0;

  if (!S.count_stream) CU(cudaStreamCreateWithFlags(&S.count_stream, cudaStreamNonBlocking));
  cudaEvent_t last_count_ev = nullptr;
  const size_t wc_len = (size_t)fp.max_depth + 3;
  CU(S.d_wave.ensure(wc_len * sizeof(WaveCounters)));
  WaveCounters *wc = S.d_wave.as<WaveCounters>();
  const RayQueue no_queue{nullptr, nullptr, nullptr, 0};

  for (uint64_t tile_lo = 0; tile_lo < fp.n_local_tiles; tile_lo += tiles_per_batch) {
    uint64_t tile_hi = std::min<uint64_t>(fp.n_local_tiles, tile_lo + tiles_per_batch);
    const uint32_t slot_lo = (uint32_t)(tile_lo * per_tile), n_slots = (uint32_t)((tile_hi - tile_lo) * per_tile);
    // exact number of primary rays of this batch (samples of pixels inside the image): no sync needed
    uint64_t n0 = 0;
    for (uint64_t lt = tile_lo; lt < tile_hi; ++lt) {
      uint32_t tile = fp.tile_first + (uint32_t)lt * fp.tile_stride;
      uint32_t ty = tile / fp.tiles_x, tx = tile - ty * fp.tiles_x;
      uint32_t wpx = std::min<uint32_t>(NRB_TILE, fp.width - tx * NRB_TILE), hpx = std::min<uint32_t>(NRB_TILE, fp.height - ty * NRB_TILE);
      n0 += (uint64_t)wpx * hpx * fp.spp;
    }
    primary += n0;
    if (last_count_ev) CU(cudaStreamWaitEvent(st, last_count_ev, 0));  // the side stream has read the previous batch's counters
    CU(cudaMemsetAsync(wc, 0, wc_len * sizeof(WaveCounters), st));  // the only counter reset of the batch

    // Pipelined waves.  Wave k >= 1 consumes queue k%2 holding n_k = wc[k].n_rays rays (written by the
    // shade of wave k-1); wave 0 generates its rays from the sample slots.  Every kernel reads its
    // counts from device memory, so wave k is enqueued as soon as the host knows n_{k-1}
    // (n_k <= child_factor * n_{k-1}) — i.e. while wave k-1 is still running.
    // Per wave: ONE trace launch (shadow rays of wave k-1 + closest hits of wave k) and ONE shade launch.
    std::vector<cudaEvent_t> count_ev;
    const size_t wave_base = wave_counts_used;
    uint64_t known_prev = n0;
    int pending_shadow = -1;  // wave whose shadow queue has not been traced yet
    auto trace_span = [&](bool primary, RayQueue q, WaveCounters *wcc, WaveCounters *wcs) -> int {
      ShadowQueue sq{S.d_sq[0].as<float4>(), S.d_sq[1].as<float4>(), S.d_sq[2].as<float4>(), S.sq_cap};
      cudaEvent_t e0 = get_event(S, ev_used), e1 = get_event(S, ev_used);
      CU(cudaEventRecord(e0, st));
      if (S.has_nmap)
        launch_trace_general(S.view, fp, primary, q, S.d_hits.as<float4>(), wcc, slot_lo, n_slots, sq, accum, wcs, S.sm_count * 3, st);
      else
        launch_trace(S.view, S.has_shapes, fp, primary, q, S.d_hits.as<float4>(), wcc, slot_lo, n_slots, sq, accum, wcs,
                   TraceOpts{primary ? refill_primary : refill_rays, refill_shadow, reverse_shadow, small_queue}, S.grid_trace, st);
      CU(cudaEventRecord(e1, st));
      trace_spans.emplace_back(e0, e1);
      ++launches;
      return NRB_OK;
    };
    for (uint32_t k = 0; k < fp.max_depth; ++k) {
      const int cur = (int)(k & 1u);
      uint64_t bound;
      if (k == 0) {
        bound = n_slots;  // hit records are indexed by slot in wave 0
      } else {
        if (k >= 2) {
          CU(cudaEventSynchronize(count_ev[k - 1]));
          known_prev = S.h_wave_counts[wave_base + k - 1];
        }
        if (known_prev == 0 || S.child_factor == 0) break;
        bound = known_prev * (uint64_t)S.child_factor;
        // Tail launches spill the refraction ray of every hit that wants BOTH children while the lane follows the reflection.
        // A chain reflects at most min(levels left, refl_chain_max) times, which bounds the spill queue exactly; scenes where
        // that bound is too large for a queue (mirror-glass with tiny attenuation) stay on the wave path, whose queues are
        // sized per wave and cannot overflow.
        const uint64_t spill_limit = env_size("NRB_SPILL_CAP", 64u << 20);
        auto spill_factor = [&](uint32_t level) -> uint64_t {
          return S.child_factor < 2 ? 0ull : std::min<uint64_t>(fp.max_depth - level, S.refl_chain_max);
        };
        if (k >= tail_min_wave && bound <= tail_threshold && bound * spill_factor(k) <= spill_limit) {
          // (every ray of the tail phase descends from a ray of wave k-1, so at most known_prev pixels can still change;
          //  with more than that the patch would rival the image itself)
          if (early_h_out && early_used && !*early_used && fp.n_local_tiles <= tiles_per_batch && d_out && !fp.packed &&
              known_prev <= (uint64_t)n_acc) {
            // everything but the tail phase is in the accumulator: send the image now, patch it afterwards
            launch_resolve(accum, n_acc, fp.spp, d_out, st);
            ++launches;
            CU(cudaEventRecord(S.ev_early, st));
            CU(cudaStreamWaitEvent(S.copy_stream, S.ev_early, 0));
            CU(cudaMemcpyAsync(early_h_out, d_out, (size_t)n_acc * 3 * sizeof(float), cudaMemcpyDeviceToHost, S.copy_stream));
            CU(cudaEventRecord(S.ev_copied, S.copy_stream));
            *early_used = true;
          }
          // ---- tail: every lane follows its own ray chain to the end (see tail_kernel) -----------
          // The chains append their shadow rays behind the ones the last shade left untraced (same queue,
          // same counter); ONE shadow launch per tail launch traces them all.
          for (uint32_t t = k; t < fp.max_depth; ++t) {
            const int tc = (int)(t & 1u);
            if (wave_base + t + 1 >= S.h_wave_cap) return fail(NRB_ERR_INVALID_ARG, "max_depth too large for the wave-count buffer");
            if (bound * spill_factor(t) >= (1ull << 32)) return fail(NRB_ERR_QUEUE_OVERFLOW, "tail spill queue would exceed 2^32 entries");
            CU(ensure_ray_queue(S, 1 - tc, (uint32_t)std::max<uint64_t>(bound * spill_factor(t), 1024)));  // spill queue, exact bound
            if (S_total && S.sq_cap < 65536) {
              if (pending_shadow >= 0) {  // the buffer is about to move: trace what it holds first
                rc = trace_span(false, no_queue, nullptr, &wc[pending_shadow]);
                if (rc) return rc;
                pending_shadow = -1;
              }
              CU(cudaStreamSynchronize(st));
              for (int c = 0; c < 3; ++c) CU(S.d_sq[c].ensure((size_t)65536 * 16));
              S.sq_cap = 65536;
            }
            const int sh_wave = pending_shadow >= 0 ? pending_shadow : (int)t;  // counter the chains' shadow rays go to
            ShadowQueue sq{S.d_sq[0].as<float4>(), S.d_sq[1].as<float4>(), S.d_sq[2].as<float4>(), S.sq_cap};
            cudaEvent_t e0 = get_event(S, ev_used), e1 = get_event(S, ev_used);
            CU(cudaEventRecord(e0, st));
            launch_tail(S.view, S.has_shapes, fp, ray_queue(S, tc), &wc[t], ray_queue(S, 1 - tc), sq, dc, accum, &wc[sh_wave],
                        S.grid_tail, st);
            CU(cudaEventRecord(e1, st));
            tail_spans.emplace_back(e0, e1);
            ++launches;
            if (S_total) {
              rc = trace_span(false, no_queue, nullptr, &wc[sh_wave]);  // pending shadow rays + the chains' shadow rays
              if (rc) return rc;
            }
            pending_shadow = -1;
            CU(cudaMemcpyAsync(&S.h_wave_counts[wave_base + t], &wc[t].n_rays, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            wave_counts_used = std::max(wave_counts_used, wave_base + t + 1);
            if (S.child_factor < 2) break;  // a hit never spawns both children: nothing can be spilled, no sync needed
            // rays spilled by hits that spawned both children start the next tail launch
            CU(cudaMemcpyAsync(&S.h_wave_counts[wave_base + t + 1], &wc[t + 1].n_rays, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            uint32_t spilled = S.h_wave_counts[wave_base + t + 1];
            if (spilled == 0) break;
            bound = spilled;
          }
          break;
        }
      }
      if (wave_base + k >= S.h_wave_cap) return fail(NRB_ERR_INVALID_ARG, "max_depth too large for the wave-count buffer");
      if (k == 0) {
        S.h_wave_counts[wave_base] = (uint32_t)n0;
        count_ev.push_back(nullptr);
      } else {
        // wc[k].n_rays is final once the shade of wave k-1 (already enqueued) has run: read it back on the side stream
        cudaEvent_t ready = get_event(S, ev_used), ev = get_event(S, ev_used);
        CU(cudaEventRecord(ready, st));
        CU(cudaStreamWaitEvent(S.count_stream, ready, 0));
        CU(cudaMemcpyAsync(&S.h_wave_counts[wave_base + k], &wc[k].n_rays, sizeof(uint32_t), cudaMemcpyDeviceToHost, S.count_stream));
        CU(cudaEventRecord(ev, S.count_stream));
        count_ev.push_back(ev);
        last_count_ev = ev;
      }
      ++wave_counts_used;
      const uint64_t emitters = (k == 0) ? n0 : bound;  // rays that can spawn children / shadow rays
      uint64_t next_need = emitters * (uint64_t)S.child_factor;
      if (next_need * 48 > mem_ceiling || next_need >= (1ull << 32))
        return fail(NRB_ERR_QUEUE_OVERFLOW, "secondary-ray queue would exceed NRB_QUEUE_BYTES");
      CU(ensure_ray_queue(S, 1 - cur, (uint32_t)next_need));
      if (bound > S.hits_cap) {
        CU(cudaStreamSynchronize(st));  // growing frees the old buffer: drain first (first frames only)
        CU(S.d_hits.ensure((size_t)bound * 16));
        S.hits_cap = (uint32_t)bound;
      }
      uint32_t n_upper = (uint32_t)bound;
      bool chunked = false;
      uint32_t chunk = n_upper;
      if (S_total) {
        uint64_t want = emitters * S_total;
        const uint64_t cap_limit = std::max<uint64_t>(shadow_cap_req, S_total);
        if (want > cap_limit) {
          // the shadow rays of this wave may not fit: use the exact count and chunk the wave
          chunked = true;
          if (k >= 1) {
            CU(cudaEventSynchronize(count_ev[k]));
            n_upper = S.h_wave_counts[wave_base + k];
          }
          want = std::min<uint64_t>((uint64_t)n_upper * S_total, cap_limit);
        }
        if (want > S.sq_cap) {
          // the queue may still hold the previous wave's untraced shadow rays: trace them before it moves
          if (pending_shadow >= 0) {
            rc = trace_span(false, no_queue, nullptr, &wc[pending_shadow]);
            if (rc) return rc;
            pending_shadow = -1;
          }
          CU(cudaStreamSynchronize(st));
          for (int c = 0; c < 3; ++c) CU(S.d_sq[c].ensure((size_t)want * 16));
          S.sq_cap = (uint32_t)want;
        }
        if (chunked) chunk = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_upper, S.sq_cap / S_total));
      }
      if (n_upper == 0) break;
      ShadowQueue sq{S.d_sq[0].as<float4>(), S.d_sq[1].as<float4>(), S.d_sq[2].as<float4>(), S.sq_cap};
      const RayQueue qin = (k == 0) ? no_queue : ray_queue(S, cur);
      // trace: shadow rays left by the previous wave + closest hits of this wave
      rc = trace_span(k == 0, qin, &wc[k], pending_shadow >= 0 ? &wc[pending_shadow] : nullptr);
      if (rc) return rc;
      pending_shadow = -1;
      if (!chunked) {
        cudaEvent_t h0 = get_event(S, ev_used), h1 = get_event(S, ev_used);
        CU(cudaEventRecord(h0, st));
        launch_shade(S.view, S.has_shapes, fp, k == 0, qin, S.d_hits.as<float4>(), &wc[k], slot_lo, n_slots, 0, n_upper,
                     ray_queue(S, 1 - cur), sq, dc, accum, S.grid_shade, st);
        CU(cudaEventRecord(h1, st));
        shade_spans.emplace_back(h0, h1);
        ++launches;
        if (S_total) pending_shadow = (int)k;
      } else {
        // rare path (area lights x huge waves): shade a slice, trace its shadow rays, repeat
        for (uint32_t lo = 0; lo < n_upper; lo += chunk) {
          uint32_t hi = (uint32_t)std::min<uint64_t>(n_upper, (uint64_t)lo + chunk);
          cudaEvent_t h0 = get_event(S, ev_used), h1 = get_event(S, ev_used);
          CU(cudaEventRecord(h0, st));
          launch_shade(S.view, S.has_shapes, fp, k == 0, qin, S.d_hits.as<float4>(), &wc[k], slot_lo, n_slots, lo, hi,
                       ray_queue(S, 1 - cur), sq, dc, accum, S.grid_shade, st);
          CU(cudaEventRecord(h1, st));
          shade_spans.emplace_back(h0, h1);
          ++launches;
          rc = trace_span(false, no_queue, nullptr, &wc[k]);
          if (rc) return rc;
          CU(cudaMemsetAsync(&wc[k].n_shadow, 0, sizeof(uint32_t), st));
          CU(cudaMemsetAsync(&wc[k].fetch_shadow, 0, sizeof(uint32_t), st));
        }
      }
    }
    if (pending_shadow >= 0) {
      rc = trace_span(false, no_queue, nullptr, &wc[pending_shadow]);
      if (rc) return rc;
    }
  }
  if (early_used && *early_used) {
    CU(cudaStreamWaitEvent(st, S.ev_copied, 0));  // the early image must have landed before its pixels are overwritten
    launch_patch_host_image(accum, d_out, n_acc, fp.spp, early_d_out, st);
  } else if (segments_h_out && fp.packed && d_out8) {
    // RGB8 form: 48-byte segments
    const uint32_t cols_per_rank = fp.tiles_x / fp.tile_stride;
    launch_resolve_tiles_to_segments_rgb8(accum, fp, cols_per_rank, d_out8, st);
    const size_t seg = (size_t)NRB_TILE * 3;
    CU(cudaMemcpy2DAsync((char *)segments_h_out + (size_t)fp.tile_first * seg, (size_t)fp.tile_stride * seg, d_out8, seg, seg,
                         (size_t)fp.height * cols_per_rank, cudaMemcpyDeviceToHost, st));
  } else if (to_image && fp.packed && d_out8)
    launch_resolve_tiles_to_image_rgb8(accum, fp, d_out8, st);
  else if (d_out8)
    launch_resolve_rgb8(accum, n_acc, fp.spp, d_out8, st);
  else if (segments_h_out && fp.packed) {
    // this rank's tile columns as back-to-back 192-byte segments, then ONE strided 2-D DMA into the shared host image
    const uint32_t cols_per_rank = fp.tiles_x / fp.tile_stride;
    launch_resolve_tiles_to_segments(accum, fp, cols_per_rank, d_out, st);
    const size_t seg = (size_t)NRB_TILE * 3 * sizeof(float);
    CU(cudaMemcpy2DAsync((char *)segments_h_out + (size_t)fp.tile_first * seg, (size_t)fp.tile_stride * seg, d_out, seg, seg,
                         (size_t)fp.height * cols_per_rank, cudaMemcpyDeviceToHost, st));
  } else if (to_image && fp.packed)
    launch_resolve_tiles_to_image(accum, fp, d_out, st);
  else
    launch_resolve(accum, n_acc, fp.spp, d_out, st);
  ++launches;
  CU(cudaEventRecord(S.ev_end, st));
  CU(cudaMemcpyAsync(S.h_counters, dc, sizeof(Counters), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  CU(cudaStreamSynchronize(S.count_stream));
  CU(cudaGetLastError());
  for (size_t i = 0; i < wave_counts_used; ++i) waves += S.h_wave_counts[i] ? 1u : 0u;
  if (dump) {
    float total = 0.0f;
    cudaEventElapsedTime(&total, S.ev_begin, S.ev_end);
    fprintf(stderr, "[nrb] frame %.3f ms, %zu waves enqueued\n", total, wave_counts_used);
    unsigned long long dbg[4];
    debug_visit_counters(dbg, true);
    if (dbg[2])
      fprintf(stderr, "[nrb] traversals %llu: %.1f node visits, %.1f triangle tests each, longest %llu steps (since last dump)\n",
              dbg[2], (double)dbg[0] / dbg[2], (double)dbg[1] / dbg[2], dbg[3]);
    for (size_t i = 0; i < trace_spans.size(); ++i) {
      float tt = 0, th = 0, t0 = 0;
      cudaEventElapsedTime(&tt, trace_spans[i].first, trace_spans[i].second);
      if (i < shade_spans.size()) cudaEventElapsedTime(&th, shade_spans[i].first, shade_spans[i].second);
      cudaEventElapsedTime(&t0, S.ev_begin, trace_spans[i].first);
      fprintf(stderr, "[nrb]  wave %2zu n=%9u start=%.3f trace=%.3f shade=%.3f\n", i,
              i < wave_counts_used ? S.h_wave_counts[i] : 0u, t0, tt, th);
    }
  }
  if (S.h_counters->overflow) return fail(NRB_ERR_QUEUE_OVERFLOW, "a device queue overflowed (internal capacity bug)");
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->rays_primary = primary;
    stats->rays_reflect = S.h_counters->rays_reflect;
    stats->rays_refract = S.h_counters->rays_refract;
    stats->rays_shadow = S.h_counters->rays_shadow;
    stats->rays_shadow_culled = S.h_counters->rays_shadow_culled;
    stats->paths_truncated = S.h_counters->paths_truncated;
    stats->waves = waves;
    stats->kernel_launches = launches;
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, S.ev_begin, S.ev_end);
    stats->ms_device = ms;
    float tt = 0.0f;
    for (auto &sp : trace_spans) {
      float t = 0.0f;
      cudaEventElapsedTime(&t, sp.first, sp.second);
      tt += t;
    }
    auto span_ms = [](const std::vector<std::pair<cudaEvent_t, cudaEvent_t>> &v) {
      float sum = 0.0f;
      for (auto &sp : v) {
        float t = 0.0f;
        cudaEventElapsedTime(&t, sp.first, sp.second);
        sum += t;
      }
      return sum;
    };
    stats->ms_trace = tt;
    stats->ms_tail = span_ms(tail_spans);
    stats->ms_shade_kernel = span_ms(shade_spans);
    stats->ms_shade = ms - tt - stats->ms_tail;
    stats->launches_trace = (uint32_t)trace_spans.size();
    stats->launches_tail = (uint32_t)tail_spans.size();
    stats->launches_shade = (uint32_t)shade_spans.size();
    stats->rays_tail = S.h_counters->rays_tail;
    stats->bvh_nodes = S.n_bvh_nodes;
    stats->triangles = S.n_tris;
    stats->scene_bytes = S.scene_bytes;
  }
  return NRB_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// extern "C" entry points
// ---------------------------------------------------------------------------------------------
extern "C" {

const char *nrb_last_error(void) { return g_err.c_str(); }

const char *nrb_version(void) { return "nrays_b200 abi1 sm_100a"; }

int nrb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int nrb_scene_create(const NrbSceneDesc *desc, int device, NrbScene **out) {
  return nrb_scene_create_opts(desc, device, nullptr, out);
}

int nrb_scene_create_opts(const NrbSceneDesc *desc, int device, const NrbBuildOptions *opts, NrbScene **out) {
  if (!desc || !out) return fail(NRB_ERR_INVALID_ARG, "desc/out is NULL");
  uint32_t builder = opts ? opts->builder : (uint32_t)NRB_BUILDER_SAH;
  if (!opts) {  // the environment only fills in for a caller that expressed no choice
    if (const char *e = getenv("NRB_BUILDER"))
      builder = (std::string(e) == "lbvh") ? NRB_BUILDER_LBVH : (std::string(e) == "ploc") ? NRB_BUILDER_PLOC : NRB_BUILDER_SAH;
  }
  if (builder != NRB_BUILDER_SAH && builder != NRB_BUILDER_LBVH && builder != NRB_BUILDER_PLOC) return fail(NRB_ERR_INVALID_ARG, "unknown builder");
  if (desc->struct_size != sizeof(NrbSceneDesc) || desc->abi_version != NRB_ABI_VERSION)
    return fail(NRB_ERR_INVALID_ARG, "NrbSceneDesc struct_size / abi_version mismatch");
  int ndev = nrb_device_count();
  if (ndev == 0) return fail(NRB_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU path)");
  if (device < 0 || device >= ndev) return fail(NRB_ERR_INVALID_ARG, "device index out of range");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(NRB_ERR_NO_DEVICE, "device is not sm_100 class (kernels are built for sm_100a only)");
  std::unique_ptr<NrbScene> S(new NrbScene);
  S->device = device;
  S->sm_count = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&S->own_stream, cudaStreamNonBlocking));
  S->stream = S->own_stream;
  CU(cudaEventCreate(&S->ev_begin));
  CU(cudaEventCreate(&S->ev_end));
  CU(cudaHostAlloc((void **)&S->h_counters, sizeof(Counters), cudaHostAllocDefault));
  CU(S->d_counters.ensure(sizeof(Counters)));
  HostScene H;
  auto t0 = std::chrono::steady_clock::now();
  int rc = flatten_scene(*desc, H, builder);
  if (rc) return rc;
  if (getenv("NRB_CHECK_BVH")) {  // structural self-check of whatever builder ran (tests)
    std::string why;
    if (check_bvh(H, why) || check_device_nodes(H, why)) return fail(NRB_ERR_INTERNAL, "internal BVH invariant violated: " + why);
  }
  dump_bvh(H);
  PhaseTimer upt;
  rc = upload_scene(*desc, H, *S);
  if (rc) return rc;
  upt.lap("device format + upload");
  S->build_info.bvh_nodes = H.nodes.size();
  S->build_info.triangles = H.n_source_tris;
  S->build_info.shapes = H.shapes.size();
  S->build_info.planes = H.planes.size();
  S->build_info.transparent_candidates = H.candidates.size();
  S->build_info.max_depth = (uint32_t)(H.depth_tri + H.depth_mid + H.depth_top);
  S->build_info.build_ms = (float)std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  S->build_info.gpu_build_ms = H.gpu_build_ms;
  S->build_info.builder = builder;
  S->build_info.node_format = (uint32_t)S->view.node_format;
  S->grid_trace = S->sm_count * trace_blocks_per_sm(S->has_shapes);
  S->grid_tail = S->sm_count * (S->has_shapes ? 3 : kTailMinBlocks);
  S->grid_shade = S->sm_count * shade_blocks_per_sm(S->has_shapes);
  *out = S.release();
  return NRB_OK;
}

void nrb_scene_destroy(NrbScene *scene) { delete scene; }

int nrb_scene_build_info(const NrbScene *scene, NrbBuildInfo *info) {
  if (!scene || !info) return fail(NRB_ERR_INVALID_ARG, "scene/info is NULL");
  *info = scene->build_info;
  return NRB_OK;
}

int nrb_scene_validate(const NrbSceneDesc *desc, NrbBuildInfo *info) {
  if (!desc) return fail(NRB_ERR_INVALID_ARG, "desc is NULL");
  if (desc->struct_size != sizeof(NrbSceneDesc) || desc->abi_version != NRB_ABI_VERSION)
    return fail(NRB_ERR_INVALID_ARG, "NrbSceneDesc struct_size / abi_version mismatch");
  HostScene H;
  auto t0 = std::chrono::steady_clock::now();
  int rc = flatten_scene(*desc, H);
  if (rc) return rc;
  double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  std::string why;
  if (check_bvh(H, why) || check_device_nodes(H, why)) return fail(NRB_ERR_INTERNAL, "internal BVH invariant violated: " + why);
  dump_bvh(H);
  if (getenv("NRB_BUILD_TIMES")) {
    const NodeGrid g = make_node_grid(H.nodes);
    if (g.ok) fprintf(stderr, "[nrb] 16-bit grid: mean linear inflation of the child boxes %.3fx\n", grid_inflation(H.nodes, g));
  }
  if (info) {
    std::memset(info, 0, sizeof(*info));
    info->bvh_nodes = H.nodes.size();
    info->triangles = H.n_source_tris;
    info->shapes = H.shapes.size();
    info->planes = H.planes.size();
    info->transparent_candidates = H.candidates.size();
    info->max_depth = (uint32_t)(H.depth_tri + H.depth_mid + H.depth_top);
    info->build_ms = (float)ms;
  }
  return NRB_OK;
}

int nrb_scene_set_background(NrbScene *scene, const float rgb[3]) {
  if (!scene || !rgb) return fail(NRB_ERR_INVALID_ARG, "scene/rgb is NULL");
  for (int k = 0; k < 3; ++k) scene->view.background[k] = rgb[k];
  return NRB_OK;
}

int nrb_scene_set_stream(NrbScene *scene, void *cuda_stream) {
  if (!scene) return fail(NRB_ERR_INVALID_ARG, "scene is NULL");
  CU(cudaSetDevice(scene->device));
  CU(cudaStreamSynchronize(scene->stream));
  scene->stream = cuda_stream ? (cudaStream_t)cuda_stream : scene->own_stream;
  return NRB_OK;
}

int nrb_render_device(NrbScene *scene, const NrbCamera *camera, float *d_out_rgb, NrbStats *stats) {
  if (!scene || !camera || !d_out_rgb) return fail(NRB_ERR_INVALID_ARG, "scene/camera/out is NULL");
  return render_device(*scene, *camera, nullptr, d_out_rgb, nullptr, nullptr, stats);
}

int nrb_render(NrbScene *scene, const NrbCamera *camera, float *out_rgb, NrbStats *stats) {
  if (!scene || !camera || !out_rgb) return fail(NRB_ERR_INVALID_ARG, "scene/camera/out is NULL");
  CU(cudaSetDevice(scene->device));
  size_t bytes = (size_t)camera->width * camera->height * 3 * sizeof(float);
  CU(scene->d_out.ensure(std::max<size_t>(bytes, 16)));
  // With a pinned destination the copy can run beside the tail phase (render_device: early_h_out); a pageable one would
  // make cudaMemcpyAsync block the host in the middle of the frame, so it keeps the plain render -> copy order.
  float *mapped = nullptr;  // the destination as the device sees it, if it is pinned and mapped
  if (env_size("NRB_EARLY_COPY", 1) != 0) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, out_rgb) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
      void *dp = nullptr;
      if (cudaHostGetDevicePointer(&dp, out_rgb, 0) == cudaSuccess) mapped = (float *)dp;
    }
    cudaGetLastError();
  }
  if (mapped && !scene->copy_stream) {
    CU(cudaStreamCreateWithFlags(&scene->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&scene->ev_early, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&scene->ev_copied, cudaEventDisableTiming));
  }
  bool early = false;
  int rc = render_device(*scene, *camera, nullptr, scene->d_out.as<float>(), nullptr, nullptr, stats, false,
                         mapped ? out_rgb : nullptr, mapped, mapped ? &early : nullptr);
  if (rc) {
    if (early) cudaEventSynchronize(scene->ev_copied);
    return rc;
  }
  if (!early) {
    CU(cudaMemcpyAsync(out_rgb, scene->d_out.p, bytes, cudaMemcpyDeviceToHost, scene->stream));
    CU(cudaStreamSynchronize(scene->stream));
  }
  // early: the copy ran beside the tail phase and the patch kernel has stored what the tail changed (render_device
  // returns after synchronising the stream): the host image is complete
  return NRB_OK;
}

int nrb_render_rgb8(NrbScene *scene, const NrbCamera *camera, uint8_t *out_rgb8, NrbStats *stats) {
  if (!scene || !camera || !out_rgb8) return fail(NRB_ERR_INVALID_ARG, "scene/camera/out is NULL");
  CU(cudaSetDevice(scene->device));
  size_t bytes = (size_t)camera->width * camera->height * 3;
  CU(scene->d_out8.ensure(std::max<size_t>(bytes, 16)));
  int rc = render_device(*scene, *camera, nullptr, nullptr, scene->d_out8.as<uint8_t>(), nullptr, stats);
  if (rc) return rc;
  CU(cudaMemcpyAsync(out_rgb8, scene->d_out8.p, bytes, cudaMemcpyDeviceToHost, scene->stream));
  CU(cudaStreamSynchronize(scene->stream));
  return NRB_OK;
}

int nrb_render_tiles_device(NrbScene *scene, const NrbCamera *camera, const NrbTileSet *tiles, float *d_out_tiles,
                            uint32_t *n_local_tiles, NrbStats *stats) {
  if (!scene || !camera || !tiles || !d_out_tiles) return fail(NRB_ERR_INVALID_ARG, "scene/camera/tiles/out is NULL");
  return render_device(*scene, *camera, tiles, d_out_tiles, nullptr, n_local_tiles, stats);
}

int nrb_render_tiles_to_image(NrbScene *scene, const NrbCamera *camera, const NrbTileSet *tiles, float *d_image_rgb,
                              NrbStats *stats) {
  if (!scene || !camera || !tiles || !d_image_rgb) return fail(NRB_ERR_INVALID_ARG, "scene/camera/tiles/image is NULL");
  return render_device(*scene, *camera, tiles, d_image_rgb, nullptr, nullptr, stats, true);
}

int nrb_render_tiles_to_host(NrbScene *scene, const NrbCamera *camera, const NrbTileSet *tiles, float *host_image_rgb,
                             NrbStats *stats) {
  if (!scene || !camera || !tiles || !host_image_rgb) return fail(NRB_ERR_INVALID_ARG, "scene/camera/tiles/image is NULL");
  const uint32_t tiles_x = (camera->width + NRB_TILE - 1) / NRB_TILE;
  if (tiles->stride == 0 || tiles->first >= tiles->stride || camera->width % NRB_TILE != 0 || tiles_x % tiles->stride != 0)
    return fail(NRB_ERR_UNSUPPORTED, "tiles_to_host needs width % 16 == 0, tiles_x % stride == 0 and first < stride "
                                     "(each rank then owns whole tile columns); use nrb_render_tiles_to_image");
  CU(cudaSetDevice(scene->device));
  const uint32_t tiles_y = (camera->height + NRB_TILE - 1) / NRB_TILE;
  const size_t local_px = (size_t)(tiles_x / tiles->stride) * tiles_y * NRB_TILE * NRB_TILE;
  CU(scene->d_out.ensure(std::max<size_t>(local_px * 3 * sizeof(float), 16)));
  return render_device(*scene, *camera, tiles, scene->d_out.as<float>(), nullptr, nullptr, stats, false, nullptr, nullptr, nullptr,
                       host_image_rgb);
}

int nrb_render_tiles_to_image_rgb8(NrbScene *scene, const NrbCamera *camera, const NrbTileSet *tiles, uint8_t *d_image_rgb8,
                                   NrbStats *stats) {
  if (!scene || !camera || !tiles || !d_image_rgb8) return fail(NRB_ERR_INVALID_ARG, "scene/camera/tiles/image is NULL");
  return render_device(*scene, *camera, tiles, nullptr, d_image_rgb8, nullptr, stats, true);
}

int nrb_render_tiles_to_host_rgb8(NrbScene *scene, const NrbCamera *camera, const NrbTileSet *tiles, uint8_t *host_image_rgb8,
                                  NrbStats *stats) {
  if (!scene || !camera || !tiles || !host_image_rgb8) return fail(NRB_ERR_INVALID_ARG, "scene/camera/tiles/image is NULL");
  const uint32_t tiles_x = (camera->width + NRB_TILE - 1) / NRB_TILE;
  if (tiles->stride == 0 || tiles->first >= tiles->stride || camera->width % NRB_TILE != 0 || tiles_x % tiles->stride != 0)
    return fail(NRB_ERR_UNSUPPORTED, "tiles_to_host_rgb8 needs width % 16 == 0, tiles_x % stride == 0 and first < stride "
                                     "(each rank then owns whole tile columns); use nrb_render_tiles_to_image_rgb8");
  CU(cudaSetDevice(scene->device));
  const uint32_t tiles_y = (camera->height + NRB_TILE - 1) / NRB_TILE;
  const size_t local_px = (size_t)(tiles_x / tiles->stride) * tiles_y * NRB_TILE * NRB_TILE;
  CU(scene->d_out8.ensure(std::max<size_t>(local_px * 3, 16)));
  return render_device(*scene, *camera, tiles, nullptr, scene->d_out8.as<uint8_t>(), nullptr, stats, false, nullptr, nullptr, nullptr,
                       host_image_rgb8);
}

int nrb_ipc_alloc(int device, uint64_t bytes, void **d_ptr, NrbIpcHandle *handle) {
  if (!d_ptr || !handle || bytes == 0) return fail(NRB_ERR_INVALID_ARG, "ipc_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(NrbIpcHandle), "NrbIpcHandle must hold a cudaIpcMemHandle_t");
  CU(cudaSetDevice(device));
  void *p = nullptr;
  CU(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(NRB_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  CU(cudaMemset(p, 0, bytes));
  std::memcpy(handle->bytes, &h, sizeof(h));
  *d_ptr = p;
  return NRB_OK;
}

int nrb_ipc_open(int device, const NrbIpcHandle *handle, void **d_ptr) {
  if (!d_ptr || !handle) return fail(NRB_ERR_INVALID_ARG, "ipc_open: bad arguments");
  CU(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle->bytes, sizeof(h));
  void *p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(NRB_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
  }
  *d_ptr = p;
  return NRB_OK;
}

int nrb_ipc_close(int device, void *d_ptr) {
  if (!d_ptr) return NRB_OK;
  CU(cudaSetDevice(device));
  CU(cudaIpcCloseMemHandle(d_ptr));
  return NRB_OK;
}

int nrb_ipc_free(int device, void *d_ptr) {
  if (!d_ptr) return NRB_OK;
  CU(cudaSetDevice(device));
  CU(cudaFree(d_ptr));
  return NRB_OK;
}

int nrb_host_register(int device, void *host_ptr, uint64_t bytes, void **d_ptr) {
  if (!host_ptr || !d_ptr || bytes == 0) return fail(NRB_ERR_INVALID_ARG, "host_register: bad arguments");
  CU(cudaSetDevice(device));
  cudaError_t e = cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(NRB_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
  }
  void *dp = nullptr;
  e = cudaHostGetDevicePointer(&dp, host_ptr, 0);
  if (e != cudaSuccess) {
    cudaGetLastError();
    cudaHostUnregister(host_ptr);
    return fail(NRB_ERR_CUDA, std::string("cudaHostGetDevicePointer: ") + cudaGetErrorString(e));
  }
  *d_ptr = dp;
  return NRB_OK;
}

int nrb_host_unregister(int device, void *host_ptr) {
  if (!host_ptr) return NRB_OK;
  CU(cudaSetDevice(device));
  CU(cudaHostUnregister(host_ptr));
  return NRB_OK;
}

uint32_t nrb_tile_count(uint32_t width, uint32_t height) {
  return ((width + NRB_TILE - 1) / NRB_TILE) * ((height + NRB_TILE - 1) / NRB_TILE);
}

uint32_t nrb_tile_count_local(uint32_t width, uint32_t height, const NrbTileSet *tiles) {
  uint32_t n = nrb_tile_count(width, height);
  if (!tiles || tiles->stride == 0 || tiles->first >= n) return tiles ? 0 : n;
  return (n - tiles->first + tiles->stride - 1) / tiles->stride;
}

int nrb_untile_device(int device, void *cuda_stream, const float *d_gathered, uint32_t n_ranks,
                      uint32_t tiles_per_rank, uint32_t width, uint32_t height, float *d_out_rgb) {
  if (!d_gathered || !d_out_rgb || n_ranks == 0) return fail(NRB_ERR_INVALID_ARG, "untile: bad arguments");
  CU(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  launch_untile(d_gathered, n_ranks, tiles_per_rank, width, height, d_out_rgb, st);
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  return NRB_OK;
}

void *nrb_host_alloc(uint64_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    g_err = "cudaHostAlloc failed";
    return nullptr;
  }
  return p;
}

void nrb_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
