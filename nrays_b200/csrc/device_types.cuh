// device_types.cuh — HBM data layout of the flattened scene and of the wavefront queues.
//
// Everything the kernels touch is f32 / u32 in 16-byte units so every access is a single 128-bit
// load (LDG.E.128).  See DESIGN.md "Data layout in HBM" for the byte accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nrb {

// ---- BVH ------------------------------------------------------------------------------------
// One 64-byte node holds BOTH children's boxes (Aila-Laine layout), so one traversal step fetches one
// cache-line-aligned record (two 256-bit loads) and tests two boxes.
// The builders (bvh_build.cpp, lbvh.cu) and the host-side checks use the lo/hi form:
//   n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)
//   n1 = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//   n2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)
//   n3 = (child0, child1, -, -) as int bits
// The DEVICE copy is converted at upload (api.cu: to_device_nodes) to centres + half extents, which turns
// the slab test into 9 FFMA + 4 min/max per box (kernels.cu: test_children).  Node formats (SceneView.node_format, chosen per scene):
//   0 (64 B): n0 = (c0.centre.xyz, c1.centre.x)   n1 = (c1.centre.yz, c0.half.xy)   n2 = (c0.half.z, c1.half.xyz)   n3 as above
//   2 (48 B used, 64 B stride): (c0.cx, c0.cy, c1.cx, c1.cy) (c0.cz, c1.cz, bf16x2(c0.hx, c0.hy), bf16x2(c1.hx, c1.hy))
//             (bf16x2(c0.hz, c1.hz), child0, child1, -) (unused); half extents rounded UP to bf16
//   3 (32 B, 32-byte stride; mesh-only scenes): six words (hi << 16 | lo), one per child and axis — c0.x c0.y c0.z c1.x c1.y c1.z —
//             then child0, child1.  lo / hi are cells of ONE 16-bit grid over the scene (SceneView.grid_lo / grid_cell), snapped
//             outward plus one cell, so the record is fetched with one 256-bit load and half the bytes reach the registers.
//   4         format 3 records walked by the speculative loop (scenes beyond L2)
// Child code c:  c >= 0 -> inner node index;  c < 0 -> leaf, ~c = (first << 3) | ((count-1) << 1) | is_shape
//   is_shape = 0: triangles [first, first+count) of the leaf-ordered triangle array (count <= 4)
//   is_shape = 1: analytic shape `first` of the shape table (count == 1)
struct __align__(16) BvhNode {
  float4 n0, n1, n2;
  int4 n3;
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 bytes");

constexpr int kEmpty = 0x7FFFFFFF;  // stack sentinel / "no root"
constexpr int kMaxLeafTris = 4;
constexpr int kStackSize = 64;

__host__ __device__ inline int make_leaf(uint32_t first, uint32_t count, bool is_shape) {
  return ~(int)((first << 3) | ((count - 1u) << 1) | (is_shape ? 1u : 0u));
}

// Triangle: 48 bytes, world space, edge form for the two-sided Ericson test.
//   t0 = (v0.xyz, scene-node id as int bits), t1 = (e1.xyz, 0), t2 = (e2.xyz, 0);  e1 = v1-v0, e2 = v2-v0
struct __align__(16) Tri {
  float4 t0, t1, t2;
};
// Per-triangle vertex uvs, fetched only by the shade / shadow-filter code: 24 bytes.
struct __align__(8) TriUV {  // three 64-bit loads
  float u0, v0, u1, v1, u2, v2;
};

// Analytic SceneNode geometry (ball, cuboid, cylinder, capsule, cone, plane): 80 bytes.
struct __align__(16) Shape {
  int kind;   // NRB_SHAPE_*
  int node;   // SceneNode index
  int solid;  // SceneNode.solid
  int _pad;
  float p[4];    // shape parameters (see include/nrays_b200.h)
  float rot[9];  // row-major rotation of the isometry
  float trans[3];
};
static_assert(sizeof(Shape) == 80, "Shape must be 80 bytes");

// SceneNode fields the shade code needs (src/scene_node.rs:8-19): 32 bytes.
struct __align__(16) NodeInfo {
  int material;
  float refl_mix;
  float refl_att;
  float alpha;
  float refr_coeff;
  int flags;  // bit0: shadow-transparent candidate (own BVH root)
  int nmap_tex;  // SceneNode.nmap (src/scene_node.rs:60-70): texture whose mean rgb is subtracted from the node's toi, or -1
  int _pad;
};

struct __align__(16) Material {  // 64 bytes
  int kind;
  float ambient[3];
  float diffuse[3];
  float specular[3];
  float shininess;
  int tex;
  int alpha_tex;
  int _pad[3];
};
static_assert(sizeof(Material) == 64, "Material layout");

struct Texture {
  uint32_t w, h;
  int interp, overflow;
  uint32_t offset;  // texel offset into the RGBA32F pool
  uint32_t avail;   // texels from offset to the end of the pool
};

struct Light {
  float pos[3];
  float radius;
  float color[3];
  uint32_t racsample;
};

// A shadow-transparent candidate SceneNode: closest hit per node decides filter vs occlusion
// (src/scene.rs:304-339).
struct Candidate {
  float lo[3], hi[3];
  int root;  // child code of its own sub-tree (triangle mesh) or shape leaf
  int node;  // SceneNode index
};

struct SceneView {
  const void *nodes;    // device node records: 64-byte stride (formats 0, 2) or 32-byte stride (formats 3, 4)
  int node_format;      // 0: fp32 centres + half extents (64 B read per visit); 2: bf16 half extents (48 B read per visit);
                        // 3: child boxes on the scene's 16-bit grid (32 B read per visit); 4: format 3 + speculative loop
  float grid_lo[3];     // formats 3 / 4: plane coordinate = grid_lo + q * grid_cell, q in [0, 65535]
  float grid_cell[3];
  const Tri *tris;
  const TriUV *tri_uvs;
  const Shape *shapes;
  const NodeInfo *node_info;
  const Material *materials;
  const Texture *textures;
  const float4 *texels;
  const Light *lights;
  const int *planes;            // shape indices of all planes (always tested, never in the BVH)
  const Candidate *candidates;  // shadow-transparent candidates that are not planes
  const Candidate *nmaps;       // nodes with a depth-shift texture: own sub-root, box = the REFERENCE's node AABB (pruning order)
  int root_all;                 // closest-hit entry (kEmpty if the scene has only planes)
  int root_opaque;              // any-hit entry for shadow rays (kEmpty if none)
  int n_planes;
  int n_candidates;
  int n_nmap;
  int n_lights;
  int shadow_samples;  // sum over lights of racsample^2
  float background[3];
};

// ---- wavefront queues (SoA of 16-byte columns) -------------------------------------------------
// Ray record, 48 bytes in three columns:
//   a = (o.x, o.y, o.z, d.x)   b = (d.y, d.z, weight, energy)   c = (refr, gid, path, depth) [gid/path/depth as uint bits]
// gid = pixel_index * spp + sample (global, independent of sharding) -> RNG counter + accumulation address.
struct RayQueue {
  float4 *a, *b, *c;
  uint32_t capacity;
};
// Hit record, 16 bytes: (t, prim, u, v).  prim (uint bits): triangle index in leaf order, or
// 0x80000000 | shape index, or kMiss.
constexpr uint32_t kMiss = 0xFFFFFFFFu;
constexpr uint32_t kSkip = 0xFFFFFFFEu;  // primary slot outside the image (ragged tile): no ray
constexpr uint32_t kShapeBit = 0x80000000u;
// Shadow ray record, 48 bytes in three columns:
//   a = (o.xyz, tmax)   b = (d.xyz, pixel address as uint bits)   c = (contribution rgb, -)
struct ShadowQueue {
  float4 *a, *b, *c;
  uint32_t capacity;
};

// Per-wave queue counters: wave k reads its ray count from wc[k].n_rays (written by the shade of wave
// k-1), its shade writes wc[k].n_shadow and wc[k+1].n_rays.  One zeroed array per batch: no counter
// resets between kernels.
struct WaveCounters {
  uint32_t n_rays;         // rays in the queue this wave consumes
  uint32_t n_shadow;       // shadow rays emitted by this wave's shade
  uint32_t fetch_closest;  // persistent-kernel work cursors
  uint32_t fetch_shadow;
};

// Frame statistics block (one per scene handle)
struct Counters {
  uint32_t overflow;  // set if any queue append was dropped
  uint32_t _pad[3];
  unsigned long long rays_reflect, rays_refract, rays_shadow, paths_truncated;
  unsigned long long rays_shadow_culled;  // light samples not cast because their contribution is exactly zero
  unsigned long long rays_tail;           // closest-hit queries answered inside the tail kernel
  unsigned long long dbg_nodes, dbg_tris, dbg_rays;  // -DNRB_COUNT_VISITS builds only
};

// Exact division of a 32-bit number by a run-time constant: q = hi64(M * n), M = floor(2^64 / d) + 1
// (Lemire et al.; exact for every 32-bit n and d >= 2).  Three instructions instead of a ~25-instruction
// software division; the kernels divide by spp / tiles_x / width several times per ray.
struct FastDiv {
  uint32_t mul_lo, mul_hi;
  uint32_t d;
  uint32_t _pad;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  unsigned long long m = d >= 2 ? (~0ull / d) + 1ull : 0ull;  // floor((2^64 - 1) / d) + 1 == floor(2^64 / d) + 1 unless d | 2^64
  if (d >= 2 && (d & (d - 1)) == 0) m = (1ull << 63) / (d >> 1);  // power of two: exactly 2^64 / d
  f.mul_lo = (uint32_t)m, f.mul_hi = (uint32_t)(m >> 32);
  f._pad = 0;
  return f;
}

// Per-frame constants
struct FrameParams {
  uint32_t width, height, spp, max_depth;
  uint32_t tiles_x, tiles_y;
  uint32_t tile_first, tile_stride;  // tile set rendered by this call
  uint32_t n_local_tiles;
  uint32_t packed;  // 0: accumulate into a row-major image; 1: into packed tiles [n_local][16][16]
  float window;
  float inv_w, inv_h;
  float eye[3];
  // primary direction = normalize(dx * ndc.x + dy * ndc.y + d0) (host-derived in f64 from the
  // inverse view-projection so the device never subtracts eye from a near-plane point)
  float dx[3], dy[3], d0[3];
  float wx, wy, w0;  // homogeneous w = wx * ndc.x + wy * ndc.y + w0; the direction flips when w < 0
  uint32_t seed_lo, seed_hi;
  FastDiv div_spp, div_tiles_x, div_width, div_tile_stride;
};

}  // namespace nrb
