// kernels.cu — the sm_100a kernels of the nrays render path (wavefront formulation).
//
//   K1 primary_ray()          src/scene.rs:67-89      jitter + unproject, fused into wave 0 of K2 and K4
//   K2 trace_kernel/closest   src/scene.rs:163-166, 262-283 + ncollide3d BVT/RayCast
//   K3 trace_kernel/shadow    src/scene.rs:147-161, 285-339 (Scene::intersects_ray + transparent filter)
//   K4 shade_kernel           src/scene.rs:168-252, src/phong_material.rs:72-151, src/light.rs:56-63,
//                             src/texture2d.rs:207-256, normal/uv materials
//   K5 resolve_kernel         src/scene.rs:94 (tot_c / spp) [+ src/image.rs:64-77 RGB8 quantisation]
//   K5+K6 resolve_tiles_to_image_kernel   (multi-GPU) this rank's tiles / spp stored straight into the owner's row-major
//                             image, local or peer memory over NVLink: the path's one exchange, fused into the resolve
//   K6 untile_kernel          (multi-GPU, all-gather fallback) packed 16x16 tiles -> row-major image
//
// The reference recursion `trace` is linear in its reflection / refraction children, so each ray
// carries a scalar weight and every hit / miss / unoccluded light sample adds its weighted colour
// straight into the pixel accumulator (SURVEY §7 "Hard parts").
#include <cooperative_groups.h>

#include "device_math.cuh"
#include "kernels.h"

namespace nrb {

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
NRB_DI uint32_t lane_id() { return threadIdx.x & 31u; }

// 8-bit Morton decode (4 bits x, 4 bits y) for the in-tile pixel order
NRB_DI uint32_t compact4(uint32_t v) {
  v &= 0x55u;
  v = (v | (v >> 1)) & 0x33u;
  v = (v | (v >> 2)) & 0x0Fu;
  return v;
}

NRB_DI uint32_t accum_index(const FrameParams &fp, uint32_t ipt) {
  if (!fp.packed) return ipt;
  uint32_t y = fdiv(ipt, fp.div_width), x = ipt - y * fp.width;
  uint32_t tile = (y / NRB_TILE) * fp.tiles_x + (x / NRB_TILE);
  uint32_t lt = fdiv(tile - fp.tile_first, fp.div_tile_stride);
  return lt * (NRB_TILE * NRB_TILE) + (y % NRB_TILE) * NRB_TILE + (x % NRB_TILE);
}

NRB_DI void accum_add(float4 *accum, uint32_t idx, V3 c) {
  if (c.x == 0.0f && c.y == 0.0f && c.z == 0.0f) return;
  atomicAdd(&accum[idx], make_float4(c.x, c.y, c.z, 0.0f));  // one 128-bit RED (sm_90+)
}

// ---------------------------------------------------------------------------------------------
// K1 — primary rays (src/scene.rs:67-89), generated on the fly.
// Wave 0 never materialises its rays: the closest-hit kernel and the shade kernel both evaluate this
// function for sample slot `slot` (identical arithmetic -> identical ray), so 48 B/ray of queue
// writes and 80 B/ray of reads never touch HBM.  Slot order: tile-major, Morton inside the 16x16
// tile, samples innermost, so a warp covers a compact pixel block.
// ---------------------------------------------------------------------------------------------
NRB_DI bool primary_ray(const FrameParams &fp, uint32_t slot, V3 &o, V3 &d, uint32_t &gid) {
  // slot = (local tile * 256 + pixel in tile) * spp + sample
  uint32_t q = fdiv(slot, fp.div_spp), s = slot - q * fp.spp;
  uint32_t lt = q / (NRB_TILE * NRB_TILE), p = q % (NRB_TILE * NRB_TILE);
  uint32_t tile = fp.tile_first + lt * fp.tile_stride;
  uint32_t ty = fdiv(tile, fp.div_tiles_x), tx = tile - ty * fp.tiles_x;
  uint32_t x = tx * NRB_TILE + compact4(p), y = ty * NRB_TILE + compact4(p >> 1);
  if (x >= fp.width || y >= fp.height) return false;
  uint32_t ipt = y * fp.width + x;
  float jx = 0.0f, jy = 0.0f;
  if (fp.window != 0.0f) {
    uint32_t rnd[4];
    philox4x32_10(ipt, s, 0u, 0u, fp.seed_lo, fp.seed_hi ^ kStreamPrimary, rnd);
    jx = (u24(rnd[0]) - 0.5f) * fp.window;
    jy = (u24(rnd[1]) - 0.5f) * fp.window;
  }
  float fx = (float)x + jx, fy = (float)y + jy;
  float ndx = (fx * fp.inv_w - 0.5f) * 2.0f;
  float ndy = -(fy * fp.inv_h - 0.5f) * 2.0f;
  V3 dir = mk(fp.dx[0], fp.dx[1], fp.dx[2]) * ndx + mk(fp.dy[0], fp.dy[1], fp.dy[2]) * ndy + mk(fp.d0[0], fp.d0[1], fp.d0[2]);
  float w = fp.wx * ndx + fp.wy * ndy + fp.w0;
  dir = normalize(dir);
  if (w < 0.0f) dir = -dir;
  o = mk(fp.eye[0], fp.eye[1], fp.eye[2]);
  d = dir;
  gid = ipt * fp.spp + s;
  return true;
}

// ---------------------------------------------------------------------------------------------
// BVH traversal (per-lane while-while with an explicit stack)
// ---------------------------------------------------------------------------------------------
#ifdef NRB_COUNT_VISITS
__device__ unsigned long long g_dbg[4];  // nodes, tris, rays, -
#endif

struct Hit {
  float t;
  uint32_t prim;
  float u, v;
};

// One node record as the traversal loop sees it (device_types.cuh "node formats").  Three layouts are compiled and the scene
// picks one at creation (api.cu: upload_scene):
//   format 0: 64 bytes, two 256-bit loads — the fastest form while the scene lives in L2 and neighbouring rays visit the same
//             nodes (C3: 2.24 ms; format 2 under the same loop 2.3, format 3 2.57).
//   format 2: 48 bytes used of a 64-byte record: the six half extents are stored as bf16 (rounded up: the box only grows, by
//             < 0.8 % of its half extent), fetched with one 256-bit + one 128-bit load.  The record keeps its 64-byte stride so
//             a divergent lane still touches ONE cache line per visit.  Scenes beyond L2 that have analytic shapes.
//   format 3: 32 bytes, ONE 256-bit load: the twelve planes are 16-bit cells of one grid over the whole scene (api.cu:
//             to_device_nodes).  A cell becomes a float with one PRMT (0x4B00 in front of it: 2^23 + q) whose selector picks the
//             near or the far plane for this ray's sign, so the slab test is 12 PRMT + 4 FFMA2 + 4 FFMA + 8 min/max and needs
//             no centre / half-extent form; the 2^23 is folded into the ray's constant, which costs half a cell of rounding —
//             the cells are snapped outward by a whole one.  Mesh-only scenes (analytic shapes may have unbounded boxes).
//   format 4: format-3 records under the speculative loop — the choice for mesh-only scenes beyond L2 (C4: 7.14 ms against 7.32
//             for format 2, with half the node memory).
// What the formats taught (profiles/README.md): the L1 data stage hands 128 bytes per clock to the register file whether or not
// the lanes agree on the address (64 bytes x 32 lanes = 16 clocks per warp and visit; ncu: 128 M of the 163 M data-stage
// wavefronts of the primary launch of C3 are node records), so it reads 85 % busy — yet halving those bytes buys nothing on C3:
// the loop is bound by instruction issue, and the PRMTs of format 3 land on the ALU pipe (min/max, compares, selects, integer
// adds), which issues at half rate and is the busiest pipe of the visit.  Fewer bytes pay only where lines miss L1 / L2 (C4).
// Formats 0 and 2 run the slab test on the packed FP32 pipe form (FFMA2, PTX fma.rn.f32x2) where it is free: measured on B200 the
// FFMA2 issues at half the FFMA rate (1.86 vs 3.88 warp-instructions / clk / SM, scripts/ubench_ffma2.cu), so it saves issue
// slots but no FMA-pipe time, and on this loop it is neutral (format 1 = format 0 with FFMA2: 2.37 vs 2.34 ms on C3).
// Format 0 therefore keeps the scalar FFMA form of round 1; format 2 uses FFMA2 (its operands arrive as pairs anyway).
typedef unsigned long long u64;
NRB_DI u64 pk2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
NRB_DI void upk2(u64 r, float &lo, float &hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r)); }
NRB_DI u64 fma2(u64 a, u64 b, u64 c) {  // FFMA2: two fp32 FMAs in one issue slot (sm_100+)
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// Triangle fetch of the leaf test (48-byte record: v0, e1, e2).  Loading with L1::no_allocate — triangles are used once per ray —
// was measured and dropped (C3 2.26-2.28 vs 2.22 ms, C4 unchanged).
NRB_DI float4 ld_tri(const float4 *p) { return __ldg(p); }

template <int FMT>
struct NodeRec;
template <>
struct NodeRec<0> {
  float4 n0, n1, n2;
  int c0, c1;
};
template <>
struct NodeRec<2> {
  u64 cxy0, cxy1, cz01;  // (c0.x,c0.y) (c1.x,c1.y) (c0.z,c1.z)
  u64 hxy0, hxy1, hz01;  // half extents, same pairing
  int c0, c1;
};

template <>
struct NodeRec<3> {
  uint32_t w[6];  // (hi << 16 | lo) cells: c0.x c0.y c0.z c1.x c1.y c1.z
  int c0, c1;
};
// FMT -> record layout / loop form
constexpr int node_layout(int fmt) { return fmt == 4 ? 3 : fmt; }
constexpr bool speculative_loop(int fmt) { return fmt == 2 || fmt == 4; }

template <int LAYOUT>
NRB_DI NodeRec<LAYOUT> load_node(const SceneView &sc, int node);

template <>
NRB_DI NodeRec<0> load_node<0>(const SceneView &sc, int node) {
  NodeRec<0> r;
  const char *np = reinterpret_cast<const char *>(sc.nodes) + (size_t)(unsigned)node * 64u;
  // two 256-bit loads (LDG.E.256, sm_100+) fetch the whole record; the child codes arrive with the boxes
  int pad0, pad1;  // the record's two unused words
  (void)pad0, (void)pad1;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.n0.x), "=f"(r.n0.y), "=f"(r.n0.z), "=f"(r.n0.w), "=f"(r.n1.x), "=f"(r.n1.y), "=f"(r.n1.z), "=f"(r.n1.w)
               : "l"(np));
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.n2.x), "=f"(r.n2.y), "=f"(r.n2.z), "=f"(r.n2.w), "=r"(r.c0), "=r"(r.c1), "=r"(pad0), "=r"(pad1)
               : "l"(np + 32));
  return r;
}

template <>
NRB_DI NodeRec<2> load_node<2>(const SceneView &sc, int node) {
  NodeRec<2> r;
  const char *na = reinterpret_cast<const char *>(sc.nodes) + (size_t)(unsigned)node * 64u;
  uint32_t wxy0, wxy1, wz, pad;
  (void)pad;
  u64 hw;
  asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r.cxy0), "=l"(r.cxy1), "=l"(r.cz01), "=l"(hw) : "l"(na));
  asm volatile("ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(wz), "=r"(r.c0), "=r"(r.c1), "=r"(pad) : "l"(na + 32));
  wxy0 = (uint32_t)hw, wxy1 = (uint32_t)(hw >> 32);
  // bf16 pairs -> fp32 pairs.  Low half: << 16, written as a multiply so it issues on the FMA pipe (IMAD) — the ALU pipe
  // (min / max, compares, selects: half rate) is the busier one in this loop.  High half: the word as it is; the low half's
  // bits land in the mantissa tail and make the half extent up to 2^-7 larger — the box only grows.
  r.hxy0 = pk2(__uint_as_float(wxy0 * 65536u), __uint_as_float(wxy0));
  r.hxy1 = pk2(__uint_as_float(wxy1 * 65536u), __uint_as_float(wxy1));
  r.hz01 = pk2(__uint_as_float(wz * 65536u), __uint_as_float(wz));
  return r;
}

template <>
NRB_DI NodeRec<3> load_node<3>(const SceneView &sc, int node) {
  NodeRec<3> r;
  const char *na = reinterpret_cast<const char *>(sc.nodes) + (size_t)(unsigned)node * 32u;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.c0), "=r"(r.c1)
               : "l"(na));
  return r;
}

// Per-ray constants of the slab test: t(plane) = plane * (1/d) - o * (1/d).
// Grid form (node layout 3): t(cell q) = (2^23 + q) * idx + oodx with idx = cell / d and oodx = (grid_lo - o) / d - 2^23 * idx.
struct RayPre {
  float idx, idy, idz, oodx, oody, oodz;
};
NRB_DI RayPre ray_pre(V3 o, V3 d) {
  const float ooeps = 1.0e-24f;
  RayPre p;
  p.idx = 1.0f / (fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x));
  p.idy = 1.0f / (fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y));
  p.idz = 1.0f / (fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z));
  p.oodx = o.x * p.idx, p.oody = o.y * p.idy, p.oodz = o.z * p.idz;
  return p;
}

template <int LAYOUT>
NRB_DI RayPre ray_pre_for(const SceneView &sc, V3 o, V3 d) {
  if (LAYOUT != 3) return ray_pre(o, d);
  const float ooeps = 1.0e-24f;
  const float ix = 1.0f / (fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x));
  const float iy = 1.0f / (fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y));
  const float iz = 1.0f / (fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z));
  RayPre p;
  p.idx = sc.grid_cell[0] * ix, p.idy = sc.grid_cell[1] * iy, p.idz = sc.grid_cell[2] * iz;
  p.oodx = fmaf(-8388608.0f, p.idx, (sc.grid_lo[0] - o.x) * ix);
  p.oody = fmaf(-8388608.0f, p.idy, (sc.grid_lo[1] - o.y) * iy);
  p.oodz = fmaf(-8388608.0f, p.idz, (sc.grid_lo[2] - o.z) * iz);
  return p;
}

// What a node visit needs besides RayPre: nothing for the float layouts; for the grid layout the PRMT selectors that turn
// one half of a (hi << 16 | lo) word into 2^23 + q — the near plane is `lo` where the ray runs in +axis direction.
template <int LAYOUT>
struct RayAux {};
template <>
struct RayAux<3> {
  uint32_t nx, ny, nz, fx, fy, fz;  // near / far plane selectors
};
template <int LAYOUT>
NRB_DI RayAux<LAYOUT> ray_aux(const RayPre &) {
  return RayAux<LAYOUT>();
}
template <>
NRB_DI RayAux<3> ray_aux<3>(const RayPre &p) {
  RayAux<3> a;  // __byte_perm(word, 0x4B000000, s): 0x7410 -> 0x4B00'lo, 0x7432 -> 0x4B00'hi
  a.nx = p.idx < 0.0f ? 0x7432u : 0x7410u;
  a.ny = p.idy < 0.0f ? 0x7432u : 0x7410u;
  a.nz = p.idz < 0.0f ? 0x7432u : 0x7410u;
  a.fx = a.nx ^ 0x22u, a.fy = a.ny ^ 0x22u, a.fz = a.nz ^ 0x22u;
  // opaque to the optimiser: otherwise it re-derives the six selectors from the signs at EVERY visit (3 FSETP + 3 SEL + 3 LOP3)
  asm volatile("" : "+r"(a.nx), "+r"(a.ny), "+r"(a.nz), "+r"(a.fx), "+r"(a.fy), "+r"(a.fz));
  return a;
}

// Both children's slab tests in centre / half-extent form: t(centre) -+ half * |1/d| — 9 FFMA + 4
// min/max per box and no lo/hi sort (the FMNMX pipe, not the FMA pipe, limits the classic form).
NRB_DI void test_children(const NodeRec<0> &n, const RayPre &p, const RayAux<0> &, float tbest, float &c0min, float &c0max,
                          float &c1min, float &c1max) {
  const float aidx = fabsf(p.idx), aidy = fabsf(p.idy), aidz = fabsf(p.idz);
  float c0tx = fmaf(n.n0.x, p.idx, -p.oodx), c0ty = fmaf(n.n0.y, p.idy, -p.oody), c0tz = fmaf(n.n0.z, p.idz, -p.oodz);
  float c1tx = fmaf(n.n0.w, p.idx, -p.oodx), c1ty = fmaf(n.n1.x, p.idy, -p.oody), c1tz = fmaf(n.n1.y, p.idz, -p.oodz);
  c0min = fmaxf(fmaxf(fmaf(-n.n1.z, aidx, c0tx), fmaf(-n.n1.w, aidy, c0ty)), fmaxf(fmaf(-n.n2.x, aidz, c0tz), 0.0f));
  c0max = fminf(fminf(fmaf(n.n1.z, aidx, c0tx), fmaf(n.n1.w, aidy, c0ty)), fminf(fmaf(n.n2.x, aidz, c0tz), tbest));
  c1min = fmaxf(fmaxf(fmaf(-n.n2.y, aidx, c1tx), fmaf(-n.n2.z, aidy, c1ty)), fmaxf(fmaf(-n.n2.w, aidz, c1tz), 0.0f));
  c1max = fminf(fminf(fmaf(n.n2.y, aidx, c1tx), fmaf(n.n2.z, aidy, c1ty)), fminf(fmaf(n.n2.w, aidz, c1tz), tbest));
}
// The same test on pairs: 9 FFMA2 + 8 min/max.  ptxas folds the |1/d| and -|1/d| operands into FFMA2's abs / neg modifiers
// and the (1/d.z, 1/d.z) pair into its scalar-broadcast operand form, so the ray constants stay in six registers.
NRB_DI void test_children(const NodeRec<2> &n, const RayPre &p, const RayAux<2> &, float tbest, float &c0min, float &c0max,
                          float &c1min, float &c1max) {
  const u64 idxy = pk2(p.idx, p.idy), idzz = pk2(p.idz, p.idz);
  const u64 noodxy = pk2(-p.oodx, -p.oody), noodzz = pk2(-p.oodz, -p.oodz);
  const u64 aidxy = pk2(fabsf(p.idx), fabsf(p.idy)), aidzz = pk2(fabsf(p.idz), fabsf(p.idz));
  const u64 naidxy = pk2(-fabsf(p.idx), -fabsf(p.idy)), naidzz = pk2(-fabsf(p.idz), -fabsf(p.idz));
  const u64 t0 = fma2(n.cxy0, idxy, noodxy), t1 = fma2(n.cxy1, idxy, noodxy), tz = fma2(n.cz01, idzz, noodzz);
  float l0x, l0y, l1x, l1y, l0z, l1z, h0x, h0y, h1x, h1y, h0z, h1z;
  upk2(fma2(n.hxy0, naidxy, t0), l0x, l0y);
  upk2(fma2(n.hxy0, aidxy, t0), h0x, h0y);
  upk2(fma2(n.hxy1, naidxy, t1), l1x, l1y);
  upk2(fma2(n.hxy1, aidxy, t1), h1x, h1y);
  upk2(fma2(n.hz01, naidzz, tz), l0z, l1z);
  upk2(fma2(n.hz01, aidzz, tz), h0z, h1z);
  c0min = fmaxf(fmaxf(l0x, l0y), fmaxf(l0z, 0.0f));
  c0max = fminf(fminf(h0x, h0y), fminf(h0z, tbest));
  c1min = fmaxf(fmaxf(l1x, l1y), fmaxf(l1z, 0.0f));
  c1max = fminf(fminf(h1x, h1y), fminf(h1z, tbest));
}

// Grid layout: t = (2^23 + q) * s + k per plane; the planes arrive sorted (near, far), so no lo/hi min/max.  x and y run as
// pairs on FFMA2 with the ray's (s.x, s.y) / (k.x, k.y) register pairs, z on scalar FFMA — no duplicated constants.
NRB_DI void test_children(const NodeRec<3> &n, const RayPre &p, const RayAux<3> &a, float tbest, float &c0min, float &c0max,
                          float &c1min, float &c1max) {
  const uint32_t magic = 0x4B000000u;
  const u64 sxy = pk2(p.idx, p.idy), kxy = pk2(p.oodx, p.oody);
#define NRB_Q(w, sel) __uint_as_float(__byte_perm(w, magic, sel))
  float n0x, n0y, f0x, f0y, n1x, n1y, f1x, f1y;
  upk2(fma2(pk2(NRB_Q(n.w[0], a.nx), NRB_Q(n.w[1], a.ny)), sxy, kxy), n0x, n0y);
  upk2(fma2(pk2(NRB_Q(n.w[0], a.fx), NRB_Q(n.w[1], a.fy)), sxy, kxy), f0x, f0y);
  upk2(fma2(pk2(NRB_Q(n.w[3], a.nx), NRB_Q(n.w[4], a.ny)), sxy, kxy), n1x, n1y);
  upk2(fma2(pk2(NRB_Q(n.w[3], a.fx), NRB_Q(n.w[4], a.fy)), sxy, kxy), f1x, f1y);
  const float n0z = fmaf(NRB_Q(n.w[2], a.nz), p.idz, p.oodz), f0z = fmaf(NRB_Q(n.w[2], a.fz), p.idz, p.oodz);
  const float n1z = fmaf(NRB_Q(n.w[5], a.nz), p.idz, p.oodz), f1z = fmaf(NRB_Q(n.w[5], a.fz), p.idz, p.oodz);
#undef NRB_Q
  c0min = fmaxf(fmaxf(n0x, n0y), fmaxf(n0z, 0.0f));
  c0max = fminf(fminf(f0x, f0y), fminf(f0z, tbest));
  c1min = fmaxf(fmaxf(n1x, n1y), fmaxf(n1z, 0.0f));
  c1max = fminf(fminf(f1x, f1y), fminf(f1z, tbest));
}

// Slab test of one padded box [lo, hi] against the segment [0, tmax] with the ray's precomputed reciprocals — the
// same arithmetic and padding as the node boxes (origin inside counts as a hit).
NRB_DI bool box_hit(const RayPre &p, const float lo[3], const float hi[3], float tmax) {
  float ax = fmaf(lo[0], p.idx, -p.oodx), bx = fmaf(hi[0], p.idx, -p.oodx);
  float ay = fmaf(lo[1], p.idy, -p.oody), by = fmaf(hi[1], p.idy, -p.oody);
  float az = fmaf(lo[2], p.idz, -p.oodz), bz = fmaf(hi[2], p.idz, -p.oodz);
  float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
  float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
  return t1 >= t0;
}

// ANY = true: return at the first hit with toi <= tmax (shadow rays vs opaque geometry).
// ANY = false: closest hit with toi < tmax (strict, best_first_search keeps the first of equals).
template <bool HAS_SHAPES, bool ANY, int FMT>
NRB_DI bool traverse(const SceneView &sc, int root, V3 o, V3 d, float tmax, Hit &hit) {
  int stack[kStackSize];
  int sp = 0;
  stack[0] = kEmpty;
  int node = root;
  bool found = false;
  constexpr int L = node_layout(FMT);
  const RayPre pre = ray_pre_for<L>(sc, o, d);
  const RayAux<L> aux = ray_aux<L>(pre);
  float tbest = tmax;
#ifdef NRB_COUNT_VISITS
  unsigned dbg_n = 0, dbg_t = 0;
#endif

  while (node != kEmpty) {
    // ---- inner nodes: one 64-byte record = both children's boxes ----
    while ((unsigned)node < (unsigned)kEmpty) {
#ifdef NRB_COUNT_VISITS
      ++dbg_n;
#endif
      const NodeRec<L> n = load_node<L>(sc, node);
      float c0min, c0max, c1min, c1max;
      test_children(n, pre, aux, tbest, c0min, c0max, c1min, c1max);
      bool h0 = c0max >= c0min, h1 = c1max >= c1min;
      if (!h0 && !h1) {
        node = stack[sp--];
      } else {
        node = h0 ? n.c0 : n.c1;
        if (h0 && h1) {
          int far = n.c1;
          if (!ANY && c1min < c0min) {  // any-hit needs no front-to-back order
            far = n.c0;
            node = n.c1;
          }
          stack[++sp] = far;
        }
      }
    }
    // ---- leaves ----
    while (node < 0) {
      uint32_t code = (uint32_t)~node;
      uint32_t first = code >> 3, cnt = ((code >> 1) & 3u) + 1u;
      if (HAS_SHAPES && (code & 1u)) {
        Inter it;
        if (cast_shape(sc.shapes[first], o, d, it) && (ANY ? it.toi <= tbest : it.toi < tbest)) {
          tbest = it.toi;
          hit.t = it.toi, hit.prim = kShapeBit | first, hit.u = it.u, hit.v = it.v;
          found = true;
          if (ANY) return true;
        }
      } else {
        const float4 *tp = reinterpret_cast<const float4 *>(sc.tris + first);
        for (uint32_t k = 0; k < cnt; ++k) {
#ifdef NRB_COUNT_VISITS
          ++dbg_t;
#endif
          float4 t0 = ld_tri(tp + 3 * k), t1 = ld_tri(tp + 3 * k + 1), t2 = ld_tri(tp + 3 * k + 2);
          float toi, bv, bw;
          if (cast_tri<ANY>(mk(t0.x, t0.y, t0.z), mk(t1.x, t1.y, t1.z), mk(t2.x, t2.y, t2.z), o, d, tbest, toi, bv, bw)) {
            tbest = toi;
            hit.t = toi, hit.prim = first + k, hit.u = bv, hit.v = bw;
            found = true;
            if (ANY) return true;
          }
        }
      }
      node = stack[sp--];
    }
  }
#ifdef NRB_COUNT_VISITS
  atomicAdd(&g_dbg[0], (unsigned long long)dbg_n);
  atomicAdd(&g_dbg[1], (unsigned long long)dbg_t);
  atomicAdd(&g_dbg[2], 1ull);
  atomicMax(&g_dbg[3], (unsigned long long)(dbg_n + dbg_t));
#endif
  return found;
}

// Resumable form of the same loop for the persistent trace kernel.  The lane's traversal state lives in
// `LaneTrav` (+ its stack) across calls; trav_run() advances it until this lane's traversal is complete
// (s.node == kEmpty) OR the number of lanes still traversing drops below `min_active` — then every lane
// returns so the warp can hand new rays to its idle lanes (persistent threads with dynamic fetch: the
// "active-ray compaction" of the north star).  `any` is a per-lane run-time flag here: shadow rays run
// any-hit under root_opaque and closest-hit under a transparent candidate's sub-root.
//
// Register budget: the node loop needs only (1/d, o/d, tbest, node, sp).  Everything touched per LEAF or per
// RAY — origin, direction, the best hit's (prim, u, v) — is parked in the lane's local-memory block in front
// of the traversal stack (slots kLm*), because the compiler would otherwise spill the loop's own operands.
struct LaneTrav {
  RayPre pre;
  int node, sp;
  float tbest;  // any: tmax (inclusive); closest: current best toi (exclusive)
  int leaf;     // format 2 (speculative loop): postponed leaf code (< 0), or 0
};
constexpr int kLmO = 0, kLmD = 3, kLmPrim = 6, kLmU = 7, kLmV = 8, kLmSentinel = 9;  // then the stack proper
constexpr int kLmSize = kLmSentinel + 1 + kStackSize;

NRB_DI V3 lm_vec(const int *lm, int at) {
  return mk(__int_as_float(lm[at]), __int_as_float(lm[at + 1]), __int_as_float(lm[at + 2]));
}
NRB_DI void lm_set_ray(int *lm, V3 o, V3 d) {
  lm[kLmO] = __float_as_int(o.x), lm[kLmO + 1] = __float_as_int(o.y), lm[kLmO + 2] = __float_as_int(o.z);
  lm[kLmD] = __float_as_int(d.x), lm[kLmD + 1] = __float_as_int(d.y), lm[kLmD + 2] = __float_as_int(d.z);
}
NRB_DI void lm_set_hit(int *lm, uint32_t prim, float u, float v) {
  lm[kLmPrim] = (int)prim, lm[kLmU] = __float_as_int(u), lm[kLmV] = __float_as_int(v);
}

NRB_DI void trav_start(LaneTrav &s, int *lm, int root, float tlimit) {
  lm[kLmSentinel] = kEmpty;
  lm[kLmPrim] = (int)kMiss;
  s.sp = kLmSentinel;
  s.node = root;
  s.tbest = tlimit;
  s.leaf = 0;
}

// One leaf of the traversal: up to 4 triangles (or one analytic shape).  Returns true if an any-hit query is finished.
template <bool HAS_SHAPES>
NRB_DI bool trav_leaf(const SceneView &sc, int leaf, int *lm, bool any, float &tbest) {
  const uint32_t code = (uint32_t)~leaf;
  const uint32_t first = code >> 3, cnt = ((code >> 1) & 3u) + 1u;
  bool hit_any = false;
  const V3 o = lm_vec(lm, kLmO), d = lm_vec(lm, kLmD);
  if (HAS_SHAPES && (code & 1u)) {
    Inter it;
    if (cast_shape(sc.shapes[first], o, d, it) && (any ? it.toi <= tbest : it.toi < tbest)) {
      tbest = it.toi;
      lm_set_hit(lm, kShapeBit | first, it.u, it.v);
      hit_any = any;
    }
  } else {
    const float4 *tp = reinterpret_cast<const float4 *>(sc.tris + first);
    for (uint32_t k = 0; k < cnt && !hit_any; ++k) {
      float4 t0 = ld_tri(tp + 3 * k), t1 = ld_tri(tp + 3 * k + 1), t2 = ld_tri(tp + 3 * k + 2);
      float toi, bv, bw;
      if (cast_tri_rt(mk(t0.x, t0.y, t0.z), mk(t1.x, t1.y, t1.z), mk(t2.x, t2.y, t2.z), o, d, tbest, any, toi, bv, bw)) {
        tbest = toi;
        lm_set_hit(lm, first + k, bv, bw);
        hit_any = any;
      }
    }
  }
  return hit_any;
}

// One node visit: both children tested, the nearer one entered, the farther one pushed.
// PREDICATED: the push / pop as predicated PTX instead of branches (used under the speculative loop, where it is worth 3 % on
// C4; under the plain loop it is neutral on C3 and costs 1 % on C5).
template <int L, bool PREDICATED>
NRB_DI void trav_visit(const SceneView &sc, const RayPre &pre, const RayAux<L> &aux, float tbest, int &node, int &sp, int *lm) {
  const NodeRec<L> n = load_node<L>(sc, node);
  float c0min, c0max, c1min, c1max;
  test_children(n, pre, aux, tbest, c0min, c0max, c1min, c1max);
  if (PREDICATED) {
    // lm[sp] is the pop source when neither child is hit and — after the increment — the push target when both are, so one
    // address serves both.
    const size_t lml = __cvta_generic_to_local(lm);
    asm volatile(
        "{\n"
        " .reg .pred h0, h1, sw, both, none, t1, nh0;\n"
        " .reg .b32 nearc, farc;\n"
        " .reg .b64 ad;\n"
        " setp.ge.f32 h0, %3, %2;\n"
        " setp.ge.f32 h1, %5, %4;\n"
        " setp.lt.f32 sw, %4, %2;\n"
        " and.pred both, h0, h1;\n"
        " or.pred none, h0, h1;\n"
        " not.pred none, none;\n"
        " not.pred nh0, h0;\n"
        " and.pred t1, both, sw;\n"
        " or.pred t1, t1, nh0;\n"
        " selp.b32 nearc, %7, %6, t1;\n"
        " selp.b32 farc, %6, %7, t1;\n"
        " @both add.s32 %1, %1, 1;\n"
        " mad.wide.s32 ad, %1, 4, %8;\n"
        " @both st.local.b32 [ad], farc;\n"
        " @none ld.local.b32 nearc, [ad];\n"
        " @none add.s32 %1, %1, -1;\n"
        " mov.b32 %0, nearc;\n"
        "}"
        : "=r"(node), "+r"(sp)
        : "f"(c0min), "f"(c0max), "f"(c1min), "f"(c1max), "r"(n.c0), "r"(n.c1), "l"(lml)
        : "memory");
    return;
  }
  bool h0 = c0max >= c0min, h1 = c1max >= c1min;
  if (!h0 && !h1) {
    node = lm[sp--];
  } else {
    node = h0 ? n.c0 : n.c1;
    if (h0 && h1) {
      int far = n.c1;
      if (c1min < c0min) {
        far = n.c0;
        node = n.c1;
      }
      lm[++sp] = far;
    }
  }
}

// Format 0 (coherent scenes): plain while-while — descend until a leaf, test it, repeat.
// Format 2 (scenes beyond L2, whose rays diverge): SPECULATIVE while-while (Aila & Laine, HPG 2009).  In hair only ~11 of 32
// lanes are in the node loop at any time (ncu, C4): a lane that reaches its leaf waits there for the slowest lane of the warp.
// Here it postpones that leaf, pops the next node and keeps descending — with a not-yet-shrunk interval, which is conservative —
// until no lane of the loop is still looking for its first leaf; then the postponed leaves (and a second one, if the lane
// sits on one) are tested together.
template <bool HAS_SHAPES, int FMT>
NRB_DI void trav_run(const SceneView &sc, LaneTrav &s, int *lm, bool any, int min_active) {
  constexpr int L = node_layout(FMT);
  int node = s.node, sp = s.sp;
  float tbest = s.tbest;
  const RayAux<L> aux = ray_aux<L>(s.pre);
  if (!speculative_loop(FMT)) {
    while (node != kEmpty) {
      while ((unsigned)node < (unsigned)kEmpty) trav_visit<L, false>(sc, s.pre, aux, tbest, node, sp, lm);
      while (node < 0) node = trav_leaf<HAS_SHAPES>(sc, node, lm, any, tbest) ? kEmpty : lm[sp--];
      if (__popc(__activemask()) < min_active) break;  // dynamic fetch: let the warp refill its idle lanes
    }
  } else {
    int leaf = s.leaf;  // postponed leaf (< 0) or 0
    while (node != kEmpty || leaf < 0) {
      while ((unsigned)node < (unsigned)kEmpty) {
        trav_visit<L, true>(sc, s.pre, aux, tbest, node, sp, lm);
        if (node < 0 && leaf == 0) {  // first leaf: postpone it and go on with the next node
          leaf = node;
          node = lm[sp--];
        }
        if (!__any_sync(__activemask(), leaf == 0)) break;  // every lane still descending holds a leaf: test them now
      }
      if (node < 0 && leaf == 0) {  // arrived on a leaf without descending (a root that is a leaf, or a popped leaf)
        leaf = node;
        node = lm[sp--];
      }
      while (leaf < 0) {
        if (trav_leaf<HAS_SHAPES>(sc, leaf, lm, any, tbest)) {
          node = kEmpty;
          leaf = 0;
          break;
        }
        leaf = 0;
        if (node < 0) {  // the lane stopped on a second leaf
          leaf = node;
          node = lm[sp--];
        }
      }
      if (__popc(__activemask()) < min_active) break;
    }
    s.leaf = leaf;
  }
  s.node = node, s.sp = sp, s.tbest = tbest;
}

// Hands queue entries to the idle lanes of a persistent warp.  The warp keeps a private pool
// [pool_base, pool_base + pool_left) of entries taken from the global cursor 32 * kFetchPackets at a time
// (one atomic per chunk); idle lanes get consecutive indices.
// Entries a warp takes per cursor atomic: kFetchPackets packets while the queue gives every resident warp several
// fetches, single packets below that so the last fetches of a small launch balance better (tile-sharded frames).
NRB_DI uint32_t fetch_chunk(uint32_t count, uint32_t small_queue) { return count >= small_queue ? 32u * kFetchPackets : 32u; }

struct RayPool {  // one per warp, in shared memory (keeps three registers out of the traversal loop)
  uint32_t base, left, more;  // more: the global cursor may still have entries
};
constexpr uint32_t kNoRay = 0xFFFFFFFFu;

NRB_DI void pool_reset(RayPool *pool) {
  __syncwarp();
  if (lane_id() == 0) pool->base = 0u, pool->left = 0u, pool->more = 1u;
  __syncwarp();
}

NRB_DI bool pool_empty(const RayPool *pool) { return pool->more == 0u && pool->left == 0u; }

NRB_DI uint32_t pool_assign(RayPool *pool, bool idle, uint32_t *cursor, uint32_t count, uint32_t chunk) {
  const uint32_t lane = lane_id();
  const uint32_t mask = __ballot_sync(0xFFFFFFFFu, idle);
  uint32_t mine = kNoRay;
  if (mask == 0u) return mine;
  uint32_t base = pool->base, left = pool->left, more = pool->more;
  __syncwarp();
  const uint32_t need = __popc(mask), rank = __popc(mask & ((1u << lane) - 1u));
  uint32_t take = min(need, left);
  if (idle && rank < take) mine = base + rank;
  base += take, left -= take;
  if (take < need && more) {
    uint32_t nb = 0;
    if (lane == 0) nb = atomicAdd(cursor, chunk);
    nb = __shfl_sync(0xFFFFFFFFu, nb, 0);
    if (nb >= count) {
      more = 0u;
    } else {
      base = nb, left = min(chunk, count - nb);
      const uint32_t take2 = min(need - take, left);
      if (idle && rank >= take && rank < take + take2) mine = base + (rank - take);
      base += take2, left -= take2;
    }
  }
  if (lane == 0) pool->base = base, pool->left = left, pool->more = more;
  __syncwarp();
  return mine;
}

// SceneNode.nmap (src/scene_node.rs:60-70): defined after the surface reconstruction below.
NRB_DI void nmap_closest(const SceneView &sc, V3 o, V3 d, Hit &hit);

// ---------------------------------------------------------------------------------------------
// K2 — closest hit of one ray (Scene::trace's best_first_search, src/scene.rs:164-166)
// ---------------------------------------------------------------------------------------------
template <bool HAS_SHAPES, int FMT>
NRB_DI float4 closest_hit(const SceneView &sc, V3 o, V3 d) {
  Hit hit;
  hit.t = 3.402823466e+38f, hit.prim = kMiss, hit.u = hit.v = 0.0f;
  if (HAS_SHAPES) {
    // planes have infinite AABBs: always tested, never in the BVH
    for (int p = 0; p < sc.n_planes; ++p) {
      int si = sc.planes[p];
      Inter it;
      if (cast_shape(sc.shapes[si], o, d, it) && it.toi < hit.t) {
        hit.t = it.toi, hit.prim = kShapeBit | (uint32_t)si, hit.u = 0.0f, hit.v = 0.0f;
      }
    }
  }
  if (sc.root_all != kEmpty) traverse<HAS_SHAPES, false, FMT>(sc, sc.root_all, o, d, hit.t, hit);
  if (HAS_SHAPES && sc.n_nmap > 0) nmap_closest(sc, o, d, hit);
  return make_float4(hit.t, __uint_as_float(hit.prim), hit.u, hit.v);
}

// ---------------------------------------------------------------------------------------------
// surface reconstruction shared by shade and the shadow filter
// ---------------------------------------------------------------------------------------------
struct Surface {
  V3 n;
  float u, v;
  bool has_uv;
  int node;
  int material;  // triangles carry a copy of NodeInfo.material (Tri.t1.w); -1: look it up in NodeInfo
};

template <bool HAS_SHAPES>
NRB_DI void reconstruct(const SceneView &sc, V3 o, V3 d, uint32_t prim, float bu, float bv, Surface &s) {
  if (HAS_SHAPES && (prim & kShapeBit)) {
    const Shape &sh = sc.shapes[prim & ~kShapeBit];
    Inter it;
    it.n = mk(0, 0, 0), it.u = it.v = 0.0f, it.has_uv = false;
    cast_shape(sh, o, d, it);
    s.n = it.n, s.u = it.u, s.v = it.v, s.has_uv = it.has_uv, s.node = sh.node, s.material = -1;
    return;
  }
  const float4 *tp = reinterpret_cast<const float4 *>(sc.tris + prim);
  float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
  V3 e1 = mk(t1.x, t1.y, t1.z), e2 = mk(t2.x, t2.y, t2.z);
  V3 n = cross(e1, e2);
  float t = dot(o - mk(t0.x, t0.y, t0.z), n);  // same expression as cast_tri: normal faces the ray origin
  V3 nn = normalize(n);
  s.n = t < 0.0f ? -nn : nn;
  const float2 *uvp = reinterpret_cast<const float2 *>(sc.tri_uvs + prim);
  const float2 uv0 = __ldg(uvp), uv1 = __ldg(uvp + 1), uv2 = __ldg(uvp + 2);
  float bw0 = -bu - bv + 1.0f;
  s.u = uv0.x * bw0 + uv1.x * bu + uv2.x * bv;
  s.v = uv0.y * bw0 + uv1.y * bu + uv2.y * bv;
  s.has_uv = true;
  s.node = __float_as_int(t0.w);
  s.material = __float_as_int(t1.w);
}

// Material::ambiant (src/phong_material.rs:39-70, normal_material.rs:7-15, uv_material.rs:8-21)
NRB_DI float4 mat_ambiant(const SceneView &sc, const Material &m, const Surface &s) {
  if (m.kind == NRB_MAT_NORMAL)
    return make_float4((1.0f + s.n.x) / 2.0f, (1.0f + s.n.y) / 2.0f, (1.0f + s.n.z) / 2.0f, 1.0f);
  if (m.kind == NRB_MAT_UV) return s.has_uv ? make_float4(s.u, s.v, 0.0f, 1.0f) : make_float4(0, 0, 0, 0);
  if (s.has_uv) {
    float4 tc = make_float4(1, 1, 1, 1);
    if (m.tex >= 0) {
      tc = tex_sample(sc, m.tex, s.u, s.v);
      tc.w = 1.0f;
    }
    if (m.alpha_tex >= 0) tc.w = tex_sample(sc, m.alpha_tex, s.u, s.v).w;
    return make_float4(m.ambient[0] * tc.x, m.ambient[1] * tc.y, m.ambient[2] * tc.z, tc.w);
  }
  return make_float4(m.ambient[0], m.ambient[1], m.ambient[2], 1.0f);
}

// ---------------------------------------------------------------------------------------------
// SceneNode.nmap — the "normal map" of the reference is a depth shift (src/scene_node.rs:60-70): after the node's own
// cast, `toi -= mean(nmap.sample(uv).rgb)`.  The shifted toi is what best_first_search compares, so a node with a
// depth-shift texture cannot live in the flat trees: it keeps its own sub-root and is resolved after the rest of the
// scene, in the order and with the pruning of the reference's search — nodes are popped by increasing AABB entry
// distance and the search stops at the first one whose entry distance is >= the best cost so far;
// `Candidate.lo/hi` hold the REFERENCE's node AABB (centre transformed, half extents times |R|) for that purpose.
// The loader never enables nmap (examples/loader3d.rs:553); this is the general, non-resumable path.
// ---------------------------------------------------------------------------------------------
NRB_DI bool aabb_toi_solid(const float lo[3], const float hi[3], V3 o, V3 d, float &toi) {  // AABB::toi_with_ray(.., solid = true), SURVEY B.3
  float tmin = 0.0f, tmax = 3.402823466e+38f;
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    const float oi = comp(o, ax), di = comp(d, ax);
    if (di == 0.0f) {
      if (oi < lo[ax] || oi > hi[ax]) return false;
    } else {
      const float inv = 1.0f / di;
      float t0 = (lo[ax] - oi) * inv, t1 = (hi[ax] - oi) * inv;
      if (t0 > t1) {
        const float tmp = t0;
        t0 = t1, t1 = tmp;
      }
      tmin = fmaxf(tmin, t0);
      tmax = fminf(tmax, t1);
      if (tmin > tmax) return false;
    }
  }
  toi = tmin;
  return true;
}

// Closest hit of one nmap node with its toi shifted; false if the node is missed.
NRB_DI bool nmap_cast(const SceneView &sc, const Candidate &nm, V3 o, V3 d, Hit &h) {
  h.t = 3.402823466e+38f, h.prim = kMiss, h.u = h.v = 0.0f;
  if (!traverse<true, false, 0>(sc, nm.root, o, d, 3.402823466e+38f, h)) return false;
  Surface sf;
  reconstruct<true>(sc, o, d, h.prim, h.u, h.v, sf);
  if (sf.has_uv) {
    const float4 c = tex_sample(sc, sc.node_info[nm.node].nmap_tex, sf.u, sf.v);
    h.t -= (c.x + c.y + c.z) / 3.0f;
  }
  return true;
}

NRB_DI void nmap_closest(const SceneView &sc, V3 o, V3 d, Hit &hit) {
  uint32_t done = 0u;
  for (int it = 0; it < sc.n_nmap; ++it) {
    int next = -1;
    float next_tb = 3.402823466e+38f;
    for (int j = 0; j < sc.n_nmap; ++j) {
      if ((done >> j) & 1u) continue;
      float tb;
      if (!aabb_toi_solid(sc.nmaps[j].lo, sc.nmaps[j].hi, o, d, tb)) {
        done |= 1u << j;
        continue;
      }
      if (tb < next_tb) next_tb = tb, next = j;
    }
    if (next < 0 || !(next_tb < hit.t)) break;  // popped cost >= best cost: the search ends
    done |= 1u << next;
    Hit h;
    if (nmap_cast(sc, sc.nmaps[next], o, d, h) && h.t < hit.t) hit = h;
  }
}

// Shadow query over the nmap nodes (src/scene.rs:304-339 with the shifted toi).  Returns true if occluded.
NRB_DI bool nmap_shadow(const SceneView &sc, V3 o, V3 d, float tmax, V3 &filter) {
  for (int j = 0; j < sc.n_nmap; ++j) {
    const Candidate nm = sc.nmaps[j];
    float tb;
    if (!aabb_toi_solid(nm.lo, nm.hi, o, d, tb)) continue;
    Hit h;
    if (!nmap_cast(sc, nm, o, d, h) || !(h.t <= tmax)) continue;
    Surface sf;
    reconstruct<true>(sc, o, d, h.prim, h.u, h.v, sf);
    const NodeInfo ni = sc.node_info[nm.node];
    const float4 c = mat_ambiant(sc, sc.materials[ni.material], sf);
    const float alpha = c.w * ni.alpha;
    if (!(alpha < 1.0f)) return true;
    const float k = 1.0f - alpha;
    filter = mk(filter.x * c.x * k, filter.y * c.y * k, filter.z * c.z * k);
  }
  return false;
}

// ---------------------------------------------------------------------------------------------
// K3 — shadow rays (Scene::intersects_ray, src/scene.rs:147-161 + :285-339)
// Outcome: any SceneNode whose closest hit is opaque with toi <= maxtoi => shadowed;
// otherwise filter = product over nodes whose closest hit within maxtoi is transparent.
// Opaque-certain geometry lives under root_opaque (any-hit is exact for it); every node that can be
// transparent has its own sub-root and is resolved by its own closest hit.
// ---------------------------------------------------------------------------------------------
template <bool HAS_SHAPES, int FMT>
NRB_DI bool shadow_candidate(const SceneView &sc, int root, int node_id, V3 o, V3 d, float tmax, V3 &filter) {
  Hit h;
  h.t = 0, h.prim = kMiss, h.u = h.v = 0;
  // closest hit of this node with toi <= tmax
  if (!traverse<HAS_SHAPES, false, FMT>(sc, root, o, d, nextafterf(tmax, 3.402823466e+38f), h)) return false;
  Surface s;
  reconstruct<HAS_SHAPES>(sc, o, d, h.prim, h.u, h.v, s);
  const NodeInfo ni = sc.node_info[node_id];
  float4 c = mat_ambiant(sc, sc.materials[ni.material], s);
  float alpha = c.w * ni.alpha;
  if (alpha < 1.0f) {
    float k = 1.0f - alpha;
    filter = mk(filter.x * c.x * k, filter.y * c.y * k, filter.z * c.z * k);
    return false;
  }
  return true;  // opaque
}

// The transparent-candidate part of the shadow query: every candidate SceneNode whose box the segment
// crosses is resolved by its own closest hit (filter or occlusion).
template <bool HAS_SHAPES, int FMT>
NRB_DI bool shadow_candidates(const SceneView &sc, V3 o, V3 d, float tmax, V3 &filter) {
  bool occluded = false;
  if (sc.n_candidates > 0) {
    const RayPre pre = ray_pre(o, d);
    for (int c = 0; c < sc.n_candidates && !occluded; ++c) {
      const Candidate cd = sc.candidates[c];
      if (!box_hit(pre, cd.lo, cd.hi, tmax)) continue;  // bv cost of the candidate's box
      occluded = shadow_candidate<HAS_SHAPES, FMT>(sc, cd.root, cd.node, o, d, tmax, filter);
    }
  }
  return occluded;
}

template <bool HAS_SHAPES, int FMT>
NRB_DI void shadow_query(const SceneView &sc, V3 o, V3 d, float tmax, uint32_t pix, V3 contrib, float4 *accum) {
  bool occluded = false;
  V3 filter = mk(1, 1, 1);
  if (HAS_SHAPES) {
    for (int p = 0; p < sc.n_planes && !occluded; ++p) {
      int si = sc.planes[p];
      const Shape &sh = sc.shapes[si];
      if (sc.node_info[sh.node].flags & 1) {
        // transparent candidate plane: its (only) hit filters or occludes
        Inter it;
        if (cast_shape(sh, o, d, it) && it.toi <= tmax) {
          Surface s;
          s.n = it.n, s.u = it.u, s.v = it.v, s.has_uv = it.has_uv, s.node = sh.node, s.material = -1;
          const NodeInfo ni = sc.node_info[sh.node];
          float4 c = mat_ambiant(sc, sc.materials[ni.material], s);
          float alpha = c.w * ni.alpha;
          if (alpha < 1.0f) {
            float k = 1.0f - alpha;
            filter = mk(filter.x * c.x * k, filter.y * c.y * k, filter.z * c.z * k);
          } else {
            occluded = true;
          }
        }
      } else {
        Inter it;
        if (cast_shape(sh, o, d, it) && it.toi <= tmax) occluded = true;
      }
    }
  }
  if (!occluded && sc.root_opaque != kEmpty) {
    Hit h;
    occluded = traverse<HAS_SHAPES, true, FMT>(sc, sc.root_opaque, o, d, tmax, h);
  }
  if (!occluded) occluded = shadow_candidates<HAS_SHAPES, FMT>(sc, o, d, tmax, filter);
  if (HAS_SHAPES && !occluded && sc.n_nmap > 0) occluded = nmap_shadow(sc, o, d, tmax, filter);
  if (!occluded) accum_add(accum, pix, cmul(contrib, filter));
}

// ---------------------------------------------------------------------------------------------
// The persistent trace kernel: ONE launch per wave drains the shadow queue of the previous wave
// (any-hit + transparent filter, adds into the pixel accumulator) and the ray queue of this wave
// (closest hit -> hit records).  Grid = SMs x resident CTAs.  Every lane owns one ray at a time; when
// fewer than `min_active` lanes of a warp are still traversing, the warp pauses and refills its idle
// lanes from the queue (pool_assign), so SIMD utilisation does not decay to the slowest ray of a packet.
// Even CTAs start on the shadow queue, odd CTAs on the ray queue, then swap: in small (latency-bound)
// waves both queues progress at once.  PRIMARY: wave 0 generates its rays from the sample slot instead
// of reading a queue.
// ---------------------------------------------------------------------------------------------

// Shadow-ray state machine.  Phase -1: any-hit under root_opaque.  Phase c >= 0: closest hit
// of transparent candidate c (its own sub-root), resolved through Material::ambiant.  Returns true if a
// traversal was started, false if the ray is finished (occluded, or accumulated into its pixel).
template <bool HAS_SHAPES, int FMT>
NRB_DI bool shadow_advance(const SceneView &sc, const ShadowQueue &sq, uint32_t idx, LaneTrav &s, int *lm, int &cand,
                           bool first, bool reverse, float4 *accum) {
  // The transparent filter is folded into the entry's contribution in place (the entry belongs to this lane),
  // so the state carried across trav_run calls is just (idx, cand).
  // reverse: the opaque any-hit phase walks the segment from its LIGHT end.  Occlusion of a segment does not
  // depend on the direction it is walked in, but rays that leave one point light together visit the same nodes
  // in the same order, like primary rays, instead of starting at scattered surface points.  The candidate
  // phases need the hit closest to the surface and always use the forward ray.
  const uint32_t prim = first ? kMiss : (uint32_t)lm[kLmPrim];
  if (first || (cand < 0 && prim == kMiss && reverse && sc.n_candidates > 0)) {
    const float4 a = sq.a[idx], b = sq.b[idx];
    V3 o = mk(a.x, a.y, a.z), d = mk(b.x, b.y, b.z);
    if (first && reverse && sc.root_opaque != kEmpty) o = o + d * a.w, d = -d;
    lm_set_ray(lm, o, d);
    s.pre = ray_pre_for<node_layout(FMT)>(sc, o, d);
  }
  if (first) {
    cand = -1;
    if (sc.root_opaque != kEmpty) {
      trav_start(s, lm, sc.root_opaque, sq.a[idx].w);
      return true;
    }
  } else if (cand < 0) {
    if (prim != kMiss) return false;  // an opaque occluder
  } else if (prim != kMiss) {
    // closest hit of this candidate node within tmax: filter or occlude (src/scene.rs:313-337)
    Surface sf;
    reconstruct<HAS_SHAPES>(sc, lm_vec(lm, kLmO), lm_vec(lm, kLmD), prim, __int_as_float(lm[kLmU]), __int_as_float(lm[kLmV]), sf);
    const NodeInfo ni = sc.node_info[sc.candidates[cand].node];
    float4 c = mat_ambiant(sc, sc.materials[ni.material], sf);
    float alpha = c.w * ni.alpha;
    if (!(alpha < 1.0f)) return false;
    float k = 1.0f - alpha;
    float4 e = sq.c[idx];
    sq.c[idx] = make_float4(e.x * c.x * k, e.y * c.y * k, e.z * c.z * k, e.w);
  }
  if (cand + 1 < sc.n_candidates) {
    const float tmax = sq.a[idx].w;
    for (++cand; cand < sc.n_candidates; ++cand) {
      const Candidate cd = sc.candidates[cand];
      // bv cost of the candidate's box (float planes: the grid layout's s.pre is in cells)
      if (!box_hit(node_layout(FMT) == 3 ? ray_pre(lm_vec(lm, kLmO), lm_vec(lm, kLmD)) : s.pre, cd.lo, cd.hi, tmax)) continue;
      trav_start(s, lm, cd.root, nextafterf(tmax, 3.402823466e+38f));  // closest hit with toi <= tmax
      return true;
    }
  }
  const float4 b = sq.b[idx], c = sq.c[idx];
  accum_add(accum, __float_as_uint(b.w), mk(c.x, c.y, c.z));
  return false;
}

// Planes of the shadow query (always tested).  Returns true if the ray is occluded.
NRB_DI bool shadow_planes(const SceneView &sc, const ShadowQueue &sq, uint32_t idx, V3 o, V3 d, float tmax) {
  for (int p = 0; p < sc.n_planes; ++p) {
    const Shape &sh = sc.shapes[sc.planes[p]];
    Inter it;
    if (!(cast_shape(sh, o, d, it) && it.toi <= tmax)) continue;
    const NodeInfo ni = sc.node_info[sh.node];
    if (!(ni.flags & 1)) return true;
    // transparent candidate plane: its (only) hit filters or occludes
    Surface sf;
    sf.n = it.n, sf.u = it.u, sf.v = it.v, sf.has_uv = it.has_uv, sf.node = sh.node, sf.material = -1;
    float4 c = mat_ambiant(sc, sc.materials[ni.material], sf);
    float alpha = c.w * ni.alpha;
    if (!(alpha < 1.0f)) return true;
    float k = 1.0f - alpha;
    float4 e = sq.c[idx];
    sq.c[idx] = make_float4(e.x * c.x * k, e.y * c.y * k, e.z * c.z * k, e.w);
  }
  return false;
}

template <bool HAS_SHAPES, int FMT>
NRB_DI void drain_shadow(const SceneView &sc, const ShadowQueue &sq, float4 *accum, WaveCounters *wc_shadow, int min_active,
                         bool reverse, uint32_t small_queue, RayPool *pool, int *lm) {
  const uint32_t count = min(wc_shadow->n_shadow, sq.capacity);
  LaneTrav s;
  s.node = kEmpty;
  pool_reset(pool);
  bool active = false;
  uint32_t idx = 0;
  int cand = -1;
  while (true) {
    const uint32_t mine = pool_assign(pool, !active, &wc_shadow->fetch_shadow, count, fetch_chunk(count, small_queue));
    if (mine != kNoRay) {
      idx = mine;
      bool occluded = false;
      if (HAS_SHAPES) {
        const float4 a = sq.a[idx], b = sq.b[idx];
        occluded = shadow_planes(sc, sq, idx, mk(a.x, a.y, a.z), mk(b.x, b.y, b.z), a.w);
      }
      active = !occluded && shadow_advance<HAS_SHAPES, FMT>(sc, sq, idx, s, lm, cand, true, reverse, accum);
    }
    if (__ballot_sync(0xFFFFFFFFu, active) == 0u) {
      if (pool_empty(pool)) break;
      continue;
    }
    if (active) {
      trav_run<HAS_SHAPES, FMT>(sc, s, lm, cand < 0, min_active);
      if (s.node == kEmpty) active = shadow_advance<HAS_SHAPES, FMT>(sc, sq, idx, s, lm, cand, false, reverse, accum);
    }
  }
}

template <bool HAS_SHAPES, bool PRIMARY, int FMT>
NRB_DI void drain_closest(const SceneView &sc, const FrameParams &fp, const RayQueue &q, float4 *hits,
                          WaveCounters *wc_closest, uint32_t slot_lo, uint32_t n_slots, int min_active, uint32_t small_queue,
                          RayPool *pool, int *lm) {
  const uint32_t count = PRIMARY ? n_slots : wc_closest->n_rays;
  LaneTrav s;
  s.node = kEmpty;
  pool_reset(pool);
  bool active = false;
  uint32_t idx = 0;
  while (true) {
    const uint32_t mine = pool_assign(pool, !active, &wc_closest->fetch_closest, count, fetch_chunk(count, small_queue));
    if (mine != kNoRay) {
      idx = mine;
      bool valid = true;
      V3 o, d;
      if (PRIMARY) {
        uint32_t gid;
        valid = primary_ray(fp, slot_lo + idx, o, d, gid);
      } else {
        const float4 a = q.a[idx], b = q.b[idx];
        o = mk(a.x, a.y, a.z), d = mk(a.w, b.x, b.y);
      }
      if (!valid) {
        // slots outside the image (ragged tiles) are marked so shade skips them
        hits[idx] = make_float4(0.0f, __uint_as_float(kSkip), 0.0f, 0.0f);
      } else {
        lm_set_ray(lm, o, d);
        s.pre = ray_pre_for<node_layout(FMT)>(sc, o, d);
        trav_start(s, lm, sc.root_all, 3.402823466e+38f);
        if (HAS_SHAPES) {
          // planes have infinite AABBs: always tested, never in the BVH
          for (int p = 0; p < sc.n_planes; ++p) {
            int si = sc.planes[p];
            Inter it;
            if (cast_shape(sc.shapes[si], o, d, it) && it.toi < s.tbest) {
              s.tbest = it.toi;
              lm_set_hit(lm, kShapeBit | (uint32_t)si, 0.0f, 0.0f);
            }
          }
        }
        active = true;
      }
    }
    if (__ballot_sync(0xFFFFFFFFu, active) == 0u) {
      if (pool_empty(pool)) break;
      continue;
    }
    if (active) {
      trav_run<HAS_SHAPES, FMT>(sc, s, lm, false, min_active);
      if (s.node == kEmpty) {
        const uint32_t prim = (uint32_t)lm[kLmPrim];
        const bool hit = prim != kMiss;
        hits[idx] = make_float4(s.tbest, __uint_as_float(prim), hit ? __int_as_float(lm[kLmU]) : 0.0f,
                                hit ? __int_as_float(lm[kLmV]) : 0.0f);
        active = false;
      }
    }
  }
}

template <bool HAS_SHAPES, bool PRIMARY, int FMT>
__global__ void __launch_bounds__(kTraceBlock, HAS_SHAPES ? 6 : kTraceMinBlocks)
    trace_kernel(SceneView sc, FrameParams fp, RayQueue q, float4 *hits, WaveCounters *wc_closest, uint32_t slot_lo,
                 uint32_t n_slots, ShadowQueue sq, float4 *accum, WaveCounters *wc_shadow, TraceOpts opts) {
  __shared__ RayPool pools[kTraceBlock / 32];
  RayPool *pool = &pools[threadIdx.x >> 5];
  int lm[kLmSize];  // the lane's ray, best hit and traversal stack (see LaneTrav)
  const bool closest_first = blockIdx.x & 1u;
  for (int phase = 0; phase < 2; ++phase) {
    if ((phase == 0) == closest_first) {
      if (wc_closest)
        drain_closest<HAS_SHAPES, PRIMARY, FMT>(sc, fp, q, hits, wc_closest, slot_lo, n_slots, opts.min_active_closest, opts.small_queue, pool, lm);
    } else {
      if (wc_shadow)
        drain_shadow<HAS_SHAPES, FMT>(sc, sq, accum, wc_shadow, opts.min_active_shadow, opts.reverse_shadow != 0, opts.small_queue, pool, lm);
    }
  }
}

// General form of the trace kernel for scenes the resumable loop does not cover (nodes with a depth-shift texture):
// same queues and counters, one 32-ray packet per fetch, every lane runs the whole query (closest_hit / shadow_query).
template <bool PRIMARY>
__global__ void __launch_bounds__(kTraceBlock, 3)
    trace_general_kernel(SceneView sc, FrameParams fp, RayQueue q, float4 *hits, WaveCounters *wc_closest, uint32_t slot_lo,
                         uint32_t n_slots, ShadowQueue sq, float4 *accum, WaveCounters *wc_shadow) {
  const uint32_t lane = lane_id();
  if (wc_shadow) {
    const uint32_t count = min(wc_shadow->n_shadow, sq.capacity);
    while (true) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&wc_shadow->fetch_shadow, 32u);
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (base >= count) break;
      const uint32_t i = base + lane;
      if (i < count) {
        const float4 a = sq.a[i], b = sq.b[i], c = sq.c[i];
        shadow_query<true, 0>(sc, mk(a.x, a.y, a.z), mk(b.x, b.y, b.z), a.w, __float_as_uint(b.w), mk(c.x, c.y, c.z), accum);
      }
    }
  }
  if (wc_closest) {
    const uint32_t count = PRIMARY ? n_slots : wc_closest->n_rays;
    while (true) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&wc_closest->fetch_closest, 32u);
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (base >= count) break;
      const uint32_t i = base + lane;
      if (i < count) {
        V3 o, d;
        bool valid = true;
        if (PRIMARY) {
          uint32_t gid;
          valid = primary_ray(fp, slot_lo + i, o, d, gid);
        } else {
          const float4 a = q.a[i], b = q.b[i];
          o = mk(a.x, a.y, a.z), d = mk(a.w, b.x, b.y);
        }
        hits[i] = valid ? closest_hit<true, 0>(sc, o, d) : make_float4(0.0f, __uint_as_float(kSkip), 0.0f, 0.0f);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K4 — shade + secondary-ray generation (Scene::trace body after the cast, src/scene.rs:168-192)
// ---------------------------------------------------------------------------------------------
struct RayState {  // RayWithEnergy (src/ray_with_energy.rs:4-8) + wavefront bookkeeping
  V3 o, d;
  float weight, energy, refr;
  uint32_t gid, path, depth;
};

NRB_DI RayState load_ray(const RayQueue &q, uint32_t i) {
  float4 ra = q.a[i], rb = q.b[i], rc = q.c[i];
  RayState r;
  r.o = mk(ra.x, ra.y, ra.z), r.d = mk(ra.w, rb.x, rb.y);
  r.weight = rb.z, r.energy = rb.w, r.refr = rc.x;
  r.gid = __float_as_uint(rc.y), r.path = __float_as_uint(rc.z), r.depth = __float_as_uint(rc.w);
  return r;
}

NRB_DI void store_ray(const RayQueue &q, uint32_t i, const RayState &r) {
  q.a[i] = make_float4(r.o.x, r.o.y, r.o.z, r.d.x);
  q.b[i] = make_float4(r.d.y, r.d.z, r.weight, r.energy);
  q.c[i] = make_float4(r.refr, __uint_as_float(r.gid), __uint_as_float(r.path), __uint_as_float(r.depth));
}

// Everything the emission steps need about one shaded hit.
struct Shaded {
  V3 pt, n;         // hit point, normal facing the ray origin
  V3 kd, ks;        // Kd (x) texture, Ks
  float shininess;
  float w_obj;      // weight of the surface's own colour: w * a' * (1 - mix)
  float alpha, a1;  // alpha = obj.w * node.alpha; a' = alpha == 1 ? 1 : alpha
  float refl_mix, refl_att, refr_coeff;
  uint32_t pix, ipt, smp;
  bool emit_sh, cull_sh, want_refl, want_refr, trunc_refl, trunc_refr;
};

// Material evaluation + the combine weights of Scene::trace (src/scene.rs:171-190).  Adds the
// background (miss) or the ambient term (hit) to the pixel; decides what the ray emits.
template <bool HAS_SHAPES>
NRB_DI void shade_eval(const SceneView &sc, const FrameParams &fp, const RayState &r, float4 h, bool active,
                       float4 *accum, Shaded &out) {
  const uint32_t prim = __float_as_uint(h.y);
  out.ipt = fdiv(r.gid, fp.div_spp), out.smp = r.gid - out.ipt * fp.spp;
  out.pix = active ? accum_index(fp, out.ipt) : 0u;
  const bool is_hit = active && prim != kMiss;
  if (active && !is_hit) {
    // miss -> background (src/scene.rs:169)
    accum_add(accum, out.pix, mk(sc.background[0], sc.background[1], sc.background[2]) * r.weight);
  }
  Surface s;
  s.n = mk(0, 0, 1), s.u = s.v = 0.0f, s.has_uv = false, s.node = 0, s.material = -1;
  NodeInfo ni;
  ni.material = 0, ni.refl_mix = 0, ni.refl_att = 0, ni.alpha = 1, ni.refr_coeff = 1, ni.flags = 0;
  float4 tex_color = make_float4(1, 1, 1, 1);
  float obj_w = 1.0f;
  V3 obj_rgb = mk(0, 0, 0);
  bool phong = false;
  out.pt = mk(0, 0, 0);
  out.kd = out.ks = mk(0, 0, 0);
  out.shininess = 0.0f;
  if (is_hit) {
    reconstruct<HAS_SHAPES>(sc, r.o, r.d, prim, h.z, h.w, s);
    out.pt = r.o + r.d * h.x;
    ni = sc.node_info[s.node];
    const Material m = sc.materials[s.material >= 0 ? s.material : ni.material];
    if (m.kind == NRB_MAT_PHONG) {
      // PhongMaterial::compute, ambient part (src/phong_material.rs:85-103)
      phong = true;
      if (s.has_uv && m.tex >= 0) tex_color = tex_sample(sc, m.tex, s.u, s.v);
      if (s.has_uv && m.alpha_tex >= 0) obj_w = tex_sample(sc, m.alpha_tex, s.u, s.v).w;
      obj_rgb = mk(m.ambient[0] * tex_color.x, m.ambient[1] * tex_color.y, m.ambient[2] * tex_color.z);
      out.kd = mk(m.diffuse[0] * tex_color.x, m.diffuse[1] * tex_color.y, m.diffuse[2] * tex_color.z);
      out.ks = mk(m.specular[0], m.specular[1], m.specular[2]);
      out.shininess = m.shininess;
    } else {
      float4 c = mat_ambiant(sc, m, s);  // Material::compute default (src/material.rs:8-16)
      obj_rgb = mk(c.x, c.y, c.z);
      obj_w = c.w;
    }
  }
  out.n = s.n;
  // combine weights (src/scene.rs:178-190): out = alpha==1 ? col : col*alpha + refr*(1-alpha),
  // col = obj*(1-mix) + refl*mix
  out.alpha = obj_w * ni.alpha;
  out.a1 = (out.alpha == 1.0f) ? 1.0f : out.alpha;
  out.w_obj = r.weight * out.a1 * (1.0f - ni.refl_mix);
  out.refl_mix = ni.refl_mix, out.refl_att = ni.refl_att, out.refr_coeff = ni.refr_coeff;
  if (is_hit) accum_add(accum, out.pix, obj_rgb * out.w_obj);
  // PhongMaterial::compute queries every light sample; a sample whose weight w * a' * (1 - mix) is exactly zero (a hit
  // on a fully transparent texel, a perfect mirror) adds exactly nothing, so its shadow query is not cast (counted apart)
  const bool lit = is_hit && phong && sc.shadow_samples > 0;
  out.emit_sh = lit && out.w_obj != 0.0f;
  out.cull_sh = lit && !out.emit_sh;
  // reflection (Scene::trace_reflection, src/scene.rs:196-218)
  bool want_refl = is_hit && ni.refl_mix != 0.0f && r.energy > 0.1f;
  out.trunc_refl = want_refl && (r.depth + 1u >= fp.max_depth);
  out.want_refl = want_refl && !out.trunc_refl;
  // refraction (Scene::trace_refraction, src/scene.rs:221-252)
  bool want_refr = is_hit && out.alpha != 1.0f;
  out.trunc_refr = want_refr && (r.depth + 1u >= fp.max_depth);
  out.want_refr = want_refr && !out.trunc_refr;
}

// One light sample of PhongMaterial::compute (src/phong_material.rs:108-143, src/light.rs:56-63):
// the shadow segment and the colour it adds if unoccluded (filter applied by the shadow query).
NRB_DI void light_sample(const SceneView &sc, const FrameParams &fp, const RayState &r, const Shaded &s, int li,
                         const Light &L, uint32_t k, float inv_ns, V3 &so, V3 &ldir, float &dist, V3 &c) {
  V3 pos = mk(L.pos[0], L.pos[1], L.pos[2]);
  if (L.radius != 0.0f) {
    uint32_t rnd[4];
    // r.path doubles per bounce and wraps after 32 of them: the key takes depth / 32 so deeper paths keep their own streams
    philox4x32_10(s.ipt, s.smp, r.path, ((uint32_t)li << 16) | (k & 0xFFFFu), fp.seed_lo,
                  fp.seed_hi ^ kStreamLight ^ ((r.depth >> 5) * 0x9E3779B9u), rnd);
    pos = pos + mk(u24(rnd[0]), u24(rnd[1]), u24(rnd[2])) * L.radius;
  }
  ldir = pos - s.pt;
  float len = sqrtf(dot(ldir, ldir));
  ldir = ldir * (1.0f / len);
  dist = len - 0.001f;
  float ndl = dot(ldir, s.n);
  float dcoeff = fmaxf(ndl, 0.0f);
  c = s.kd * dcoeff;
  V3 rl = normalize_fast(-ldir + s.n * (2.0f * ndl));
  float scoeff = -dot(rl, r.d);
  if (scoeff > 0.0f) c = c + s.ks * __powf(scoeff, s.shininess);  // ex2(Ns * lg2(s)): rel. error ~2e-7 * Ns
  c = cmul(mk(L.color[0], L.color[1], L.color[2]), c) * (inv_ns * s.w_obj);
  so = s.pt + ldir * 0.001f;
}

// Writes the ray's shadow_samples light samples into slots [sbase, sbase + shadow_samples).
// TRACE_ON_OVERFLOW (tail kernel): a sample that does not fit the queue is traced on the spot.
template <bool HAS_SHAPES, bool TRACE_ON_OVERFLOW, int FMT>
NRB_DI void emit_shadow_rays(const SceneView &sc, const FrameParams &fp, const RayState &r, const Shaded &s,
                             const ShadowQueue &sq, uint32_t sbase, Counters *ctr, float4 *accum) {
  uint32_t k_out = 0;
  for (int li = 0; li < sc.n_lights; ++li) {
    const Light L = sc.lights[li];
    uint32_t ns = L.racsample * L.racsample;
    float inv_ns = 1.0f / (float)ns;
    for (uint32_t k = 0; k < ns; ++k, ++k_out) {
      V3 so, ldir, c;
      float dist;
      light_sample(sc, fp, r, s, li, L, k, inv_ns, so, ldir, dist, c);
      uint32_t si = sbase + k_out;
      if (si < sq.capacity) {
        sq.a[si] = make_float4(so.x, so.y, so.z, dist);
        sq.b[si] = make_float4(ldir.x, ldir.y, ldir.z, __uint_as_float(s.pix));
        sq.c[si] = make_float4(c.x, c.y, c.z, 0.0f);
      } else if (TRACE_ON_OVERFLOW) {
        shadow_query<HAS_SHAPES, FMT>(sc, so, ldir, dist, s.pix, c, accum);
      } else {
        ctr->overflow = 1u;
      }
    }
  }
}

NRB_DI RayState reflect_ray(const RayState &r, const Shaded &s) {  // src/scene.rs:204-214
  RayState c;
  float dn = dot(r.d, s.n);
  c.d = r.d - s.n * (2.0f * dn);
  c.o = s.pt + c.d * 0.001f;
  c.weight = r.weight * s.a1 * s.refl_mix;
  c.energy = r.energy - s.refl_att;
  c.refr = r.refr;
  c.gid = r.gid, c.path = r.path * 2u, c.depth = r.depth + 1u;
  return c;
}

NRB_DI RayState refract_ray(const RayState &r, const Shaded &s) {  // src/scene.rs:229-248
  RayState c;
  float n1, n2;
  if (r.refr == 1.0f) {
    n1 = 1.0f, n2 = s.refr_coeff;
  } else {
    n1 = s.refr_coeff, n2 = 1.0f;
  }
  V3 along = s.n * dot(r.d, s.n);
  V3 tangent = r.d - along;
  c.d = normalize(along + tangent * (n2 / n1));
  c.o = s.pt + c.d * 0.001f;
  c.weight = r.weight * (1.0f - s.alpha);
  c.energy = r.energy;
  c.refr = n2;
  c.gid = r.gid, c.path = r.path * 2u + 1u, c.depth = r.depth + 1u;
  return c;
}

// Persistent grid; the ray count is read from device memory so the host can enqueue a wave without
// knowing how many rays the previous one produced.  Queue slots are reserved per BLOCK (warp totals
// meet in shared memory, one global atomic per block per queue and iteration) and the ray statistics
// live in registers until the kernel ends: same-address global atomics would otherwise serialise in L2.
struct ShadeShared {
  uint32_t cnt[2];   // [0] shadow entries, [1] secondary rays requested by this block in this iteration
  uint32_t base[2];  // their first slots in the global queues
};

NRB_DI void flush_stats(Counters *ctr, uint32_t c_shadow, uint32_t c_refl, uint32_t c_refr, uint32_t c_trunc, uint32_t c_culled) {
  c_shadow = __reduce_add_sync(0xFFFFFFFFu, c_shadow);
  c_culled = __reduce_add_sync(0xFFFFFFFFu, c_culled);
  c_refl = __reduce_add_sync(0xFFFFFFFFu, c_refl);
  c_refr = __reduce_add_sync(0xFFFFFFFFu, c_refr);
  c_trunc = __reduce_add_sync(0xFFFFFFFFu, c_trunc);
  if (lane_id() == 0) {
    if (c_shadow) atomicAdd(&ctr->rays_shadow, (unsigned long long)c_shadow);
    if (c_culled) atomicAdd(&ctr->rays_shadow_culled, (unsigned long long)c_culled);
    if (c_refl) atomicAdd(&ctr->rays_reflect, (unsigned long long)c_refl);
    if (c_refr) atomicAdd(&ctr->rays_refract, (unsigned long long)c_refr);
    if (c_trunc) atomicAdd(&ctr->paths_truncated, (unsigned long long)c_trunc);
  }
}

template <bool HAS_SHAPES, bool PRIMARY>
__global__ void __launch_bounds__(kShadeBlock, kShadeMinBlocks) shade_kernel(SceneView sc, FrameParams fp, RayQueue qin,
                                                                            const float4 *hits, WaveCounters *wc,
                                                                            uint32_t slot_lo, uint32_t n_slots,
                                                                            uint32_t lo, uint32_t hi, RayQueue qout,
                                                                            ShadowQueue sq, Counters *ctr,
                                                                            float4 *accum) {
  uint32_t *const tail_out = &wc[1].n_rays;  // next wave's ray count
  uint32_t *const tail_shadow = &wc[0].n_shadow;
  __shared__ ShadeShared sm;
  if (threadIdx.x < 2) sm.cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t end = min(hi, PRIMARY ? n_slots : wc[0].n_rays);
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t lane = lane_id();
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t S = (uint32_t)sc.shadow_samples;
  uint32_t c_shadow = 0, c_refl = 0, c_refr = 0, c_trunc = 0, c_culled = 0;

  // the hit record of the NEXT iteration is requested one iteration ahead: it is the only streaming (HBM) load of the
  // loop and its latency would otherwise open every iteration (14 % of the stall samples)
  float4 h_next = make_float4(0, 0, 0, 0);
  {
    const uint32_t i0 = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i0 < end) h_next = hits[i0];
  }
  for (uint32_t bb = lo + blockIdx.x * blockDim.x; bb < end; bb += stride) {  // block-uniform trip count
    const uint32_t i = bb + threadIdx.x;
    bool active = i < end;
    const float4 h = h_next;
    if (i + stride < end) h_next = hits[i + stride];
    RayState r;
    r.o = mk(0, 0, 0), r.d = mk(0, 0, 1);
    r.weight = 1.0f, r.energy = 1.0f, r.refr = 1.0f;  // RayWithEnergy::new (src/ray_with_energy.rs:11-13)
    r.gid = 0u, r.path = 1u, r.depth = 0u;
    if (PRIMARY) {
      active = active && __float_as_uint(h.y) != kSkip;
      if (active) primary_ray(fp, slot_lo + i, r.o, r.d, r.gid);  // regenerated from the slot, never stored
    } else if (active) {
      r = load_ray(qin, i);
    }
    Shaded s;
    shade_eval<HAS_SHAPES>(sc, fp, r, h, active, accum, s);
    c_trunc += (s.trunc_refl ? 1u : 0u) + (s.trunc_refr ? 1u : 0u);
    c_refl += s.want_refl ? 1u : 0u;
    c_refr += s.want_refr ? 1u : 0u;
    c_shadow += (s.emit_sh || s.cull_sh) ? S : 0u;
    c_culled += s.cull_sh ? S : 0u;

    // ---- reserve queue slots: warp ballots -> shared-memory totals -> one global atomic per block ----
    const uint32_t m_sh = __ballot_sync(0xFFFFFFFFu, s.emit_sh);
    const uint32_t m_rl = __ballot_sync(0xFFFFFFFFu, s.want_refl);
    const uint32_t m_rr = __ballot_sync(0xFFFFFFFFu, s.want_refr);
    uint32_t woff_sh = 0, woff_ch = 0;
    if (lane == 0) {
      uint32_t tsh = __popc(m_sh) * S, tch = __popc(m_rl) + __popc(m_rr);
      if (tsh) woff_sh = atomicAdd(&sm.cnt[0], tsh);
      if (tch) woff_ch = atomicAdd(&sm.cnt[1], tch);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t t = sm.cnt[0];
      sm.base[0] = t ? atomicAdd(tail_shadow, t) : 0u;
      sm.cnt[0] = 0;
    } else if (threadIdx.x == 32) {
      uint32_t t = sm.cnt[1];
      sm.base[1] = t ? atomicAdd(tail_out, t) : 0u;
      sm.cnt[1] = 0;
    }
    __syncthreads();
    const uint32_t sbase = sm.base[0] + __shfl_sync(0xFFFFFFFFu, woff_sh, 0) + __popc(m_sh & lt_mask) * S;
    const uint32_t cbase = sm.base[1] + __shfl_sync(0xFFFFFFFFu, woff_ch, 0) + __popc(m_rl & lt_mask) + __popc(m_rr & lt_mask);

    if (s.emit_sh) emit_shadow_rays<HAS_SHAPES, false, 0>(sc, fp, r, s, sq, sbase, ctr, accum);
    if (s.want_refl) {
      if (cbase < qout.capacity)
        store_ray(qout, cbase, reflect_ray(r, s));
      else
        ctr->overflow = 1u;
    }
    if (s.want_refr) {
      uint32_t fi = cbase + (s.want_refl ? 1u : 0u);
      if (fi < qout.capacity)
        store_ray(qout, fi, refract_ray(r, s));
      else
        ctr->overflow = 1u;
    }
  }
  flush_stats(ctr, c_shadow, c_refl, c_refr, c_trunc, c_culled);
}

// ---------------------------------------------------------------------------------------------
// Tail kernel.  Once a wave is small (a few thousand rays) every further wave costs the latency of its
// SLOWEST ray plus two launches, and a foliage chain needs ~10 of them.  Here each lane follows its own
// ray to the end instead — closest hit, shade, next bounce — so the tail costs the longest single
// chain (max of sums) instead of the sum of per-wave maxima.  Shadow rays are appended to the queue (and
// counter `wc_sh`) that already holds the last shade's untraced shadow rays; ONE launch traces them all afterwards; when a hit spawns both children the refraction ray is spilled to `qspill`
// (processed by the next tail launch).
// ---------------------------------------------------------------------------------------------
template <bool HAS_SHAPES, int FMT>
__global__ void __launch_bounds__(kTraceBlock, HAS_SHAPES ? 3 : kTailMinBlocks) tail_kernel(SceneView sc, FrameParams fp, RayQueue qin,
                                                                              WaveCounters *wc, RayQueue qspill,
                                                                              ShadowQueue sq, Counters *ctr,
                                                                              float4 *accum, WaveCounters *wc_sh) {
  const uint32_t lane = lane_id();
  const uint32_t count = wc[0].n_rays;
  const uint32_t S = (uint32_t)sc.shadow_samples;
  uint32_t c_shadow = 0, c_refl = 0, c_refr = 0, c_trunc = 0, c_culled = 0, c_tail = 0;
  while (true) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&wc[0].fetch_closest, 32u);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= count) break;
    const uint32_t i = base + lane;
    if (i < count) {
      RayState r = load_ray(qin, i);
      while (true) {
        float4 h = closest_hit<HAS_SHAPES, FMT>(sc, r.o, r.d);
        ++c_tail;
        Shaded s;
        shade_eval<HAS_SHAPES>(sc, fp, r, h, true, accum, s);
        c_trunc += (s.trunc_refl ? 1u : 0u) + (s.trunc_refr ? 1u : 0u);
        if (s.cull_sh) c_shadow += S, c_culled += S;
        if (s.emit_sh) {
          c_shadow += S;
          // lanes that reach this point together reserve their slots with ONE atomic (coalesced group)
          cooperative_groups::coalesced_group g = cooperative_groups::coalesced_threads();
          uint32_t sbase = 0;
          if (g.thread_rank() == 0) sbase = atomicAdd(&wc_sh->n_shadow, S * g.size());
          sbase = g.shfl(sbase, 0) + S * g.thread_rank();
          emit_shadow_rays<HAS_SHAPES, true, FMT>(sc, fp, r, s, sq, sbase, ctr, accum);
        }
        if (s.want_refl && s.want_refr) {
          ++c_refl, ++c_refr;
          uint32_t slot = atomicAdd(&wc[1].n_rays, 1u);
          if (slot < qspill.capacity)
            store_ray(qspill, slot, refract_ray(r, s));
          else
            ctr->overflow = 1u;
          r = reflect_ray(r, s);
        } else if (s.want_refl) {
          ++c_refl;
          r = reflect_ray(r, s);
        } else if (s.want_refr) {
          ++c_refr;
          r = refract_ray(r, s);
        } else {
          break;
        }
      }
    }
  }
  flush_stats(ctr, c_shadow, c_refl, c_refr, c_trunc, c_culled);
  c_tail = __reduce_add_sync(0xFFFFFFFFu, c_tail);
  if (lane == 0 && c_tail) atomicAdd(&ctr->rays_tail, (unsigned long long)c_tail);
}

// ---------------------------------------------------------------------------------------------
// K5 — resolve: pixel = tot_c / spp (src/scene.rs:94); optional RGB8 (src/image.rs:64-77)
// ---------------------------------------------------------------------------------------------
__global__ void resolve_kernel(const float4 *accum, uint32_t n, float spp, float *out_rgb) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = accum[i];
  out_rgb[3 * (size_t)i + 0] = a.x / spp;
  out_rgb[3 * (size_t)i + 1] = a.y / spp;
  out_rgb[3 * (size_t)i + 2] = a.z / spp;
}

__global__ void resolve_rgb8_kernel(const float4 *accum, uint32_t n, float spp, uint8_t *out_rgb8) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = accum[i];
  float c[3] = {a.x / spp, a.y / spp, a.z / spp};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float v = fminf(fmaxf(c[k] * 255.0f, 0.0f), 255.0f);  // inf(sup(color, 0), 255) then `as usize` truncation
    out_rgb8[3 * (size_t)i + k] = (uint8_t)(unsigned)v;
  }
}

// nrb_render with a pinned destination: the image's device->host copy is started when the frame enters its tail phase
// and overlaps it; afterwards this kernel stores the pixels the tail changed straight into the (mapped) host image.
// Neighbouring changed pixels leave a warp as contiguous segments, so the PCIe writes are mostly full packets.
__global__ void patch_host_image_kernel(const float4 *accum, const float *early_rgb, uint32_t n, float spp, float *host_rgb) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = accum[i];
  const float r = a.x / spp, g = a.y / spp, b = a.z / spp;
  const size_t o = 3 * (size_t)i;
  if (r != early_rgb[o] || g != early_rgb[o + 1] || b != early_rgb[o + 2]) host_rgb[o] = r, host_rgb[o + 1] = g, host_rgb[o + 2] = b;
}

// K5 + K6 fused for the multi-GPU path: this rank's packed tile accumulators are resolved (/ spp) straight into the
// ROW-MAJOR image `out_rgb`, which may live on another GPU (peer / IPC-mapped memory over NVLink): the finished
// pixels cross the link once, as 48-byte stores of four pixels, and nothing is gathered or un-tiled afterwards.
__global__ void resolve_tiles_to_image_kernel(const float4 *accum, FrameParams fp, float inv_spp, float *out_rgb) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;  // one thread = four pixels of one tile row
  if (t >= fp.n_local_tiles * (NRB_TILE * NRB_TILE / 4u)) return;
  const uint32_t lt = t / (NRB_TILE * NRB_TILE / 4u), r = t % (NRB_TILE * NRB_TILE / 4u);
  const uint32_t row = r / (NRB_TILE / 4u), x4 = (r % (NRB_TILE / 4u)) * 4u;
  const uint32_t tile = fp.tile_first + lt * fp.tile_stride;
  const uint32_t ty = fdiv(tile, fp.div_tiles_x), tx = tile - ty * fp.tiles_x;
  const uint32_t y = ty * NRB_TILE + row, x = tx * NRB_TILE + x4;
  if (y >= fp.height || x >= fp.width) return;
  const float4 *src = accum + (size_t)lt * (NRB_TILE * NRB_TILE) + row * NRB_TILE + x4;
  float *dst = out_rgb + 3u * ((size_t)y * fp.width + x);
  if (x + 3u < fp.width && (fp.width & 3u) == 0u) {
    const float4 a = src[0], b = src[1], c = src[2], d = src[3];
    float4 *d4 = reinterpret_cast<float4 *>(dst);  // 12 * (y * W + x) bytes: 16-byte aligned when W % 4 == 0 and x % 4 == 0
    d4[0] = make_float4(a.x * inv_spp, a.y * inv_spp, a.z * inv_spp, b.x * inv_spp);
    d4[1] = make_float4(b.y * inv_spp, b.z * inv_spp, c.x * inv_spp, c.y * inv_spp);
    d4[2] = make_float4(c.z * inv_spp, d.x * inv_spp, d.y * inv_spp, d.z * inv_spp);
  } else {
    for (uint32_t k = 0; k < 4u && x + k < fp.width; ++k) {
      const float4 a = src[k];
      dst[3 * k + 0] = a.x * inv_spp, dst[3 * k + 1] = a.y * inv_spp, dst[3 * k + 2] = a.z * inv_spp;
    }
  }
}

// End-to-end form of the exchange (nrb_render_tiles_to_host): when tiles_x is a multiple of the rank count, rank r owns
// whole tile COLUMNS r, r + N, ...; in the row-major image its pixels are 16-pixel (192-byte) segments at pitch
// N * 192 bytes.  This kernel resolves the rank's tiles into a staging buffer holding exactly those segments back to
// back — segment j = y * columns_per_rank + k — so ONE strided 2-D DMA drops them into the shared host image.
__global__ void resolve_tiles_to_segments_kernel(const float4 *accum, FrameParams fp, float inv_spp, uint32_t cols_per_rank, float *stage) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;  // one thread = four pixels of one tile row
  if (t >= fp.n_local_tiles * (NRB_TILE * NRB_TILE / 4u)) return;
  const uint32_t lt = t / (NRB_TILE * NRB_TILE / 4u), r = t % (NRB_TILE * NRB_TILE / 4u);
  const uint32_t row = r / (NRB_TILE / 4u), x4 = (r % (NRB_TILE / 4u)) * 4u;
  const uint32_t tile = fp.tile_first + lt * fp.tile_stride;
  const uint32_t ty = fdiv(tile, fp.div_tiles_x), tx = tile - ty * fp.tiles_x;
  const uint32_t y = ty * NRB_TILE + row;
  if (y >= fp.height) return;
  const uint32_t k = fdiv(tx - fp.tile_first, fp.div_tile_stride);  // tile_first < tile_stride: the rank's k-th column
  const float4 *src = accum + (size_t)lt * (NRB_TILE * NRB_TILE) + row * NRB_TILE + x4;
  float4 *dst = reinterpret_cast<float4 *>(stage + ((size_t)(y * cols_per_rank + k) * NRB_TILE + x4) * 3u);
  const float4 a = src[0], b = src[1], c = src[2], d = src[3];
  dst[0] = make_float4(a.x * inv_spp, a.y * inv_spp, a.z * inv_spp, b.x * inv_spp);
  dst[1] = make_float4(b.y * inv_spp, b.z * inv_spp, c.x * inv_spp, c.y * inv_spp);
  dst[2] = make_float4(c.z * inv_spp, d.x * inv_spp, d.y * inv_spp, d.z * inv_spp);
}

// RGB8 forms of the two fused exchanges ("resolve + RGB8 quantise fused into the gather"): the finished pixels
// leave the GPU as Image::to_png would encode them — clamp(c * 255, 0, 255) truncated to u8 (src/image.rs:64-77) — so the
// exchange moves 3 bytes per pixel instead of 12, over NVLink (tiles_to_image) or PCIe (tiles_to_segments + 2-D DMA).
NRB_DI uint32_t quant8(float c) { return (uint32_t)fminf(fmaxf(c * 255.0f, 0.0f), 255.0f); }
NRB_DI void quant_px4(const float4 *src, float inv_spp, uint32_t w[3]) {
  uint32_t b[12];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 a = src[k];
    b[3 * k] = quant8(a.x * inv_spp), b[3 * k + 1] = quant8(a.y * inv_spp), b[3 * k + 2] = quant8(a.z * inv_spp);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) w[k] = b[4 * k] | (b[4 * k + 1] << 8) | (b[4 * k + 2] << 16) | (b[4 * k + 3] << 24);
}

__global__ void resolve_tiles_to_image_rgb8_kernel(const float4 *accum, FrameParams fp, float inv_spp, uint8_t *out_rgb8) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;  // one thread = four pixels of one tile row = 12 bytes
  if (t >= fp.n_local_tiles * (NRB_TILE * NRB_TILE / 4u)) return;
  const uint32_t lt = t / (NRB_TILE * NRB_TILE / 4u), r = t % (NRB_TILE * NRB_TILE / 4u);
  const uint32_t row = r / (NRB_TILE / 4u), x4 = (r % (NRB_TILE / 4u)) * 4u;
  const uint32_t tile = fp.tile_first + lt * fp.tile_stride;
  const uint32_t ty = fdiv(tile, fp.div_tiles_x), tx = tile - ty * fp.tiles_x;
  const uint32_t y = ty * NRB_TILE + row, x = tx * NRB_TILE + x4;
  if (y >= fp.height || x >= fp.width) return;
  const float4 *src = accum + (size_t)lt * (NRB_TILE * NRB_TILE) + row * NRB_TILE + x4;
  uint8_t *dst = out_rgb8 + 3u * ((size_t)y * fp.width + x);
  if (x + 3u < fp.width && (fp.width & 3u) == 0u) {
    uint32_t w[3];
    quant_px4(src, inv_spp, w);
    uint32_t *d4 = reinterpret_cast<uint32_t *>(dst);  // 3 * (y * W + x) bytes: 4-byte aligned when W % 4 == 0 and x % 4 == 0
    d4[0] = w[0], d4[1] = w[1], d4[2] = w[2];
  } else {
    for (uint32_t k = 0; k < 4u && x + k < fp.width; ++k) {
      const float4 a = src[k];
      dst[3 * k + 0] = (uint8_t)quant8(a.x * inv_spp), dst[3 * k + 1] = (uint8_t)quant8(a.y * inv_spp), dst[3 * k + 2] = (uint8_t)quant8(a.z * inv_spp);
    }
  }
}

__global__ void resolve_tiles_to_segments_rgb8_kernel(const float4 *accum, FrameParams fp, float inv_spp, uint32_t cols_per_rank, uint8_t *stage) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= fp.n_local_tiles * (NRB_TILE * NRB_TILE / 4u)) return;
  const uint32_t lt = t / (NRB_TILE * NRB_TILE / 4u), r = t % (NRB_TILE * NRB_TILE / 4u);
  const uint32_t row = r / (NRB_TILE / 4u), x4 = (r % (NRB_TILE / 4u)) * 4u;
  const uint32_t tile = fp.tile_first + lt * fp.tile_stride;
  const uint32_t ty = fdiv(tile, fp.div_tiles_x), tx = tile - ty * fp.tiles_x;
  const uint32_t y = ty * NRB_TILE + row;
  if (y >= fp.height) return;
  const uint32_t k = fdiv(tx - fp.tile_first, fp.div_tile_stride);
  const float4 *src = accum + (size_t)lt * (NRB_TILE * NRB_TILE) + row * NRB_TILE + x4;
  uint32_t w[3];
  quant_px4(src, inv_spp, w);
  uint32_t *dst = reinterpret_cast<uint32_t *>(stage + ((size_t)(y * cols_per_rank + k) * NRB_TILE + x4) * 3u);  // 48-byte segments
  dst[0] = w[0], dst[1] = w[1], dst[2] = w[2];
}

// K6 — packed tiles of n_ranks ranks (rank r owns tiles r, r+n_ranks, ...) -> row-major image
__global__ void untile_kernel(const float *gathered, uint32_t n_ranks, uint32_t tiles_per_rank, uint32_t width,
                              uint32_t height, uint32_t tiles_x, float *out_rgb) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= width * height) return;
  uint32_t y = i / width, x = i - y * width;
  uint32_t tile = (y / NRB_TILE) * tiles_x + (x / NRB_TILE);
  uint32_t rank = tile % n_ranks, lt = tile / n_ranks;
  size_t src = ((size_t)rank * tiles_per_rank + lt) * (NRB_TILE * NRB_TILE) + (y % NRB_TILE) * NRB_TILE + (x % NRB_TILE);
  out_rgb[3 * (size_t)i + 0] = gathered[3 * src + 0];
  out_rgb[3 * (size_t)i + 1] = gathered[3 * src + 1];
  out_rgb[3 * (size_t)i + 2] = gathered[3 * src + 2];
}

// ---------------------------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------------------------
void launch_trace(const SceneView &sc, bool has_shapes, const FrameParams &fp, bool primary, RayQueue q, float4 *hits,
                  WaveCounters *wc_closest, uint32_t slot_lo, uint32_t n_slots, ShadowQueue sq, float4 *accum,
                  WaveCounters *wc_shadow, TraceOpts opts, int grid, cudaStream_t st) {
  if (!wc_closest && !wc_shadow) return;
#define NRB_LAUNCH_TRACE(HS, PR, FM)                                                                                          \
  trace_kernel<HS, PR, FM><<<grid, kTraceBlock, 0, st>>>(sc, fp, q, hits, wc_closest, slot_lo, n_slots, sq, accum, wc_shadow, \
                                                         opts)
  // formats 3 / 4 exist for mesh-only scenes (api.cu: upload_scene never picks them when the scene has analytic shapes)
  const int sel = (has_shapes ? 16 : 0) | (primary ? 8 : 0) | sc.node_format;
  switch (sel) {
    case 0: NRB_LAUNCH_TRACE(false, false, 0); break;
    case 2: NRB_LAUNCH_TRACE(false, false, 2); break;
    case 3: NRB_LAUNCH_TRACE(false, false, 3); break;
    case 4: NRB_LAUNCH_TRACE(false, false, 4); break;
    case 8: NRB_LAUNCH_TRACE(false, true, 0); break;
    case 10: NRB_LAUNCH_TRACE(false, true, 2); break;
    case 11: NRB_LAUNCH_TRACE(false, true, 3); break;
    case 12: NRB_LAUNCH_TRACE(false, true, 4); break;
    case 16: NRB_LAUNCH_TRACE(true, false, 0); break;
    case 18: NRB_LAUNCH_TRACE(true, false, 2); break;
    case 24: NRB_LAUNCH_TRACE(true, true, 0); break;
    case 26: NRB_LAUNCH_TRACE(true, true, 2); break;
    default: break;  // unreachable: upload_scene validates the format
  }
#undef NRB_LAUNCH_TRACE
}

void launch_trace_general(const SceneView &sc, const FrameParams &fp, bool primary, RayQueue q, float4 *hits,
                          WaveCounters *wc_closest, uint32_t slot_lo, uint32_t n_slots, ShadowQueue sq, float4 *accum,
                          WaveCounters *wc_shadow, int grid, cudaStream_t st) {
  if (!wc_closest && !wc_shadow) return;
  if (primary)
    trace_general_kernel<true><<<grid, kTraceBlock, 0, st>>>(sc, fp, q, hits, wc_closest, slot_lo, n_slots, sq, accum, wc_shadow);
  else
    trace_general_kernel<false><<<grid, kTraceBlock, 0, st>>>(sc, fp, q, hits, wc_closest, slot_lo, n_slots, sq, accum, wc_shadow);
}

void launch_shade(const SceneView &sc, bool has_shapes, const FrameParams &fp, bool primary, RayQueue qin,
                  const float4 *hits, WaveCounters *wc, uint32_t slot_lo, uint32_t n_slots, uint32_t lo, uint32_t hi,
                  RayQueue qout, ShadowQueue sq, Counters *ctr, float4 *accum, int grid, cudaStream_t st) {
  if (hi <= lo) return;
  if (has_shapes) {
    if (primary)
      shade_kernel<true, true><<<grid, kShadeBlock, 0, st>>>(sc, fp, qin, hits, wc, slot_lo, n_slots, lo, hi, qout, sq, ctr, accum);
    else
      shade_kernel<true, false><<<grid, kShadeBlock, 0, st>>>(sc, fp, qin, hits, wc, slot_lo, n_slots, lo, hi, qout, sq, ctr, accum);
  } else {
    if (primary)
      shade_kernel<false, true><<<grid, kShadeBlock, 0, st>>>(sc, fp, qin, hits, wc, slot_lo, n_slots, lo, hi, qout, sq, ctr, accum);
    else
      shade_kernel<false, false><<<grid, kShadeBlock, 0, st>>>(sc, fp, qin, hits, wc, slot_lo, n_slots, lo, hi, qout, sq, ctr, accum);
  }
}

void launch_tail(const SceneView &sc, bool has_shapes, const FrameParams &fp, RayQueue qin, WaveCounters *wc,
                 RayQueue qspill, ShadowQueue sq, Counters *ctr, float4 *accum, WaveCounters *wc_sh, int grid,
                 cudaStream_t st) {
  const bool f2 = sc.node_format == 2;
  if (has_shapes) {
    if (f2) tail_kernel<true, 2><<<grid, kTraceBlock, 0, st>>>(sc, fp, qin, wc, qspill, sq, ctr, accum, wc_sh);
    else tail_kernel<true, 0><<<grid, kTraceBlock, 0, st>>>(sc, fp, qin, wc, qspill, sq, ctr, accum, wc_sh);
  } else if (sc.node_format >= 3) {  // the tail follows single paths: no warp to speculate for, one form for both
    tail_kernel<false, 3><<<grid, kTraceBlock, 0, st>>>(sc, fp, qin, wc, qspill, sq, ctr, accum, wc_sh);
  } else {
    if (f2) tail_kernel<false, 2><<<grid, kTraceBlock, 0, st>>>(sc, fp, qin, wc, qspill, sq, ctr, accum, wc_sh);
    else tail_kernel<false, 0><<<grid, kTraceBlock, 0, st>>>(sc, fp, qin, wc, qspill, sq, ctr, accum, wc_sh);
  }
}

int shade_blocks_per_sm(bool has_shapes) {
  int nb = 0, nb2 = 0;
  if (has_shapes) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, shade_kernel<true, false>, kShadeBlock, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, shade_kernel<true, true>, kShadeBlock, 0);
  } else {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, shade_kernel<false, false>, kShadeBlock, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, shade_kernel<false, true>, kShadeBlock, 0);
  }
  nb = nb < nb2 ? nb : nb2;
  return nb > 0 ? nb : 1;
}

void launch_resolve(const float4 *accum, uint32_t n, uint32_t spp, float *out_rgb, cudaStream_t st) {
  if (!n) return;
  resolve_kernel<<<(n + 255) / 256, 256, 0, st>>>(accum, n, (float)spp, out_rgb);
}

void launch_resolve_rgb8(const float4 *accum, uint32_t n, uint32_t spp, uint8_t *out, cudaStream_t st) {
  if (!n) return;
  resolve_rgb8_kernel<<<(n + 255) / 256, 256, 0, st>>>(accum, n, (float)spp, out);
}

void launch_patch_host_image(const float4 *accum, const float *early_rgb, uint32_t n, uint32_t spp, float *host_rgb, cudaStream_t st) {
  if (!n) return;
  patch_host_image_kernel<<<(n + 255) / 256, 256, 0, st>>>(accum, early_rgb, n, (float)spp, host_rgb);
}

void launch_resolve_tiles_to_image(const float4 *accum, const FrameParams &fp, float *out_rgb, cudaStream_t st) {
  const uint32_t n = fp.n_local_tiles * (NRB_TILE * NRB_TILE / 4u);
  if (!n) return;
  resolve_tiles_to_image_kernel<<<(n + 255) / 256, 256, 0, st>>>(accum, fp, 1.0f / (float)fp.spp, out_rgb);
}

void launch_resolve_tiles_to_segments(const float4 *accum, const FrameParams &fp, uint32_t cols_per_rank, float *stage, cudaStream_t st) {
  const uint32_t n = fp.n_local_tiles * (NRB_TILE * NRB_TILE / 4u);
  if (!n) return;
  resolve_tiles_to_segments_kernel<<<(n + 255) / 256, 256, 0, st>>>(accum, fp, 1.0f / (float)fp.spp, cols_per_rank, stage);
}

void launch_resolve_tiles_to_image_rgb8(const float4 *accum, const FrameParams &fp, uint8_t *out_rgb8, cudaStream_t st) {
  const uint32_t n = fp.n_local_tiles * (NRB_TILE * NRB_TILE / 4u);
  if (!n) return;
  resolve_tiles_to_image_rgb8_kernel<<<(n + 255) / 256, 256, 0, st>>>(accum, fp, 1.0f / (float)fp.spp, out_rgb8);
}

void launch_resolve_tiles_to_segments_rgb8(const float4 *accum, const FrameParams &fp, uint32_t cols_per_rank, uint8_t *stage, cudaStream_t st) {
  const uint32_t n = fp.n_local_tiles * (NRB_TILE * NRB_TILE / 4u);
  if (!n) return;
  resolve_tiles_to_segments_rgb8_kernel<<<(n + 255) / 256, 256, 0, st>>>(accum, fp, 1.0f / (float)fp.spp, cols_per_rank, stage);
}

void launch_untile(const float *gathered, uint32_t n_ranks, uint32_t tiles_per_rank, uint32_t width, uint32_t height,
                   float *out_rgb, cudaStream_t st) {
  uint32_t n = width * height;
  uint32_t tiles_x = (width + NRB_TILE - 1) / NRB_TILE;
  untile_kernel<<<(n + 255) / 256, 256, 0, st>>>(gathered, n_ranks, tiles_per_rank, width, height, tiles_x, out_rgb);
}

void debug_visit_counters(unsigned long long out[4], bool reset) {
#ifdef NRB_COUNT_VISITS
  cudaMemcpyFromSymbol(out, g_dbg, sizeof(unsigned long long) * 4);
  if (reset) {
    unsigned long long z[4] = {0, 0, 0, 0};
    cudaMemcpyToSymbol(g_dbg, z, sizeof(z));
  }
#else
  out[0] = out[1] = out[2] = out[3] = 0;
#endif
}

int trace_blocks_per_sm(bool has_shapes) {
  int nb = 0, nb2 = 0;
  if (has_shapes) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_kernel<true, false, 0>, kTraceBlock, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, trace_kernel<true, true, 0>, kTraceBlock, 0);
  } else {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_kernel<false, false, 0>, kTraceBlock, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, trace_kernel<false, true, 0>, kTraceBlock, 0);
  }
  nb = nb < nb2 ? nb : nb2;
  return nb > 0 ? nb : 1;
}

}  // namespace nrb
