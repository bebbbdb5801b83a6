/*
 * nrays_oracle.cpp — CPU oracle for the nrays render hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * What it is: a restatement, function by function, of
 *     src/scene.rs            render / Scene::new / trace / trace_reflection / trace_refraction /
 *                             intersects_ray / ClosestRayTOICostFn / TransparentShadowsRayTOICostFn
 *     src/scene_node.rs       SceneNode::cast
 *     src/phong_material.rs   PhongMaterial::{ambiant, compute}
 *     src/normal_material.rs, src/uv_material.rs
 *     src/light.rs            Light::sample (the live cfg(not(feature="3d")) variant, :56-63)
 *     src/texture2d.rs        Texture2d::{at, sample}
 *     src/ray_with_energy.rs
 * and of the third-party arithmetic at their call sites, which is NOT under /root/reference:
 *     ncollide3d 0.16 (Cargo.toml:12, semver range, no Cargo.lock): BVT::new_balanced,
 *         BVT::best_first_search, AABB::toi_with_ray, RayCast for Ball / Cuboid / Plane /
 *         TriMesh (+ triangle_ray_intersection), bounding volumes;
 *     nalgebra 0.15 (Cargo.toml:13): Matrix4 * Point4, Point3::from_homogeneous, normalize;
 *     rand 0.5 (Cargo.toml:14): rand::random (OS seeded, unreproducible).
 * Each function cites the reference file:line or the SURVEY.md Appendix B item it follows.
 *
 * PARITY UNPINNED BY THE REFERENCE.  The reference ships no test, golden image or known-answer
 * vector, cannot be compiled in this environment (no rustc/cargo, no vendored
 * crates, no network) and ncollide3d's source is absent, so nothing from the reference pins this
 * oracle.  Its own pins: closed-form known answers per primitive and the Random123 Philox
 * known-answer vectors (tests/test_oracle_*.py), f64-vs-f32 self-consistency, and — since round 2 —
 * an INDEPENDENT brute-force second implementation (tests/bruteforce.py: no BVT, no shared code,
 * different intersection formulas, reads the host Scene objects instead of the flattened tables)
 * that must agree with nro_cast / nro_intersects_ray / nro_trace / nro_render on randomised
 * scenes (tests/test_bruteforce_pin.py).  That pins the implementation of SURVEY Appendix A+B;
 * it cannot pin Appendix B's reading of ncollide3d itself.
 *
 * Deliberate, documented deviations from the reference:
 *   D1  RNG: rand::random is replaced by counter-based Philox4x32-10 keyed by (seed, stream ^ (depth/32)*phi) with
 *       counter (pixel, sample, path, light<<16|k); uniforms carry 24 bits ((x>>8)*2^-24) so the
 *       f32 device draws the identical value.  With window_width = 0 and light radius = 0 the
 *       render is RNG-free and this deviation vanishes.
 *   D2  Cylinder / Cone / Capsule ray casts are closed-form; ncollide3d uses iterative GJK
 *       (tolerance-based).  Same hit point / outward normal up to GJK tolerance.
 *   D3  Recursion is capped at NrbCamera.max_depth (reference: unbounded, overflows its stack on
 *       facing mirrors).  Truncated spawns are counted in NrbStats.paths_truncated.
 *   D4  Bilinear texture fetch clamps the flat texel index to the pool instead of panicking when
 *       hi = lo+1 runs past the last row; identical wherever the reference does not panic.
 *   D5  A ray/plane toi of +inf (ray parallel to the plane, origin behind it) is a miss: in the
 *       reference it fails `cost < best_cost` against the f64::MAX initial best.
 *
 * Templated on the scalar type: T = double is the faithful mode (reference Scalar = f64,
 * src/lib.rs:33); T = float is the device twin used to separate precision drift from logic bugs.
 */
#include "nrays_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <queue>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;

// ------------------------------------------------------------------------------------------
// small vector type
// ------------------------------------------------------------------------------------------
template <class T>
struct V3 {
  T x, y, z;
  V3() : x(0), y(0), z(0) {}
  V3(T a, T b, T c) : x(a), y(b), z(c) {}
  T &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  const T &operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  V3 operator+(const V3 &o) const { return V3(x + o.x, y + o.y, z + o.z); }
  V3 operator-(const V3 &o) const { return V3(x - o.x, y - o.y, z - o.z); }
  V3 operator-() const { return V3(-x, -y, -z); }
  V3 operator*(T s) const { return V3(x * s, y * s, z * s); }
  V3 operator/(T s) const { return V3(x / s, y / s, z / s); }
};
template <class T>
T dot(const V3<T> &a, const V3<T> &b) {
  return a.x * b.x + a.y * b.y + a.z * b.z;
}
template <class T>
V3<T> cross(const V3<T> &a, const V3<T> &b) {
  return V3<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <class T>
T norm(const V3<T> &a) {
  return std::sqrt(dot(a, a));
}
template <class T>
V3<T> normalize(const V3<T> &a) {
  return a / norm(a);
}
template <class T>
V3<T> cmul(const V3<T> &a, const V3<T> &b) {
  return V3<T>(a.x * b.x, a.y * b.y, a.z * b.z);
}

using C3 = V3<float>;  // colours are always f32 (src/scene.rs uses Vector3<f32>)
struct C4 {
  float x, y, z, w;
};

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11; Random123 reference constants) — deviation D1
// ------------------------------------------------------------------------------------------
void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0, c1 = n1, c2 = n2, c3 = n3;
    k0 += W0, k1 += W1;
  }
  out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

const uint32_t STREAM_PRIMARY = 0x50524D59u;  // 'PRMY'
const uint32_t STREAM_LIGHT = 0x4C474854u;    // 'LGHT'

// 24-bit uniform in [0,1): exactly representable in f32 so oracle (f64) and device (f32) agree.
inline double u24(uint32_t x) { return (double)(x >> 8) * (1.0 / 16777216.0); }

// ------------------------------------------------------------------------------------------
// Ray, AABB
// ------------------------------------------------------------------------------------------
template <class T>
struct Ray {
  V3<T> o, d;
};

template <class T>
struct Aabb {
  V3<T> mins, maxs;
  void merge(const Aabb &b) {
    for (int i = 0; i < 3; ++i) {
      mins[i] = std::min(mins[i], b.mins[i]);
      maxs[i] = std::max(maxs[i], b.maxs[i]);
    }
  }
  V3<T> center() const { return (mins + maxs) * T(0.5); }
};

// AABB::toi_with_ray(identity, ray, solid) — SURVEY B.3 (ncollide3d ray_aabb, toi only).
// Called with solid = true at src/scene.rs:276,309.
template <class T>
bool aabb_toi(const Aabb<T> &bb, const Ray<T> &r, bool solid, T &toi) {
  T tmin = 0, tmax = std::numeric_limits<T>::max();
  for (int i = 0; i < 3; ++i) {
    if (r.d[i] == T(0)) {
      if (r.o[i] < bb.mins[i] || r.o[i] > bb.maxs[i]) return false;
    } else {
      T inv = T(1) / r.d[i];
      T t0 = (bb.mins[i] - r.o[i]) * inv;
      T t1 = (bb.maxs[i] - r.o[i]) * inv;
      if (t0 > t1) std::swap(t0, t1);
      tmin = std::max(tmin, t0);
      tmax = std::min(tmax, t1);
      if (tmin > tmax) return false;
    }
  }
  toi = (tmin == T(0) && !solid) ? tmax : tmin;
  return true;
}

// ------------------------------------------------------------------------------------------
// BVT — SURVEY B.1 (new_balanced: median partitioner) and B.2 (best_first_search)
// ------------------------------------------------------------------------------------------
template <class T>
struct Bvt {
  struct Node {
    Aabb<T> bv;
    int left = -1, right = -1;  // internal
    int leaf = -1;              // leaf payload (index supplied by the caller)
  };
  std::vector<Node> nodes;
  int root = -1;

  // BVT::new_balanced(leaves) — called at src/scene.rs:126 and inside TriMesh::new (loader3d.rs:695).
  void build(const std::vector<Aabb<T>> &bvs) {
    nodes.clear();
    root = -1;
    if (bvs.empty()) return;
    std::vector<int> ids(bvs.size());
    for (size_t i = 0; i < ids.size(); ++i) ids[i] = (int)i;
    nodes.reserve(2 * bvs.size());
    root = build_rec(0, ids, bvs);
  }

  int build_rec(int depth, std::vector<int> &leaves, const std::vector<Aabb<T>> &bvs) {
    if (leaves.size() == 1) {
      Node n;
      n.bv = bvs[leaves[0]];
      n.leaf = leaves[0];
      nodes.push_back(n);
      return (int)nodes.size() - 1;
    }
    // median_partitioner: split axis = depth % 3, median of the AABB centres (mean of the two
    // middle values when the count is even), ties alternate sides, empty side gets one leaf.
    int axis = depth % 3;
    std::vector<T> med(leaves.size());
    for (size_t i = 0; i < leaves.size(); ++i) med[i] = bvs[leaves[i]].center()[axis];
    std::sort(med.begin(), med.end());
    size_t n = med.size();
    T median = (n % 2 == 0) ? (med[n / 2 - 1] + med[n / 2]) / T(2) : med[n / 2];
    std::vector<int> left, right;
    bool insert_left = false;
    for (int id : leaves) {
      T pos = bvs[id].center()[axis];
      if (pos < median || (pos == median && insert_left)) {
        left.push_back(id);
        insert_left = false;
      } else {
        right.push_back(id);
        insert_left = true;
      }
    }
    if (left.empty()) {
      left.push_back(right.back());
      right.pop_back();
    } else if (right.empty()) {
      right.push_back(left.back());
      left.pop_back();
    }
    std::vector<int>().swap(leaves);
    int l = build_rec(depth + 1, left, bvs);
    int r = build_rec(depth + 1, right, bvs);
    Node nd;
    nd.bv = nodes[l].bv;
    nd.bv.merge(nodes[r].bv);
    nd.left = l;
    nd.right = r;
    nodes.push_back(nd);
    return (int)nodes.size() - 1;
  }

  // BVT::best_first_search — SURVEY B.2.  `bv_cost(aabb, &c)` returns false for None;
  // `leaf_cost(leaf, &c, &data)` returns false for None and otherwise fills the UserData.
  // A BinaryHeap is allocated per call exactly like upstream (kept on purpose: this oracle
  // doubles as the CPU baseline).  Returns the best leaf or -1; `best_data` gets its UserData.
  template <class Data, class BvCost, class LeafCost>
  int best_first_search(BvCost &&bv_cost, LeafCost &&leaf_cost, Data &best_data) const {
    if (root < 0) return -1;
    typedef std::pair<T, int> Item;
    auto cmp = [](const Item &a, const Item &b) { return a.first > b.first; };
    std::priority_queue<Item, std::vector<Item>, decltype(cmp)> heap(cmp);
    T best_cost = std::numeric_limits<T>::max();
    int best = -1;
    T c;
    Data data;
    if (bv_cost(nodes[root].bv, c)) heap.push(Item(c, root));
    while (!heap.empty()) {
      Item it = heap.top();
      heap.pop();
      if (it.first >= best_cost) break;
      const Node &n = nodes[it.second];
      if (n.leaf < 0) {
        if (bv_cost(nodes[n.left].bv, c) && c < best_cost) heap.push(Item(c, n.left));
        if (bv_cost(nodes[n.right].bv, c) && c < best_cost) heap.push(Item(c, n.right));
      } else {
        if (leaf_cost(n.leaf, c, data) && c < best_cost) {
          best_cost = c;
          best = n.leaf;
          best_data = data;
        }
      }
    }
    return best;
  }
};

// ------------------------------------------------------------------------------------------
// RayIntersection
// ------------------------------------------------------------------------------------------
template <class T>
struct Inter {
  T toi = 0;
  V3<T> normal;
  bool has_uv = false;
  T u = 0, v = 0;
};

// ---- Ball — SURVEY B.4 (ncollide3d ball_toi_with_ray + ball_uv) ------------------------------
template <class T>
bool cast_ball(const V3<T> &center, T radius, const Ray<T> &ray, bool solid, Inter<T> &out) {
  V3<T> dc = ray.o - center;
  T a = dot(ray.d, ray.d);
  T b = dot(dc, ray.d);
  T c = dot(dc, dc) - radius * radius;
  if (c > T(0) && b > T(0)) return false;
  T delta = b * b - a * c;
  if (delta < T(0)) return false;
  T t = (-b - std::sqrt(delta)) / a;
  bool inside = false;
  if (t <= T(0)) {
    inside = true;
    t = solid ? T(0) : (-b + std::sqrt(delta)) / a;
  }
  V3<T> n = normalize(ray.o + ray.d * t - center);
  const T pi = T(3.14159265358979323846);
  out.toi = t;
  out.has_uv = true;
  out.u = T(0.5) + std::atan2(n.z, n.x) / (pi + pi);
  out.v = T(0.5) - std::asin(n.y) / pi;
  out.normal = inside ? -n : n;
  return true;
}

// ---- Cuboid — SURVEY B.5 (ncollide3d ray_aabb with side ids, local frame) ---------------------
template <class T>
bool cast_cuboid(const V3<T> &he, const Ray<T> &r, bool solid, Inter<T> &out) {
  V3<T> mins = -he, maxs = he;
  T tmax = std::numeric_limits<T>::max();
  T tmin = -tmax;
  int near_side = 0, far_side = 0;
  bool near_diag = false, far_diag = false;
  for (int i = 0; i < 3; ++i) {
    if (r.d[i] == T(0)) {
      if (r.o[i] < mins[i] || r.o[i] > maxs[i]) return false;
    } else {
      T denom = T(1) / r.d[i];
      T tn = (mins[i] - r.o[i]) * denom;
      T tf = (maxs[i] - r.o[i]) * denom;
      bool flip = false;
      if (tn > tf) {
        flip = true;
        std::swap(tn, tf);
      }
      if (tn > tmin) {
        tmin = tn;
        near_side = flip ? -(i + 1) : (i + 1);
        near_diag = false;
      } else if (tn == tmin) {
        near_diag = true;
      }
      if (tf < tmax) {
        tmax = tf;
        far_side = flip ? (i + 1) : -(i + 1);
        far_diag = false;
      } else if (tf == tmax) {
        far_diag = true;
      }
      if (tmax < T(0) || tmin > tmax) return false;
    }
  }
  int side;
  V3<T> n;
  T toi;
  if (tmin < T(0)) {  // origin inside
    if (solid) {
      toi = 0;
      side = far_side;
    } else {
      toi = tmax;
      side = far_side;
      // exit face, normal pointing inward (against the ray) — SURVEY B.5
      if (far_diag)
        n = -normalize(r.d);
      else if (far_side < 0)
        n[-far_side - 1] = T(-1);
      else
        n[far_side - 1] = T(1);
    }
  } else {
    toi = tmin;
    side = near_side;
    if (near_diag)
      n = -normalize(r.d);
    else if (near_side < 0)
      n[-near_side - 1] = T(1);
    else
      n[near_side - 1] = T(-1);
  }
  out.toi = toi;
  out.normal = n;
  // uv by face id — SURVEY B.5
  V3<T> pt = r.o + r.d * toi;
  V3<T> dpt = pt - mins;
  V3<T> scale = maxs - mins;
  int id = std::abs(side) - 1;
  out.has_uv = true;
  if (id == 0) {
    out.u = dpt.y / scale.y, out.v = dpt.z / scale.z;
  } else if (id == 1) {
    out.u = dpt.z / scale.z, out.v = dpt.x / scale.x;
  } else {
    out.u = dpt.x / scale.x, out.v = dpt.y / scale.y;
  }
  return true;
}

// ---- Plane — SURVEY B.7 (ncollide3d Plane RayCast, local frame) --------------------------------
template <class T>
bool cast_plane(const V3<T> &n, const Ray<T> &r, bool solid, Inter<T> &out) {
  T dn = dot(n, -r.o);
  if (solid && dn > T(0)) {
    out.toi = 0;
    out.normal = V3<T>();
    return true;
  }
  T t = dn / dot(n, r.d);
  if (!(t >= T(0))) return false;
  if (!(t < std::numeric_limits<T>::max())) return false;  // D5
  out.toi = t;
  out.normal = dn > T(0) ? -n : n;
  return true;
}

// ---- Convex solids of revolution, closed form — deviation D2 ----------------------
// Each helper intersects the line with one convex set and narrows [tin, tout], recording the
// outward normal of the surface that produced each bound.
template <class T>
struct Span {
  T tin = -std::numeric_limits<T>::infinity(), tout = std::numeric_limits<T>::infinity();
  V3<T> nin, nout;
  bool empty = false;
  void clip_in(T t, const V3<T> &n) {
    if (t > tin) tin = t, nin = n;
  }
  void clip_out(T t, const V3<T> &n) {
    if (t < tout) tout = t, nout = n;
  }
};

// y in [-h, h]
template <class T>
void clip_slab_y(Span<T> &s, const Ray<T> &r, T lo, T hi) {
  if (r.d.y == T(0)) {
    if (r.o.y < lo || r.o.y > hi) s.empty = true;
    return;
  }
  T t0 = (lo - r.o.y) / r.d.y, t1 = (hi - r.o.y) / r.d.y;
  if (r.d.y > T(0)) {
    s.clip_in(t0, V3<T>(0, -1, 0));
    s.clip_out(t1, V3<T>(0, 1, 0));
  } else {
    s.clip_in(t1, V3<T>(0, 1, 0));
    s.clip_out(t0, V3<T>(0, -1, 0));
  }
}

// x^2 + z^2 <= r^2
template <class T>
void clip_inf_cylinder(Span<T> &s, const Ray<T> &r, T rad) {
  T A = r.d.x * r.d.x + r.d.z * r.d.z;
  T B = r.o.x * r.d.x + r.o.z * r.d.z;
  T C = r.o.x * r.o.x + r.o.z * r.o.z - rad * rad;
  if (A == T(0)) {
    if (C > T(0)) s.empty = true;
    return;
  }
  T disc = B * B - A * C;
  if (disc < T(0)) {
    s.empty = true;
    return;
  }
  T sq = std::sqrt(disc);
  T t0 = (-B - sq) / A, t1 = (-B + sq) / A;
  V3<T> p0 = r.o + r.d * t0, p1 = r.o + r.d * t1;
  s.clip_in(t0, normalize(V3<T>(p0.x, 0, p0.z)));
  s.clip_out(t1, normalize(V3<T>(p1.x, 0, p1.z)));
}

// ball of radius rad centred at (0, cy, 0): returns [t0, t1] or empty
template <class T>
bool line_ball(const Ray<T> &r, T cy, T rad, T &t0, T &t1) {
  V3<T> dc = r.o - V3<T>(0, cy, 0);
  T a = dot(r.d, r.d), b = dot(dc, r.d), c = dot(dc, dc) - rad * rad;
  T disc = b * b - a * c;
  if (disc < T(0)) return false;
  T sq = std::sqrt(disc);
  t0 = (-b - sq) / a;
  t1 = (-b + sq) / a;
  return true;
}

template <class T>
bool finish_span(const Span<T> &s, bool solid, Inter<T> &out) {
  if (s.empty || s.tin > s.tout || s.tout < T(0)) return false;
  if (s.tin > T(0)) {
    out.toi = s.tin;
    out.normal = s.nin;
    return true;
  }
  // origin inside: solid -> toi 0; else exit point with the OUTWARD normal
  if (solid) {
    out.toi = 0;
    out.normal = V3<T>();
    return true;
  }
  if (!(s.tout < std::numeric_limits<T>::max())) return false;
  out.toi = s.tout;
  out.normal = s.nout;
  return true;
}

template <class T>
bool cast_cylinder(T hh, T rad, const Ray<T> &r, bool solid, Inter<T> &out) {
  Span<T> s;
  clip_slab_y(s, r, -hh, hh);
  if (!s.empty) clip_inf_cylinder(s, r, rad);
  return finish_span(s, solid, out);
}

// Cone: apex (0, +hh, 0), base disc y = -hh of radius rad.  With s = hh - y the solid is
// x^2 + z^2 <= (k s)^2, 0 <= s <= 2 hh, k = rad / (2 hh).
template <class T>
bool cast_cone(T hh, T rad, const Ray<T> &r, bool solid, Inter<T> &out) {
  const T inf = std::numeric_limits<T>::infinity();
  T k = rad / (T(2) * hh), k2 = k * k;
  T s0 = hh - r.o.y;
  T A = r.d.x * r.d.x + r.d.z * r.d.z - k2 * r.d.y * r.d.y;
  T B = r.o.x * r.d.x + r.o.z * r.d.z + k2 * s0 * r.d.y;
  T C = r.o.x * r.o.x + r.o.z * r.o.z - k2 * s0 * s0;
  auto nappe_s = [&](T t) { return s0 - t * r.d.y; };
  auto lateral_n = [&](T t) {
    V3<T> p = r.o + r.d * t;
    return normalize(V3<T>(p.x, k2 * (hh - p.y), p.z));
  };
  T lo = -inf, hi = inf;  // interval of the single-nappe solid cone
  if (A == T(0)) {
    if (B == T(0)) {
      if (C > T(0)) return false;
    } else {
      T t0 = -C / (T(2) * B);
      if (B > T(0))
        hi = t0;
      else
        lo = t0;
      T probe = (B > T(0)) ? t0 - T(1) : t0 + T(1);
      if (nappe_s(probe) < T(0)) return false;
    }
  } else {
    T disc = B * B - A * C;
    if (disc < T(0)) return false;
    T sq = std::sqrt(disc);
    T t1 = (-B - sq) / A, t2 = (-B + sq) / A;
    if (t1 > t2) std::swap(t1, t2);
    if (A > T(0)) {
      if (nappe_s((t1 + t2) * T(0.5)) < T(0)) return false;
      lo = t1, hi = t2;
    } else {
      // f <= 0 outside [t1,t2]; one half-line per nappe.  s decreases with t when d.y > 0.
      if (r.d.y > T(0))
        hi = t1;
      else
        lo = t2;
    }
  }
  Span<T> s;
  if (lo > -inf) s.clip_in(lo, lateral_n(lo));
  if (hi < inf) s.clip_out(hi, lateral_n(hi));
  // base half-space y >= -hh
  if (r.d.y == T(0)) {
    if (r.o.y < -hh) return false;
  } else {
    T tb = (-hh - r.o.y) / r.d.y;
    if (r.d.y > T(0))
      s.clip_in(tb, V3<T>(0, -1, 0));
    else
      s.clip_out(tb, V3<T>(0, -1, 0));
  }
  return finish_span(s, solid, out);
}

// Capsule: segment (0, +-hh, 0) swept by a ball of radius rad = union of a clipped cylinder and
// two balls; the union of the three line intervals is one interval because the solid is convex.
template <class T>
bool cast_capsule(T hh, T rad, const Ray<T> &r, bool solid, Inter<T> &out) {
  const T inf = std::numeric_limits<T>::infinity();
  T tin = inf, tout = -inf;
  {
    Span<T> s;
    clip_slab_y(s, r, -hh, hh);
    if (!s.empty) clip_inf_cylinder(s, r, rad);
    if (!s.empty && s.tin <= s.tout) tin = std::min(tin, s.tin), tout = std::max(tout, s.tout);
  }
  T a, b;
  if (line_ball(r, hh, rad, a, b)) tin = std::min(tin, a), tout = std::max(tout, b);
  if (line_ball(r, -hh, rad, a, b)) tin = std::min(tin, a), tout = std::max(tout, b);
  if (tin > tout) return false;
  auto nrm = [&](T t) {
    V3<T> p = r.o + r.d * t;
    T cy = std::min(std::max(p.y, -hh), hh);
    return normalize(p - V3<T>(0, cy, 0));
  };
  Span<T> s;
  s.tin = tin, s.tout = tout;
  s.nin = nrm(tin), s.nout = nrm(tout);
  return finish_span(s, solid, out);
}

// ---- Triangle — SURVEY B.8 (ncollide3d triangle_ray_intersection) -------------------------------
template <class T>
bool cast_triangle(const V3<T> &a, const V3<T> &b, const V3<T> &c, const Ray<T> &ray, T &toi, V3<T> &normal,
                   V3<T> &bary) {
  V3<T> ab = b - a, ac = c - a;
  V3<T> n = cross(ab, ac);
  T d = dot(n, ray.d);
  if (d == T(0)) return false;
  V3<T> ap = ray.o - a;
  T t = dot(ap, n);
  if ((t < T(0) && d < T(0)) || (t > T(0) && d > T(0))) return false;
  d = std::abs(d);
  V3<T> e = -cross(ray.d, ap);
  T v, w;
  if (t < T(0)) {
    v = -dot(ac, e);
    if (v < T(0) || v > d) return false;
    w = dot(ab, e);
    if (w < T(0) || v + w > d) return false;
    T invd = T(1) / d;
    toi = -t * invd;
    normal = -normalize(n);
    v *= invd, w *= invd;
  } else {
    v = dot(ac, e);
    if (v < T(0) || v > d) return false;
    w = -dot(ab, e);
    if (w < T(0) || v + w > d) return false;
    T invd = T(1) / d;
    toi = t * invd;
    normal = normalize(n);
    v *= invd, w *= invd;
  }
  bary = V3<T>(-v - w + T(1), v, w);
  return true;
}

// ------------------------------------------------------------------------------------------
// Textures / materials
// ------------------------------------------------------------------------------------------
struct Texture {
  uint32_t w = 0, h = 0;
  int interp = 0, overflow = 0;
  const C4 *px = nullptr;  // points into Scene::texels
  uint64_t avail = 0;      // texels from px to the end of the pool (D4)

  // Texture2d::at — src/texture2d.rs:203-205 (flat index y*W + x, not clamped per axis)
  const C4 &at(uint64_t x, uint64_t y) const {
    uint64_t i = y * (uint64_t)w + x;
    if (i >= avail) i = avail - 1;  // D4
    return px[i];
  }

  // Texture2d::sample — src/texture2d.rs:207-256
  C4 sample(double cu, double cv) const {
    float ux = (float)cu, uy = (float)cv;
    if (overflow == NRB_OVERFLOW_CLAMP) {
      ux = std::min(std::max(ux, 0.0f), 1.0f);
      uy = std::min(std::max(uy, 0.0f), 1.0f);
    } else {
      ux = std::fmod(ux, 1.0f);
      uy = std::fmod(uy, 1.0f);
      if (ux < 0.0f) ux = 1.0f + ux;
      if (uy < 0.0f) uy = 1.0f + uy;
    }
    ux = ux * (float)(w - 1);
    uy = uy * (float)(h - 1);
    if (interp == NRB_INTERP_NEAREST) {
      return at((uint64_t)std::round(ux), (uint64_t)std::round(uy));
    }
    uint64_t lx = (uint64_t)std::floor(ux), ly = (uint64_t)std::floor(uy);
    uint64_t hx = lx + 1, hy = ly + 1;
    float sx = ux - (float)lx, sy = uy - (float)ly;
    const C4 &ul = at(lx, hy), &ur = at(hx, hy), &dr = at(hx, ly), &dl = at(lx, ly);
    C4 up = {ul.x * (1.0f - sx) + ur.x * sx, ul.y * (1.0f - sx) + ur.y * sx, ul.z * (1.0f - sx) + ur.z * sx,
             ul.w * (1.0f - sx) + ur.w * sx};
    C4 dn = {dl.x * (1.0f - sx) + dr.x * sx, dl.y * (1.0f - sx) + dr.y * sx, dl.z * (1.0f - sx) + dr.z * sx,
             dl.w * (1.0f - sx) + dr.w * sx};
    C4 o = {up.x * sy + dn.x * (1.0f - sy), up.y * sy + dn.y * (1.0f - sy), up.z * sy + dn.z * (1.0f - sy),
            up.w * sy + dn.w * (1.0f - sy)};
    return o;
  }
};

struct Material {
  int kind = 0;
  C3 ambient, diffuse, specular;
  float shininess = 0;
  int tex = -1, alpha_tex = -1;
};

template <class T>
struct Light {
  V3<T> pos;
  T radius;
  uint32_t racsample;
  C3 color;
};

// ------------------------------------------------------------------------------------------
// Scene
// ------------------------------------------------------------------------------------------
template <class T>
struct Node {  // SceneNode — src/scene_node.rs:8-19
  int shape = 0, material = 0;
  T param[3];
  T rot[9];
  V3<T> trans;
  T refr_coeff = 1;
  float refl_mix = 0, refl_att = 0, alpha = 1;
  bool solid = false;
  int nmap = -1;
  // trimesh: TriMesh::new(coords, faces, Some(uvs)) — loader3d.rs:695
  std::vector<V3<T>> verts;
  std::vector<T> uvs;           // 2 per vertex
  std::vector<uint32_t> faces;  // 3 per triangle
  Bvt<T> bvt;                   // inner BVT over triangle AABBs
  Aabb<T> aabb;                 // geometry.bounding_volume(&transform) — src/scene_node.rs:41

  V3<T> to_local_vec(const V3<T> &v) const {  // R^T v
    return V3<T>(rot[0] * v.x + rot[3] * v.y + rot[6] * v.z, rot[1] * v.x + rot[4] * v.y + rot[7] * v.z,
                 rot[2] * v.x + rot[5] * v.y + rot[8] * v.z);
  }
  V3<T> to_world_vec(const V3<T> &v) const {  // R v
    return V3<T>(rot[0] * v.x + rot[1] * v.y + rot[2] * v.z, rot[3] * v.x + rot[4] * v.y + rot[5] * v.z,
                 rot[6] * v.x + rot[7] * v.y + rot[8] * v.z);
  }
  Ray<T> to_local(const Ray<T> &r) const {  // ray.inverse_transform_by(m)
    Ray<T> l;
    l.o = to_local_vec(r.o - trans);
    l.d = to_local_vec(r.d);
    return l;
  }
};

struct Counters {
  uint64_t primary = 0, reflect = 0, refract = 0, shadow = 0, truncated = 0;
  void add(const Counters &o) {
    primary += o.primary, reflect += o.reflect, refract += o.refract, shadow += o.shadow, truncated += o.truncated;
  }
};

template <class T>
struct RayWithEnergy {  // src/ray_with_energy.rs:4-22
  Ray<T> ray;
  T refr = 1;
  float energy = 1;
};

struct RngCtx {  // D1: counter state carried down the recursion;  D3: depth cap
  uint32_t pixel = 0, sample = 0, path = 1, depth = 0;
  uint64_t seed = 0;
  uint32_t max_depth = 64;
};

struct TriHit {
  double toi;
  double n[3];
  double b[3];
};

template <class T>
struct Scene {
  std::vector<Node<T>> nodes;
  std::vector<Light<T>> lights;
  std::vector<Material> materials;
  std::vector<Texture> textures;
  std::vector<C4> texels;
  C3 background;
  Bvt<T> world;  // BVT<Arc<SceneNode>, AABB> — src/scene.rs:24

  // ---- SceneNode::cast — src/scene_node.rs:51-75 -------------------------------------------
  bool cast(const Node<T> &n, const Ray<T> &r, Inter<T> &out) const {
    bool hit = false;
    out = Inter<T>();
    switch (n.shape) {
      case NRB_SHAPE_BALL:  // centre = translation; rotation ignored
        hit = cast_ball(n.trans, n.param[0], r, n.solid, out);
        break;
      case NRB_SHAPE_CUBOID:
        hit = cast_cuboid(V3<T>(n.param[0], n.param[1], n.param[2]), n.to_local(r), n.solid, out);
        if (hit) out.normal = n.to_world_vec(out.normal);
        break;
      case NRB_SHAPE_CYLINDER:
        hit = cast_cylinder(n.param[0], n.param[1], n.to_local(r), n.solid, out);
        if (hit) out.normal = n.to_world_vec(out.normal);
        break;
      case NRB_SHAPE_CAPSULE:
        hit = cast_capsule(n.param[0], n.param[1], n.to_local(r), n.solid, out);
        if (hit) out.normal = n.to_world_vec(out.normal);
        break;
      case NRB_SHAPE_CONE:
        hit = cast_cone(n.param[0], n.param[1], n.to_local(r), n.solid, out);
        if (hit) out.normal = n.to_world_vec(out.normal);
        break;
      case NRB_SHAPE_PLANE:
        hit = cast_plane(V3<T>(n.param[0], n.param[1], n.param[2]), n.to_local(r), n.solid, out);
        if (hit) out.normal = n.to_world_vec(out.normal);
        break;
      case NRB_SHAPE_TRIMESH: {
        // TriMesh::toi_and_normal_and_uv_with_ray — SURVEY B.8: local ray, inner BVT best-first
        // search, bv cost = AABB toi (solid), leaf cost = triangle toi; `solid` ignored;
        // uv = barycentric blend of the three vertex uvs.
        struct TH {
          T toi;
          V3<T> n, b;
        } th;
        Ray<T> l = n.to_local(r);
        int best = n.bvt.template best_first_search<TH>(
            [&](const Aabb<T> &bv, T &c) { return aabb_toi(bv, l, true, c); },
            [&](int tri, T &c, TH &d) {
              const uint32_t *f = &n.faces[3 * (size_t)tri];
              if (!cast_triangle(n.verts[f[0]], n.verts[f[1]], n.verts[f[2]], l, d.toi, d.n, d.b)) return false;
              c = d.toi;
              return true;
            },
            th);
        if (best >= 0) {
          const uint32_t *f = &n.faces[3 * (size_t)best];
          out.toi = th.toi;
          out.normal = n.to_world_vec(th.n);
          out.has_uv = true;
          out.u = n.uvs[2 * f[0]] * th.b.x + n.uvs[2 * f[1]] * th.b.y + n.uvs[2 * f[2]] * th.b.z;
          out.v = n.uvs[2 * f[0] + 1] * th.b.x + n.uvs[2 * f[1] + 1] * th.b.y + n.uvs[2 * f[2] + 1] * th.b.z;
          hit = true;
        }
        break;
      }
      default:
        break;
    }
    if (!hit) return false;
    // nmap depth shift — src/scene_node.rs:60-70
    if (n.nmap >= 0 && out.has_uv) {
      C4 sc = textures[n.nmap].sample((double)out.u, (double)out.v);
      float shift = (sc.x + sc.y + sc.z) / 3.0f;
      out.toi = out.toi - (T)shift;
    }
    return true;
  }

  // ---- world.best_first_search(ClosestRayTOICostFn) — src/scene.rs:164-166, 262-283 -----------
  int closest(const Ray<T> &r, Inter<T> &inter) const {
    return world.template best_first_search<Inter<T>>(
        [&](const Aabb<T> &bv, T &c) { return aabb_toi(bv, r, true, c); },
        [&](int leaf, T &c, Inter<T> &d) {
          if (!cast(nodes[leaf], r, d)) return false;
          c = d.toi;
          return true;
        },
        inter);
  }

  // ---- Material::ambiant -------------------------------------------------------------------
  C4 ambiant(const Material &m, const Inter<T> &it) const {
    switch (m.kind) {
      case NRB_MAT_NORMAL:  // src/normal_material.rs:7-15
        return C4{(1.0f + (float)it.normal.x) / 2.0f, (1.0f + (float)it.normal.y) / 2.0f,
                  (1.0f + (float)it.normal.z) / 2.0f, 1.0f};
      case NRB_MAT_UV:  // src/uv_material.rs:8-21 (None -> origin = (0,0,0,0))
        if (it.has_uv) return C4{(float)it.u, (float)it.v, 0.0f, 1.0f};
        return C4{0, 0, 0, 0};
      default: {  // PhongMaterial::ambiant — src/phong_material.rs:39-70
        if (it.has_uv) {
          C4 tc = {1, 1, 1, 1};
          if (m.tex >= 0) {
            tc = textures[m.tex].sample((double)it.u, (double)it.v);
            tc.w = 1.0f;
          }
          if (m.alpha_tex >= 0) tc.w = textures[m.alpha_tex].sample((double)it.u, (double)it.v).w;
          return C4{m.ambient.x * tc.x, m.ambient.y * tc.y, m.ambient.z * tc.z, 1.0f * tc.w};
        }
        return C4{m.ambient.x, m.ambient.y, m.ambient.z, 1.0f};
      }
    }
  }

  // ---- Scene::intersects_ray + TransparentShadowsRayTOICostFn — src/scene.rs:147-161, 285-339 --
  // Returns true for Some(filter).
  bool intersects_ray(const Ray<T> &r, T maxtoi, C3 &filter) const {
    filter = C3(1, 1, 1);
    struct Unit {};
    Unit u;
    int hit = world.template best_first_search<Unit>(
        [&](const Aabb<T> &bv, T &c) { return aabb_toi(bv, r, true, c); },
        [&](int leaf, T &c, Unit &) {
          Inter<T> t;
          const Node<T> &b = nodes[leaf];
          if (!cast(b, r, t)) return false;
          if (t.toi <= maxtoi) {
            C4 color = ambiant(materials[b.material], t);
            float alpha = color.w * b.alpha;
            if (alpha < 1.0f) {
              filter = cmul(filter, C3(color.x, color.y, color.z)) * (1.0f - alpha);
              return false;
            }
            c = t.toi;
            return true;
          }
          return false;
        },
        u);
    return hit < 0;
  }

  // ---- PhongMaterial::compute — src/phong_material.rs:72-151 (other materials: Material::compute
  //      default = ambiant, src/material.rs:8-16) ------------------------------------------------
  C4 compute(const Material &m, const RayWithEnergy<T> &ray, const V3<T> &point, const Inter<T> &it,
             const RngCtx &ctx, Counters &cnt) const {
    if (m.kind != NRB_MAT_PHONG) return ambiant(m, it);
    C4 tex_color = {1, 1, 1, 1};
    float alpha = 1.0f;
    if (it.has_uv && m.tex >= 0) tex_color = textures[m.tex].sample((double)it.u, (double)it.v);
    if (it.has_uv && m.alpha_tex >= 0) alpha = textures[m.alpha_tex].sample((double)it.u, (double)it.v).w;
    C3 tex(tex_color.x, tex_color.y, tex_color.z);
    C3 res = cmul(m.ambient, tex);
    const V3<T> &normal = it.normal;
    for (size_t li = 0; li < lights.size(); ++li) {
      const Light<T> &light = lights[li];
      C3 acc(0, 0, 0);
      uint32_t ns = light.racsample * light.racsample;
      for (uint32_t k = 0; k < ns; ++k) {
        // Light::sample — src/light.rs:56-63: pos + random::<Vect>() * radius  (D1 for the RNG)
        uint32_t ctr[4] = {ctx.pixel, ctx.sample, ctx.path, ((uint32_t)li << 16) | (k & 0xFFFFu)};
        // path doubles per bounce and wraps after 32 of them: the key takes depth / 32 so deeper paths keep their own streams
        uint32_t key[2] = {(uint32_t)ctx.seed, (uint32_t)(ctx.seed >> 32) ^ STREAM_LIGHT ^ ((ctx.depth >> 5) * 0x9E3779B9u)};
        uint32_t rnd[4];
        philox4x32_10(ctr, key, rnd);
        V3<T> pos = light.pos + V3<T>((T)u24(rnd[0]), (T)u24(rnd[1]), (T)u24(rnd[2])) * light.radius;
        V3<T> ldir = pos - point;
        T len = norm(ldir);
        ldir = ldir / len;
        T dist = len - T(0.001);
        C3 filter;
        Ray<T> sray;
        sray.o = point + ldir * T(0.001);
        sray.d = ldir;
        cnt.shadow++;
        if (!intersects_ray(sray, dist, filter)) continue;
        T dot_ldir_norm = dot(ldir, normal);
        float dcoeff = std::max((float)dot_ldir_norm, 0.0f);
        C3 diffuse_color = cmul(m.diffuse, tex);
        C3 diffuse = diffuse_color * dcoeff;
        V3<T> lproj = normal * dot_ldir_norm;
        V3<T> rldir = normalize(-ldir + lproj * T(2));
        float scoeff = (float)(-dot(rldir, ray.ray.d));
        if (scoeff > 0.0f) {
          scoeff = std::pow(scoeff, m.shininess);
          C3 specular = m.specular * scoeff;
          acc = acc + cmul(light.color, cmul(filter, diffuse + specular));
        } else {
          acc = acc + cmul(light.color, cmul(filter, diffuse));
        }
      }
      float a = 1.0f / (float)(light.racsample * light.racsample);
      res = acc * a + res;  // res.axpy(a, &acc, 1.0)
    }
    return C4{res.x, res.y, res.z, alpha};
  }

  // ---- Scene::trace — src/scene.rs:163-193 ----------------------------------------------------
  C3 trace(const RayWithEnergy<T> &ray, const RngCtx &ctx, Counters &cnt) const {
    Inter<T> inter;
    int sn = closest(ray.ray, inter);
    if (sn < 0) return background;
    const Node<T> &n = nodes[sn];
    V3<T> pt = ray.ray.o + ray.ray.d * inter.toi;
    C4 obj = compute(materials[n.material], ray, pt, inter, ctx, cnt);
    C3 refl = trace_reflection(n.refl_mix, n.refl_att, ray, pt, inter.normal, ctx, cnt);
    float alpha = obj.w * n.alpha;
    C3 obj_color = C3(obj.x, obj.y, obj.z) * (1.0f - n.refl_mix) + refl * n.refl_mix;
    C3 refr = trace_refraction(alpha, n.refr_coeff, ray, pt, inter.normal, ctx, cnt);
    if (alpha == 1.0f) return obj_color;
    return obj_color * alpha + refr * (1.0f - alpha);
  }

  // ---- Scene::trace_reflection — src/scene.rs:196-218 -----------------------------------------
  C3 trace_reflection(float mix, float attenuation, const RayWithEnergy<T> &ray, const V3<T> &pt,
                      const V3<T> &normal, const RngCtx &ctx, Counters &cnt) const {
    if (mix != 0.0f && ray.energy > 0.1f) {
      if (ctx.depth + 1 >= ctx.max_depth) {  // D3
        cnt.truncated++;
        return C3(0, 0, 0);
      }
      V3<T> nproj = normal * dot(ray.ray.d, normal);
      V3<T> rdir = ray.ray.d - nproj * T(2);
      RayWithEnergy<T> nr;
      nr.ray.o = pt + rdir * T(0.001);
      nr.ray.d = rdir;
      nr.refr = ray.refr;
      nr.energy = ray.energy - attenuation;
      RngCtx c2 = ctx;
      c2.path = ctx.path * 2u;
      c2.depth = ctx.depth + 1;
      cnt.reflect++;
      return trace(nr, c2, cnt);
    }
    return C3(0, 0, 0);
  }

  // ---- Scene::trace_refraction — src/scene.rs:221-252 -----------------------------------------
  C3 trace_refraction(float alpha, T coeff, const RayWithEnergy<T> &ray, const V3<T> &pt, const V3<T> &normal,
                      const RngCtx &ctx, Counters &cnt) const {
    if (alpha != 1.0f) {
      if (ctx.depth + 1 >= ctx.max_depth) {  // D3
        cnt.truncated++;
        return C3(0, 0, 0);
      }
      T n1, n2;
      if (ray.refr == T(1)) {
        n1 = 1, n2 = coeff;
      } else {
        n1 = coeff, n2 = 1;
      }
      V3<T> along = normal * dot(ray.ray.d, normal);
      V3<T> tangent = ray.ray.d - along;
      V3<T> new_dir = normalize(along + tangent * (n2 / n1));
      RayWithEnergy<T> nr;
      nr.ray.o = pt + new_dir * T(0.001);
      nr.ray.d = new_dir;
      nr.refr = n2;
      nr.energy = ray.energy;
      RngCtx c2 = ctx;
      c2.path = ctx.path * 2u + 1u;
      c2.depth = ctx.depth + 1;
      cnt.refract++;
      return trace(nr, c2, cnt);
    }
    return C3(0, 0, 0);
  }
};

// ---- primary ray — src/scene.rs:68-86 (always evaluated in f64 then narrowed to T) ----------------
bool primary_ray(const NrbCamera &cam, uint32_t pixel, uint32_t sample, double o[3], double d[3]) {
  uint32_t resx = cam.width, resy = cam.height;
  uint32_t j = pixel / resx, i = pixel - j * resx;
  uint32_t ctr[4] = {pixel, sample, 0u, 0u};
  uint32_t key[2] = {(uint32_t)cam.seed, (uint32_t)(cam.seed >> 32) ^ STREAM_PRIMARY};
  uint32_t rnd[4];
  philox4x32_10(ctr, key, rnd);
  double px = (u24(rnd[0]) - 0.5) * cam.window_width;
  double py = (u24(rnd[1]) - 0.5) * cam.window_width;
  double ox = (double)i + px, oy = (double)j + py;
  double dx = (ox / (double)resx - 0.5) * 2.0;
  double dy = -(oy / (double)resy - 0.5) * 2.0;
  const double *M = cam.projection;  // column-major
  double s[4] = {dx, dy, -1.0, 1.0};
  double h[4];
  for (int r = 0; r < 4; ++r) h[r] = M[r] * s[0] + M[4 + r] * s[1] + M[8 + r] * s[2] + M[12 + r] * s[3];
  if (h[3] == 0.0) return false;  // Point3::from_homogeneous(..).unwrap() would panic (src/scene.rs:85)
  double e[3] = {h[0] / h[3], h[1] / h[3], h[2] / h[3]};
  double v[3] = {e[0] - cam.eye[0], e[1] - cam.eye[1], e[2] - cam.eye[2]};
  double len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  for (int k = 0; k < 3; ++k) o[k] = cam.eye[k], d[k] = v[k] / len;
  return true;
}

// ---- scene::render — src/scene.rs:29-116 ----------------------------------------------------------
template <class T>
int render(const Scene<T> &scene, const NrbCamera &cam, int n_threads, uint64_t first, uint64_t count, float *out,
           NrbStats *stats) {
  uint64_t npixels = (uint64_t)cam.width * cam.height;
  if (count == 0) first = 0, count = npixels;
  if (first + count > npixels) {
    g_err = "pixel range outside the image";
    return NRB_ERR_INVALID_ARG;
  }
  unsigned num_thread = n_threads > 0 ? (unsigned)n_threads : std::max(1u, std::thread::hardware_concurrency());
  uint32_t spp = cam.ray_per_pixel;
  uint32_t max_depth = cam.max_depth ? cam.max_depth : 64u;
  std::vector<Counters> counters(num_thread);
  std::vector<std::thread> children;
  std::atomic<int> failed(0);
  for (unsigned ti = 0; ti < num_thread; ++ti) {
    // contiguous ranges of npixels/num_thread + 1 pixels — src/scene.rs:61-63
    uint64_t parts = count / num_thread + 1;
    uint64_t low = std::min(parts * ti, count), up = std::min(parts * (ti + 1), count);
    children.emplace_back([&, ti, low, up]() {
      Counters &cnt = counters[ti];
      for (uint64_t ip = low; ip < up; ++ip) {
        uint32_t ipt = (uint32_t)(first + ip);
        C3 tot(0, 0, 0);
        for (uint32_t s = 0; s < spp; ++s) {
          double o[3], d[3];
          if (!primary_ray(cam, ipt, s, o, d)) {
            failed = 1;
            continue;
          }
          RayWithEnergy<T> r;  // RayWithEnergy::new: refr 1.0, energy 1.0
          r.ray.o = V3<T>((T)o[0], (T)o[1], (T)o[2]);
          r.ray.d = V3<T>((T)d[0], (T)d[1], (T)d[2]);
          RngCtx ctx;
          ctx.pixel = ipt, ctx.sample = s, ctx.seed = cam.seed, ctx.max_depth = max_depth;
          cnt.primary++;
          C3 c = scene.trace(r, ctx, cnt);
          tot = tot + c;
        }
        C3 px = tot / (float)spp;
        out[3 * (size_t)ipt + 0] = px.x;
        out[3 * (size_t)ipt + 1] = px.y;
        out[3 * (size_t)ipt + 2] = px.z;
      }
    });
  }
  for (auto &c : children) c.join();
  if (failed) {
    g_err = "projection produced w == 0 (from_homogeneous would panic)";
    return NRB_ERR_INVALID_ARG;
  }
  if (stats) {
    Counters tot;
    for (auto &c : counters) tot.add(c);
    std::memset(stats, 0, sizeof(*stats));
    stats->rays_primary = tot.primary;
    stats->rays_reflect = tot.reflect;
    stats->rays_refract = tot.refract;
    stats->rays_shadow = tot.shadow;
    stats->paths_truncated = tot.truncated;
  }
  return NRB_OK;
}

// ---- Scene::new — src/scene.rs:119-133, from the flattened tables -----------------------------------
template <class T>
bool build_scene(const NrbSceneDesc &d, Scene<T> &sc) {
  const T big = std::numeric_limits<T>::max() / T(2);
  sc.background = C3(d.background[0], d.background[1], d.background[2]);
  sc.texels.resize(d.n_texels);
  if (d.n_texels) std::memcpy(sc.texels.data(), d.texels, sizeof(C4) * d.n_texels);
  for (uint32_t i = 0; i < d.n_textures; ++i) {
    const NrbTextureDesc &t = d.textures[i];
    if (t.width < 1 || t.height < 1 || t.texel_offset + (uint64_t)t.width * t.height > d.n_texels) {
      g_err = "texture outside the texel pool";
      return false;
    }
    Texture tx;
    tx.w = t.width, tx.h = t.height, tx.interp = t.interpolation, tx.overflow = t.overflow;
    tx.px = sc.texels.data() + t.texel_offset;
    tx.avail = d.n_texels - t.texel_offset;
    sc.textures.push_back(tx);
  }
  for (uint32_t i = 0; i < d.n_materials; ++i) {
    const NrbMaterialDesc &m = d.materials[i];
    Material mm;
    mm.kind = m.kind;
    mm.ambient = C3(m.ambient[0], m.ambient[1], m.ambient[2]);
    mm.diffuse = C3(m.diffuse[0], m.diffuse[1], m.diffuse[2]);
    mm.specular = C3(m.specular[0], m.specular[1], m.specular[2]);
    mm.shininess = m.shininess;
    mm.tex = m.texture, mm.alpha_tex = m.alpha_texture;
    if (mm.tex >= (int)d.n_textures || mm.alpha_tex >= (int)d.n_textures) {
      g_err = "material texture index out of range";
      return false;
    }
    sc.materials.push_back(mm);
  }
  for (uint32_t i = 0; i < d.n_lights; ++i) {
    const NrbLightDesc &l = d.lights[i];
    Light<T> ll;
    ll.pos = V3<T>((T)l.pos[0], (T)l.pos[1], (T)l.pos[2]);
    ll.radius = (T)l.radius;
    ll.racsample = l.racsample;
    ll.color = C3(l.color[0], l.color[1], l.color[2]);
    sc.lights.push_back(ll);
  }
  sc.nodes.resize(d.n_nodes);
  std::vector<Aabb<T>> node_bvs(d.n_nodes);
  for (uint32_t i = 0; i < d.n_nodes; ++i) {
    const NrbNodeDesc &nd = d.nodes[i];
    Node<T> &n = sc.nodes[i];
    if (nd.material < 0 || nd.material >= (int)d.n_materials) {
      g_err = "node material index out of range";
      return false;
    }
    n.shape = nd.shape, n.material = nd.material;
    for (int k = 0; k < 3; ++k) n.param[k] = (T)nd.param[k];
    for (int k = 0; k < 9; ++k) n.rot[k] = (T)nd.rot[k];
    n.trans = V3<T>((T)nd.trans[0], (T)nd.trans[1], (T)nd.trans[2]);
    n.refr_coeff = (T)nd.refr_coeff;
    n.refl_mix = nd.refl_mix, n.refl_att = nd.refl_atenuation, n.alpha = nd.alpha;
    n.solid = nd.solid != 0;
    n.nmap = nd.nmap_texture;
    if (n.nmap >= (int)d.n_textures) {
      g_err = "nmap texture index out of range";
      return false;
    }
    // bounding volumes (HasBoundingVolume::bounding_volume(&transform), src/scene_node.rs:41)
    V3<T> lhe;  // local half extents
    V3<T> lc;   // local centre
    bool infinite = false;
    switch (nd.shape) {
      case NRB_SHAPE_BALL:
        n.aabb.mins = n.trans - V3<T>(n.param[0], n.param[0], n.param[0]);
        n.aabb.maxs = n.trans + V3<T>(n.param[0], n.param[0], n.param[0]);
        node_bvs[i] = n.aabb;
        continue;
      case NRB_SHAPE_CUBOID:
        lhe = V3<T>(n.param[0], n.param[1], n.param[2]);
        break;
      case NRB_SHAPE_CYLINDER:
      case NRB_SHAPE_CONE:
        lhe = V3<T>(n.param[1], n.param[0], n.param[1]);
        break;
      case NRB_SHAPE_CAPSULE:
        lhe = V3<T>(n.param[1], n.param[0] + n.param[1], n.param[1]);
        break;
      case NRB_SHAPE_PLANE:
        infinite = true;  // +-(MAX/2) on all axes
        break;
      case NRB_SHAPE_TRIMESH: {
        if (nd.first_index % 3 != 0 || nd.first_index + 3 * nd.tri_count > d.n_indices || nd.tri_count == 0) {
          g_err = "trimesh index range invalid";
          return false;
        }
        // re-index the referenced vertices so each node holds its own compact arrays
        std::vector<int64_t> remap;
        uint64_t vmax = 0;
        for (uint64_t k = 0; k < 3 * nd.tri_count; ++k) vmax = std::max<uint64_t>(vmax, d.indices[nd.first_index + k]);
        if (nd.vertex_base + vmax >= d.n_vertices) {
          g_err = "trimesh vertex index out of range";
          return false;
        }
        remap.assign(vmax + 1, -1);
        n.faces.resize(3 * nd.tri_count);
        for (uint64_t k = 0; k < 3 * nd.tri_count; ++k) {
          uint32_t vi = d.indices[nd.first_index + k];
          if (remap[vi] < 0) {
            remap[vi] = (int64_t)n.verts.size();
            const float *p = d.positions + 3 * (nd.vertex_base + vi);
            n.verts.push_back(V3<T>((T)p[0], (T)p[1], (T)p[2]));
            if (d.uvs) {
              n.uvs.push_back((T)d.uvs[2 * (nd.vertex_base + vi)]);
              n.uvs.push_back((T)d.uvs[2 * (nd.vertex_base + vi) + 1]);
            } else {
              n.uvs.push_back(0), n.uvs.push_back(0);
            }
          }
          n.faces[k] = (uint32_t)remap[vi];
        }
        std::vector<Aabb<T>> tb(nd.tri_count);
        for (uint64_t t = 0; t < nd.tri_count; ++t) {
          Aabb<T> b;
          b.mins = b.maxs = n.verts[n.faces[3 * t]];
          for (int k = 1; k < 3; ++k) {
            Aabb<T> p;
            p.mins = p.maxs = n.verts[n.faces[3 * t + k]];
            b.merge(p);
          }
          tb[t] = b;
        }
        n.bvt.build(tb);
        const Aabb<T> &rb = n.bvt.nodes[n.bvt.root].bv;
        lc = rb.center();
        lhe = (rb.maxs - rb.mins) * T(0.5);
        break;
      }
      default:
        g_err = "unknown shape kind";
        return false;
    }
    if (infinite) {
      n.aabb.mins = V3<T>(-big, -big, -big);
      n.aabb.maxs = V3<T>(big, big, big);
    } else {
      // AABB::transform_by: centre transformed, half extents multiplied by |R|
      V3<T> c = n.to_world_vec(lc) + n.trans;
      V3<T> he;
      for (int r = 0; r < 3; ++r)
        he[r] = std::abs(n.rot[3 * r]) * lhe.x + std::abs(n.rot[3 * r + 1]) * lhe.y + std::abs(n.rot[3 * r + 2]) * lhe.z;
      n.aabb.mins = c - he;
      n.aabb.maxs = c + he;
    }
    node_bvs[i] = n.aabb;
  }
  sc.world.build(node_bvs);
  return true;
}

}  // namespace

struct NroScene {
  int bits;
  std::unique_ptr<Scene<double>> s64;
  std::unique_ptr<Scene<float>> s32;
};

template <class T>
static int cast_impl(const Scene<T> &sc, uint32_t node, const double o[3], const double d[3], double *toi,
                     double normal[3], double uv[2], int *uv_present) {
  if (node >= sc.nodes.size()) return 0;
  Ray<T> r;
  r.o = V3<T>((T)o[0], (T)o[1], (T)o[2]);
  r.d = V3<T>((T)d[0], (T)d[1], (T)d[2]);
  Inter<T> it;
  if (!sc.cast(sc.nodes[node], r, it)) return 0;
  *toi = (double)it.toi;
  normal[0] = (double)it.normal.x, normal[1] = (double)it.normal.y, normal[2] = (double)it.normal.z;
  uv[0] = (double)it.u, uv[1] = (double)it.v;
  *uv_present = it.has_uv ? 1 : 0;
  return 1;
}

template <class T>
static int trace_impl(const Scene<T> &sc, const NrbCamera *cam, const double o[3], const double d[3], uint32_t pixel,
                      uint32_t sample, float rgb[3]) {
  RayWithEnergy<T> r;
  r.ray.o = V3<T>((T)o[0], (T)o[1], (T)o[2]);
  r.ray.d = V3<T>((T)d[0], (T)d[1], (T)d[2]);
  RngCtx ctx;
  ctx.pixel = pixel, ctx.sample = sample, ctx.seed = cam->seed, ctx.max_depth = cam->max_depth ? cam->max_depth : 64u;
  Counters cnt;
  C3 c = sc.trace(r, ctx, cnt);
  rgb[0] = c.x, rgb[1] = c.y, rgb[2] = c.z;
  return NRB_OK;
}

template <class T>
static int shadow_impl(const Scene<T> &sc, const double o[3], const double d[3], double maxtoi, float filter[3]) {
  Ray<T> r;
  r.o = V3<T>((T)o[0], (T)o[1], (T)o[2]);
  r.d = V3<T>((T)d[0], (T)d[1], (T)d[2]);
  C3 f;
  bool some = sc.intersects_ray(r, (T)maxtoi, f);
  filter[0] = f.x, filter[1] = f.y, filter[2] = f.z;
  return some ? 1 : 0;
}

extern "C" {

const char *nro_last_error(void) { return g_err.c_str(); }

int nro_scene_create(const NrbSceneDesc *desc, int precision_bits, NroScene **out) {
  if (!desc || !out || (precision_bits != 64 && precision_bits != 32) || desc->struct_size != sizeof(NrbSceneDesc) ||
      desc->abi_version != NRB_ABI_VERSION) {
    g_err = "bad descriptor / precision";
    return NRB_ERR_INVALID_ARG;
  }
  std::unique_ptr<NroScene> s(new NroScene);
  s->bits = precision_bits;
  bool ok;
  if (precision_bits == 64) {
    s->s64.reset(new Scene<double>);
    ok = build_scene(*desc, *s->s64);
  } else {
    s->s32.reset(new Scene<float>);
    ok = build_scene(*desc, *s->s32);
  }
  if (!ok) return NRB_ERR_INVALID_ARG;
  *out = s.release();
  return NRB_OK;
}

void nro_scene_destroy(NroScene *s) { delete s; }

int nro_render(NroScene *s, const NrbCamera *cam, int n_threads, uint64_t first_pixel, uint64_t n_pixels,
               float *out_rgb, NrbStats *stats) {
  if (!s || !cam || !out_rgb || cam->ray_per_pixel == 0 || cam->width == 0 || cam->height == 0) {
    g_err = "invalid argument (ray_per_pixel must be > 0: src/scene.rs:37)";
    return NRB_ERR_INVALID_ARG;
  }
  if (s->bits == 64) return render(*s->s64, *cam, n_threads, first_pixel, n_pixels, out_rgb, stats);
  return render(*s->s32, *cam, n_threads, first_pixel, n_pixels, out_rgb, stats);
}

int nro_cast(NroScene *s, uint32_t node, const double o[3], const double d[3], double *toi, double normal[3],
             double uv[2], int *uv_present) {
  if (s->bits == 64) return cast_impl(*s->s64, node, o, d, toi, normal, uv, uv_present);
  return cast_impl(*s->s32, node, o, d, toi, normal, uv, uv_present);
}

int nro_trace(NroScene *s, const NrbCamera *cam, const double o[3], const double d[3], uint32_t pixel,
              uint32_t sample, float rgb[3]) {
  if (s->bits == 64) return trace_impl(*s->s64, cam, o, d, pixel, sample, rgb);
  return trace_impl(*s->s32, cam, o, d, pixel, sample, rgb);
}

int nro_intersects_ray(NroScene *s, const double o[3], const double d[3], double maxtoi, float filter[3]) {
  if (s->bits == 64) return shadow_impl(*s->s64, o, d, maxtoi, filter);
  return shadow_impl(*s->s32, o, d, maxtoi, filter);
}

int nro_texture_sample(NroScene *s, uint32_t texture, double u, double v, float rgba[4]) {
  const std::vector<Texture> &tx = s->bits == 64 ? s->s64->textures : s->s32->textures;
  if (texture >= tx.size()) {
    g_err = "texture index out of range";
    return NRB_ERR_INVALID_ARG;
  }
  C4 c = tx[texture].sample(u, v);
  rgba[0] = c.x, rgba[1] = c.y, rgba[2] = c.z, rgba[3] = c.w;
  return NRB_OK;
}

int nro_aabb_toi(const double mins[3], const double maxs[3], const double o[3], const double d[3], int solid,
                 double *toi) {
  Aabb<double> b;
  b.mins = V3<double>(mins[0], mins[1], mins[2]);
  b.maxs = V3<double>(maxs[0], maxs[1], maxs[2]);
  Ray<double> r;
  r.o = V3<double>(o[0], o[1], o[2]);
  r.d = V3<double>(d[0], d[1], d[2]);
  double t;
  if (!aabb_toi(b, r, solid != 0, t)) return 0;
  *toi = t;
  return 1;
}

int nro_primary_ray(const NrbCamera *cam, uint32_t pixel, uint32_t sample, double o[3], double d[3]) {
  return primary_ray(*cam, pixel, sample, o, d) ? NRB_OK : NRB_ERR_INVALID_ARG;
}

void nro_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }

}  // extern "C"
