// device_math.cuh — f32 device arithmetic of the render path: vector helpers, Philox4x32-10,
// the per-shape ray casts (SceneNode::cast, src/scene_node.rs:51-75 → ncollide3d RayCast, SURVEY
// Appendix B), the two-sided triangle test and Texture2d::sample (src/texture2d.rs:207-256).
#pragma once
#include "../../include/nrays_b200.h"
#include "device_types.cuh"

namespace nrb {

#define NRB_DI __device__ __forceinline__

struct V3 {
  float x, y, z;
};
NRB_DI V3 mk(float x, float y, float z) { return V3{x, y, z}; }
NRB_DI V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
NRB_DI V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
NRB_DI V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
NRB_DI V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
NRB_DI V3 cmul(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
NRB_DI float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
NRB_DI V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
NRB_DI V3 normalize(V3 a) {
  float inv = 1.0f / sqrtf(dot(a, a));  // correctly rounded: ray directions decide silhouette / shadow-edge pixels
  return a * inv;
}
NRB_DI V3 normalize_fast(V3 a) { return a * rsqrtf(dot(a, a)); }  // MUFU.RSQ, <= 2 ulp: shading-only vectors
NRB_DI float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

NRB_DI uint32_t fdiv(uint32_t n, const FastDiv &f) {  // n / f.d, exact (device_types.cuh)
  if (f.d == 1u) return n;
  unsigned long long lo = (unsigned long long)f.mul_lo * n;
  unsigned long long hi = (unsigned long long)f.mul_hi * n + (lo >> 32);
  return (uint32_t)(hi >> 32);
}

// ---- Philox4x32-10 (same function as the oracle; known-answer vectors in tests) -----------------
NRB_DI void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0, c1 = lo1, c2 = n2, c3 = lo0;
    k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
  }
  out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}
constexpr uint32_t kStreamPrimary = 0x50524D59u;  // 'PRMY'
constexpr uint32_t kStreamLight = 0x4C474854u;    // 'LGHT'
NRB_DI float u24(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// ---- intersection record -------------------------------------------------------------------------
struct Inter {
  float toi;
  V3 n;
  float u, v;
  bool has_uv;
};

// isometry helpers (row-major rot)
NRB_DI V3 rot_t(const float *R, V3 v) {  // R^T v
  return V3{R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z,
            R[2] * v.x + R[5] * v.y + R[8] * v.z};
}
NRB_DI V3 rot(const float *R, V3 v) {  // R v
  return V3{R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z,
            R[6] * v.x + R[7] * v.y + R[8] * v.z};
}

// ---- Ball — SURVEY B.4 -------------------------------------------------------------------------------
NRB_DI bool cast_ball(V3 c, float radius, V3 o, V3 d, bool solid, Inter &out) {
  V3 dc = o - c;
  float a = dot(d, d);
  float b = dot(dc, d);
  float cc = dot(dc, dc) - radius * radius;
  if (cc > 0.0f && b > 0.0f) return false;
  float delta = b * b - a * cc;
  if (delta < 0.0f) return false;
  float sq = sqrtf(delta);
  float t = (-b - sq) / a;
  bool inside = false;
  if (t <= 0.0f) {
    inside = true;
    t = solid ? 0.0f : (-b + sq) / a;
  }
  V3 n = normalize(o + d * t - c);
  const float pi = 3.14159265358979323846f;
  out.toi = t;
  out.has_uv = true;
  out.u = 0.5f + atan2f(n.z, n.x) / (pi + pi);
  out.v = 0.5f - asinf(fminf(fmaxf(n.y, -1.0f), 1.0f)) / pi;
  out.n = inside ? -n : n;
  return true;
}

// ---- Cuboid — SURVEY B.5 (local frame) -------------------------------------------------------------
NRB_DI bool cast_cuboid(V3 he, V3 o, V3 d, bool solid, Inter &out) {
  float tmax = 3.402823466e+38f, tmin = -3.402823466e+38f;
  int near_side = 0, far_side = 0;
  bool near_diag = false, far_diag = false;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float di = comp(d, i), oi = comp(o, i), h = comp(he, i);
    if (di == 0.0f) {
      if (oi < -h || oi > h) return false;
    } else {
      float denom = 1.0f / di;
      float tn = (-h - oi) * denom, tf = (h - oi) * denom;
      bool flip = false;
      if (tn > tf) {
        flip = true;
        float tmp = tn;
        tn = tf;
        tf = tmp;
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
      if (tmax < 0.0f || tmin > tmax) return false;
    }
  }
  int side;
  float toi;
  V3 n = mk(0, 0, 0);
  if (tmin < 0.0f) {
    side = far_side;
    if (solid) {
      toi = 0.0f;
    } else {
      toi = tmax;
      if (far_diag) {
        n = -normalize(d);
      } else {
        int ax = (far_side < 0 ? -far_side : far_side) - 1;
        float s = far_side < 0 ? -1.0f : 1.0f;
        n = mk(ax == 0 ? s : 0.0f, ax == 1 ? s : 0.0f, ax == 2 ? s : 0.0f);
      }
    }
  } else {
    side = near_side;
    toi = tmin;
    if (near_diag) {
      n = -normalize(d);
    } else {
      int ax = (near_side < 0 ? -near_side : near_side) - 1;
      float s = near_side < 0 ? 1.0f : -1.0f;
      n = mk(ax == 0 ? s : 0.0f, ax == 1 ? s : 0.0f, ax == 2 ? s : 0.0f);
    }
  }
  out.toi = toi;
  out.n = n;
  V3 pt = o + d * toi;
  V3 dpt = pt + he;
  V3 scale = he * 2.0f;
  int id = (side < 0 ? -side : side) - 1;
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

// ---- Plane — SURVEY B.7 (local frame) ----------------------------------------------------------------
NRB_DI bool cast_plane(V3 n, V3 o, V3 d, bool solid, Inter &out) {
  float dn = -dot(n, o);
  out.has_uv = false;
  out.u = out.v = 0.0f;
  if (solid && dn > 0.0f) {
    out.toi = 0.0f;
    out.n = mk(0, 0, 0);
    return true;
  }
  float t = dn / dot(n, d);
  if (!(t >= 0.0f)) return false;
  if (!(t < 3.402823466e+38f)) return false;
  out.toi = t;
  out.n = dn > 0.0f ? -n : n;
  return true;
}

// ---- closed-form convex solids of revolution (ncollide3d uses GJK) -------------------
struct Span {
  float tin, tout;
  V3 nin, nout;
  bool empty;
};
NRB_DI void span_init(Span &s) {
  s.tin = -INFINITY, s.tout = INFINITY;
  s.nin = s.nout = mk(0, 0, 0);
  s.empty = false;
}
NRB_DI void clip_in(Span &s, float t, V3 n) {
  if (t > s.tin) s.tin = t, s.nin = n;
}
NRB_DI void clip_out(Span &s, float t, V3 n) {
  if (t < s.tout) s.tout = t, s.nout = n;
}
NRB_DI void clip_slab_y(Span &s, V3 o, V3 d, float lo, float hi) {
  if (d.y == 0.0f) {
    if (o.y < lo || o.y > hi) s.empty = true;
    return;
  }
  float t0 = (lo - o.y) / d.y, t1 = (hi - o.y) / d.y;
  if (d.y > 0.0f) {
    clip_in(s, t0, mk(0, -1, 0));
    clip_out(s, t1, mk(0, 1, 0));
  } else {
    clip_in(s, t1, mk(0, 1, 0));
    clip_out(s, t0, mk(0, -1, 0));
  }
}
NRB_DI void clip_inf_cylinder(Span &s, V3 o, V3 d, float rad) {
  float A = d.x * d.x + d.z * d.z;
  float B = o.x * d.x + o.z * d.z;
  float C = o.x * o.x + o.z * o.z - rad * rad;
  if (A == 0.0f) {
    if (C > 0.0f) s.empty = true;
    return;
  }
  float disc = B * B - A * C;
  if (disc < 0.0f) {
    s.empty = true;
    return;
  }
  float sq = sqrtf(disc);
  float t0 = (-B - sq) / A, t1 = (-B + sq) / A;
  V3 p0 = o + d * t0, p1 = o + d * t1;
  clip_in(s, t0, normalize(mk(p0.x, 0, p0.z)));
  clip_out(s, t1, normalize(mk(p1.x, 0, p1.z)));
}
NRB_DI bool line_ball(V3 o, V3 d, float cy, float rad, float &t0, float &t1) {
  V3 dc = o - mk(0, cy, 0);
  float a = dot(d, d), b = dot(dc, d), c = dot(dc, dc) - rad * rad;
  float disc = b * b - a * c;
  if (disc < 0.0f) return false;
  float sq = sqrtf(disc);
  t0 = (-b - sq) / a;
  t1 = (-b + sq) / a;
  return true;
}
NRB_DI bool finish_span(const Span &s, bool solid, Inter &out) {
  out.has_uv = false;
  out.u = out.v = 0.0f;
  if (s.empty || s.tin > s.tout || s.tout < 0.0f) return false;
  if (s.tin > 0.0f) {
    out.toi = s.tin;
    out.n = s.nin;
    return true;
  }
  if (solid) {
    out.toi = 0.0f;
    out.n = mk(0, 0, 0);
    return true;
  }
  if (!(s.tout < 3.402823466e+38f)) return false;
  out.toi = s.tout;
  out.n = s.nout;
  return true;
}
NRB_DI bool cast_cylinder(float hh, float rad, V3 o, V3 d, bool solid, Inter &out) {
  Span s;
  span_init(s);
  clip_slab_y(s, o, d, -hh, hh);
  if (!s.empty) clip_inf_cylinder(s, o, d, rad);
  return finish_span(s, solid, out);
}
NRB_DI bool cast_cone(float hh, float rad, V3 o, V3 d, bool solid, Inter &out) {
  float k = rad / (2.0f * hh), k2 = k * k;
  float s0 = hh - o.y;
  float A = d.x * d.x + d.z * d.z - k2 * d.y * d.y;
  float B = o.x * d.x + o.z * d.z + k2 * s0 * d.y;
  float C = o.x * o.x + o.z * o.z - k2 * s0 * s0;
  float lo = -INFINITY, hi = INFINITY;
  if (A == 0.0f) {
    if (B == 0.0f) {
      if (C > 0.0f) return false;
    } else {
      float t0 = -C / (2.0f * B);
      if (B > 0.0f)
        hi = t0;
      else
        lo = t0;
      float probe = (B > 0.0f) ? t0 - 1.0f : t0 + 1.0f;
      if (s0 - probe * d.y < 0.0f) return false;
    }
  } else {
    float disc = B * B - A * C;
    if (disc < 0.0f) return false;
    float sq = sqrtf(disc);
    float t1 = (-B - sq) / A, t2 = (-B + sq) / A;
    if (t1 > t2) {
      float tmp = t1;
      t1 = t2;
      t2 = tmp;
    }
    if (A > 0.0f) {
      if (s0 - 0.5f * (t1 + t2) * d.y < 0.0f) return false;
      lo = t1, hi = t2;
    } else {
      if (d.y > 0.0f)
        hi = t1;
      else
        lo = t2;
    }
  }
  Span s;
  span_init(s);
  if (lo > -INFINITY) {
    V3 p = o + d * lo;
    clip_in(s, lo, normalize(mk(p.x, k2 * (hh - p.y), p.z)));
  }
  if (hi < INFINITY) {
    V3 p = o + d * hi;
    clip_out(s, hi, normalize(mk(p.x, k2 * (hh - p.y), p.z)));
  }
  if (d.y == 0.0f) {
    if (o.y < -hh) return false;
  } else {
    float tb = (-hh - o.y) / d.y;
    if (d.y > 0.0f)
      clip_in(s, tb, mk(0, -1, 0));
    else
      clip_out(s, tb, mk(0, -1, 0));
  }
  return finish_span(s, solid, out);
}
NRB_DI bool cast_capsule(float hh, float rad, V3 o, V3 d, bool solid, Inter &out) {
  float tin = INFINITY, tout = -INFINITY;
  {
    Span s;
    span_init(s);
    clip_slab_y(s, o, d, -hh, hh);
    if (!s.empty) clip_inf_cylinder(s, o, d, rad);
    if (!s.empty && s.tin <= s.tout) tin = fminf(tin, s.tin), tout = fmaxf(tout, s.tout);
  }
  float a, b;
  if (line_ball(o, d, hh, rad, a, b)) tin = fminf(tin, a), tout = fmaxf(tout, b);
  if (line_ball(o, d, -hh, rad, a, b)) tin = fminf(tin, a), tout = fmaxf(tout, b);
  if (tin > tout) return false;
  Span s;
  span_init(s);
  s.tin = tin, s.tout = tout;
  {
    V3 p = o + d * tin;
    s.nin = normalize(p - mk(0, fminf(fmaxf(p.y, -hh), hh), 0));
    p = o + d * tout;
    s.nout = normalize(p - mk(0, fminf(fmaxf(p.y, -hh), hh), 0));
  }
  return finish_span(s, solid, out);
}

// SceneNode::cast for an analytic shape (world-space ray in, world-space normal out)
__device__ __noinline__ bool cast_shape(const Shape &sh, V3 o, V3 d, Inter &out) {
  bool solid = sh.solid != 0;
  V3 tr = mk(sh.trans[0], sh.trans[1], sh.trans[2]);
  if (sh.kind == NRB_SHAPE_BALL) return cast_ball(tr, sh.p[0], o, d, solid, out);
  V3 lo = rot_t(sh.rot, o - tr), ld = rot_t(sh.rot, d);
  bool hit;
  switch (sh.kind) {
    case NRB_SHAPE_CUBOID:
      hit = cast_cuboid(mk(sh.p[0], sh.p[1], sh.p[2]), lo, ld, solid, out);
      break;
    case NRB_SHAPE_CYLINDER:
      hit = cast_cylinder(sh.p[0], sh.p[1], lo, ld, solid, out);
      break;
    case NRB_SHAPE_CAPSULE:
      hit = cast_capsule(sh.p[0], sh.p[1], lo, ld, solid, out);
      break;
    case NRB_SHAPE_CONE:
      hit = cast_cone(sh.p[0], sh.p[1], lo, ld, solid, out);
      break;
    default:
      hit = cast_plane(mk(sh.p[0], sh.p[1], sh.p[2]), lo, ld, solid, out);
      break;
  }
  if (hit) out.n = rot(sh.rot, out.n);
  return hit;
}

// ---- Triangle — SURVEY B.8 (triangle_ray_intersection), edge form ----------------------------------
// Returns toi and the barycentric (v, w) of vertices 1 and 2.  Hits with toi >= tlimit (closest-hit:
// best_first_search keeps the first of equal costs, strict <) or toi > tlimit (INCLUSIVE, shadow
// rays: `t.toi <= self.maxtoi`, src/scene.rs:313) are rejected before the edge tests.
template <bool INCLUSIVE>
NRB_DI bool cast_tri(V3 v0, V3 e1, V3 e2, V3 o, V3 d, float tlimit, float &toi, float &bv, float &bw) {
  V3 n = cross(e1, e2);
  float dd = dot(n, d);
  V3 ap = o - v0;
  float t = dot(ap, n);
  if (dd == 0.0f) return false;
  if ((t < 0.0f && dd < 0.0f) || (t > 0.0f && dd > 0.0f)) return false;
  float D = fabsf(dd);
  float at = fabsf(t);
  if (INCLUSIVE ? !(at <= tlimit * D) : !(at < tlimit * D)) return false;
  V3 e = cross(ap, d);  // = -(d x ap)
  float s = t < 0.0f ? -1.0f : 1.0f;
  float v = s * dot(e2, e);
  float w = -s * dot(e1, e);
  if (v < 0.0f || v > D || w < 0.0f || v + w > D) return false;
  float invd = 1.0f / D;
  toi = at * invd;
  bv = v * invd;
  bw = w * invd;
  return true;
}

// Same test with the comparison chosen at run time (the persistent trace loop mixes any-hit and closest-hit lanes).
NRB_DI bool cast_tri_rt(V3 v0, V3 e1, V3 e2, V3 o, V3 d, float tlimit, bool inclusive, float &toi, float &bv, float &bw) {
  V3 n = cross(e1, e2);
  float dd = dot(n, d);
  V3 ap = o - v0;
  float t = dot(ap, n);
  if (dd == 0.0f) return false;
  if ((t < 0.0f && dd < 0.0f) || (t > 0.0f && dd > 0.0f)) return false;
  float D = fabsf(dd);
  float at = fabsf(t);
  float lim = tlimit * D;
  if (inclusive ? !(at <= lim) : !(at < lim)) return false;
  V3 e = cross(ap, d);  // = -(d x ap)
  float s = t < 0.0f ? -1.0f : 1.0f;
  float v = s * dot(e2, e);
  float w = -s * dot(e1, e);
  if (v < 0.0f || v > D || w < 0.0f || v + w > D) return false;
  float invd = 1.0f / D;
  toi = at * invd;
  bv = v * invd;
  bw = w * invd;
  return true;
}

// ---- Texture2d::sample — src/texture2d.rs:207-256 -----------------------------------------------
NRB_DI float4 tex_at(const SceneView &sc, const Texture &t, uint32_t x, uint32_t y) {
  uint32_t i = y * t.w + x;
  if (i >= t.avail) i = t.avail - 1;  // the reference would panic here
  return __ldg(&sc.texels[t.offset + i]);
}
NRB_DI float4 tex_sample(const SceneView &sc, int tex, float cu, float cv) {
  const Texture t = sc.textures[tex];
  float ux = cu, uy = cv;
  if (t.overflow == NRB_OVERFLOW_CLAMP) {
    ux = fminf(fmaxf(ux, 0.0f), 1.0f);
    uy = fminf(fmaxf(uy, 0.0f), 1.0f);
  } else {
    ux = ux - truncf(ux);  // == fmodf(ux, 1.0f) bit for bit (the fraction of a float is exact), 2 instructions
    uy = uy - truncf(uy);
    if (ux < 0.0f) ux = 1.0f + ux;
    if (uy < 0.0f) uy = 1.0f + uy;
  }
  ux = ux * (float)(t.w - 1);
  uy = uy * (float)(t.h - 1);
  if (t.interp == NRB_INTERP_NEAREST) return tex_at(sc, t, (uint32_t)roundf(ux), (uint32_t)roundf(uy));
  float fx = floorf(ux), fy = floorf(uy);
  uint32_t lx = (uint32_t)fx, ly = (uint32_t)fy;
  uint32_t hx = lx + 1, hy = ly + 1;
  float sx = ux - fx, sy = uy - fy;
  float4 ul = tex_at(sc, t, lx, hy), ur = tex_at(sc, t, hx, hy), dr = tex_at(sc, t, hx, ly), dl = tex_at(sc, t, lx, ly);
  float ax = 1.0f - sx, ay = 1.0f - sy;
  // no fma contraction across the two products: the reference rounds each product (f32 mul, then add)
  float4 up = make_float4(__fadd_rn(__fmul_rn(ul.x, ax), __fmul_rn(ur.x, sx)), __fadd_rn(__fmul_rn(ul.y, ax), __fmul_rn(ur.y, sx)),
                          __fadd_rn(__fmul_rn(ul.z, ax), __fmul_rn(ur.z, sx)), __fadd_rn(__fmul_rn(ul.w, ax), __fmul_rn(ur.w, sx)));
  float4 dn = make_float4(__fadd_rn(__fmul_rn(dl.x, ax), __fmul_rn(dr.x, sx)), __fadd_rn(__fmul_rn(dl.y, ax), __fmul_rn(dr.y, sx)),
                          __fadd_rn(__fmul_rn(dl.z, ax), __fmul_rn(dr.z, sx)), __fadd_rn(__fmul_rn(dl.w, ax), __fmul_rn(dr.w, sx)));
  return make_float4(__fadd_rn(__fmul_rn(up.x, sy), __fmul_rn(dn.x, ay)), __fadd_rn(__fmul_rn(up.y, sy), __fmul_rn(dn.y, ay)),
                     __fadd_rn(__fmul_rn(up.z, sy), __fmul_rn(dn.z, ay)), __fadd_rn(__fmul_rn(up.w, sy), __fmul_rn(dn.w, ay)));
}

}  // namespace nrb
