// bvh_sim.cpp — developer tool (not part of the product): replays the device traversal loop on the CPU
// over a BVH dumped by `NRB_DUMP_BVH=file nrb_scene_validate(...)` and counts node visits / triangle
// tests per ray for a pinhole camera (primary rays) and for the shadow segments hit -> light.
// Used to compare builder settings without GPU time.
//
//   g++ -O2 -std=c++17 -o /tmp/bvh_sim scripts/bvh_sim.cpp
//   /tmp/bvh_sim dump.bin  ex ey ez  ax ay az  fovy  W H  step  [lx ly lz]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct F4 { float x, y, z, w; };
struct I4 { int x, y, z, w; };
struct Node { F4 n0, n1, n2; I4 n3; };
struct Tri { F4 t0, t1, t2; };
static const int kEmpty = 0x7FFFFFFF;

struct V3 { float x, y, z; };
static V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static V3 norm(V3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }

static std::vector<Node> nodes;
static std::vector<Tri> tris;
static uint64_t n_visits, n_tests, n_rays;
static bool g_no_shrink = false;

static bool tri_hit(const Tri &T, V3 o, V3 d, float tlimit, bool inclusive, float &toi) {
  V3 v0{T.t0.x, T.t0.y, T.t0.z}, e1{T.t1.x, T.t1.y, T.t1.z}, e2{T.t2.x, T.t2.y, T.t2.z};
  V3 n = cross(e1, e2);
  float dd = dot(n, d);
  V3 ap = o - v0;
  float t = dot(ap, n);
  if (dd == 0.0f) return false;
  if ((t < 0 && dd < 0) || (t > 0 && dd > 0)) return false;
  float D = std::fabs(dd), at = std::fabs(t);
  if (inclusive ? !(at <= tlimit * D) : !(at < tlimit * D)) return false;
  V3 e = cross(ap, d);
  float s = t < 0 ? -1.0f : 1.0f;
  float v = s * dot(e2, e), w = -s * dot(e1, e);
  if (v < 0 || v > D || w < 0 || v + w > D) return false;
  toi = at / D;
  return true;
}

static bool traverse(int root, V3 o, V3 d, float tmax, bool any, float &tout) {
  int stack[128], sp = 0;
  stack[0] = kEmpty;
  int node = root;
  bool found = false;
  auto inv = [](float v) { return 1.0f / (std::fabs(v) > 1e-24f ? v : std::copysign(1e-24f, v)); };
  float idx = inv(d.x), idy = inv(d.y), idz = inv(d.z);
  float ox = o.x * idx, oy = o.y * idy, oz = o.z * idz;
  float tbest = tmax;
  ++n_rays;
  while (node != kEmpty) {
    while ((unsigned)node < (unsigned)kEmpty) {
      ++n_visits;
      const Node &N = nodes[node];
      auto slab = [&](float lox, float hix, float loy, float hiy, float loz, float hiz, float &tmin) {
        float ax = lox * idx - ox, bx = hix * idx - ox, ay = loy * idy - oy, by = hiy * idy - oy, az = loz * idz - oz, bz = hiz * idz - oz;
        tmin = std::fmax(std::fmax(std::fmin(ax, bx), std::fmin(ay, by)), std::fmax(std::fmin(az, bz), 0.0f));
        float tmx = std::fmin(std::fmin(std::fmax(ax, bx), std::fmax(ay, by)), std::fmin(std::fmax(az, bz), tbest));
        return tmx >= tmin;
      };
      float m0, m1;
      bool h0 = slab(N.n0.x, N.n0.y, N.n0.z, N.n0.w, N.n2.x, N.n2.y, m0);
      bool h1 = slab(N.n1.x, N.n1.y, N.n1.z, N.n1.w, N.n2.z, N.n2.w, m1);
      if (!h0 && !h1) {
        node = stack[sp--];
      } else {
        node = h0 ? N.n3.x : N.n3.y;
        if (h0 && h1) {
          int far = N.n3.y;
          if (m1 < m0) far = N.n3.x, node = N.n3.y;
          stack[++sp] = far;
        }
      }
    }
    while (node < 0) {
      uint32_t code = (uint32_t)~node, first = code >> 3, cnt = ((code >> 1) & 3u) + 1u;
      if (!(code & 1u)) {
        for (uint32_t k = 0; k < cnt; ++k) {
          ++n_tests;
          float toi;
          if (tri_hit(tris[first + k], o, d, tbest, any, toi)) {
            if (!g_no_shrink) tbest = toi;
            found = true;
            if (any) { tout = toi; return true; }
          }
        }
      }
      node = stack[sp--];
    }
  }
  tout = tbest;
  return found;
}


// ---- wide-BVH what-if (WIDE=4|8 [QUANT=1]): the BVH2 collapsed into nodes of up to WIDE children (largest-area child
// expanded first), optionally with 8-bit boxes quantised outward on a per-node power-of-two grid.  Same counters.
struct WNode { int n; float lo[8][3], hi[8][3]; int child[8]; };
static std::vector<WNode> wnodes;
static int g_wide = 0, g_quant = 0;
static void child_box(const Node &N, int k, float lo[3], float hi[3]) {
  if (k == 0) lo[0] = N.n0.x, hi[0] = N.n0.y, lo[1] = N.n0.z, hi[1] = N.n0.w, lo[2] = N.n2.x, hi[2] = N.n2.y;
  else lo[0] = N.n1.x, hi[0] = N.n1.y, lo[1] = N.n1.z, hi[1] = N.n1.w, lo[2] = N.n2.z, hi[2] = N.n2.w;
}
static int collapse(int code) {  // code: BVH2 child code; returns wide child code (internal -> wnodes index)
  if (code < 0 || code == kEmpty) return code;
  WNode w;
  w.n = 2;
  const Node &N = nodes[code];
  child_box(N, 0, w.lo[0], w.hi[0]), child_box(N, 1, w.lo[1], w.hi[1]);
  w.child[0] = N.n3.x, w.child[1] = N.n3.y;
  while (w.n < g_wide) {
    int best = -1;
    float ba = -1;
    for (int i = 0; i < w.n; ++i)
      if (w.child[i] >= 0 && w.child[i] != kEmpty) {
        float dx = w.hi[i][0] - w.lo[i][0], dy = w.hi[i][1] - w.lo[i][1], dz = w.hi[i][2] - w.lo[i][2];
        float a = dx * dy + dy * dz + dz * dx;
        if (a > ba) ba = a, best = i;
      }
    if (best < 0) break;
    const Node &C = nodes[w.child[best]];
    child_box(C, 0, w.lo[best], w.hi[best]);
    child_box(C, 1, w.lo[w.n], w.hi[w.n]);
    w.child[best] = C.n3.x, w.child[w.n] = C.n3.y;
    ++w.n;
  }
  if (g_quant) {
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (int i = 0; i < w.n; ++i)
      for (int a = 0; a < 3; ++a) lo[a] = std::fmin(lo[a], w.lo[i][a]), hi[a] = std::fmax(hi[a], w.hi[i][a]);
    for (int a = 0; a < 3; ++a) {
      float ext = hi[a] - lo[a];
      int e;
      std::frexp(ext / 255.0f, &e);
      float step = std::ldexp(1.0f, e);  // >= ext / 255
      for (int i = 0; i < w.n; ++i) {
        float ql = std::floor((w.lo[i][a] - lo[a]) / step), qh = std::ceil((w.hi[i][a] - lo[a]) / step);
        w.lo[i][a] = lo[a] + ql * step, w.hi[i][a] = lo[a] + qh * step;
      }
    }
  }
  int id = (int)wnodes.size();
  wnodes.push_back(w);
  for (int i = 0; i < w.n; ++i) {
    int c = collapse(wnodes[id].child[i]);
    wnodes[id].child[i] = c;
  }
  return id;
}
static uint64_t n_pushes;
static bool traverse_wide(int root, V3 o, V3 d, float tmax, bool any, float &tout) {
  int stack[256], sp = 0;
  stack[0] = kEmpty;
  int node = root;
  bool found = false;
  auto inv = [](float v) { return 1.0f / (std::fabs(v) > 1e-24f ? v : std::copysign(1e-24f, v)); };
  float idx = inv(d.x), idy = inv(d.y), idz = inv(d.z);
  float ox = o.x * idx, oy = o.y * idy, oz = o.z * idz;
  float tbest = tmax;
  ++n_rays;
  while (node != kEmpty) {
    while ((unsigned)node < (unsigned)kEmpty) {
      ++n_visits;
      const WNode &N = wnodes[node];
      float tm[8];
      int ord[8], nh = 0;
      for (int i = 0; i < N.n; ++i) {
        float ax = N.lo[i][0] * idx - ox, bx = N.hi[i][0] * idx - ox, ay = N.lo[i][1] * idy - oy, by = N.hi[i][1] * idy - oy,
              az = N.lo[i][2] * idz - oz, bz = N.hi[i][2] * idz - oz;
        float tmin = std::fmax(std::fmax(std::fmin(ax, bx), std::fmin(ay, by)), std::fmax(std::fmin(az, bz), 0.0f));
        float tmx = std::fmin(std::fmin(std::fmax(ax, bx), std::fmax(ay, by)), std::fmin(std::fmax(az, bz), tbest));
        if (tmx >= tmin) {
          int j = nh++;
          while (j > 0 && tm[ord[j - 1]] > tmin) ord[j] = ord[j - 1], --j;
          ord[j] = i, tm[i] = tmin;
        }
      }
      if (nh == 0) node = stack[sp--];
      else {
        for (int j = nh - 1; j >= 1; --j) stack[++sp] = N.child[ord[j]], ++n_pushes;
        node = N.child[ord[0]];
      }
    }
    while (node < 0) {
      uint32_t code = (uint32_t)~node, first = code >> 3, cnt = ((code >> 1) & 3u) + 1u;
      if (!(code & 1u)) {
        for (uint32_t k = 0; k < cnt; ++k) {
          ++n_tests;
          float toi;
          if (tri_hit(tris[first + k], o, d, tbest, any, toi)) {
            if (!g_no_shrink) tbest = toi;
            found = true;
            if (any) { tout = toi; return true; }
          }
        }
      }
      node = stack[sp--];
    }
  }
  tout = tbest;
  return found;
}

static bool trav(int root, V3 o, V3 d, float tmax, bool any, float &tout) {
  return g_wide ? traverse_wide(root, o, d, tmax, any, tout) : traverse(root, o, d, tmax, any, tout);
}

int main(int argc, char **argv) {
  if (argc < 12) return fprintf(stderr, "usage: see header\n"), 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return perror("open"), 1;
  uint64_t hdr[4];
  if (fread(hdr, sizeof(hdr), 1, f) != 1) return 1;
  nodes.resize(hdr[0]);
  tris.resize(hdr[1]);
  if (fread(nodes.data(), sizeof(Node), nodes.size(), f) != nodes.size()) return 1;
  if (fread(tris.data(), sizeof(Tri), tris.size(), f) != tris.size()) return 1;
  struct Cand { float lo[3], hi[3]; int root, node; };
  std::vector<Cand> cands;
  {
    uint64_t nc = 0;
    if (fread(&nc, sizeof(nc), 1, f) == 1 && nc < 4096) {
      cands.resize(nc);
      if (fread(cands.data(), sizeof(Cand), nc, f) != nc) cands.clear();
    }
  }
  fclose(f);
  int root_all = (int)(uint32_t)hdr[2], root_opaque = (int)(uint32_t)hdr[3];
  if (getenv("GRID16")) {  // what-if: BVH2 boxes snapped outward to a global 16-bit grid (+ GRID16 cells of padding)
    int pad = atoi(getenv("GRID16"));
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (const Node &N : nodes)
      for (int k = 0; k < 2; ++k) {
        float l[3], h[3];
        child_box(N, k, l, h);
        for (int a = 0; a < 3; ++a) lo[a] = std::fmin(lo[a], l[a]), hi[a] = std::fmax(hi[a], h[a]);
      }
    float cell[3];
    for (int a = 0; a < 3; ++a) cell[a] = (hi[a] - lo[a]) / 65000.0f;
    auto ql = [&](float v, int a) { return lo[a] + (std::floor((v - lo[a]) / cell[a]) - pad) * cell[a]; };
    auto qh = [&](float v, int a) { return lo[a] + (std::ceil((v - lo[a]) / cell[a]) + pad) * cell[a]; };
    for (Node &N : nodes) {
      N.n0.x = ql(N.n0.x, 0), N.n0.y = qh(N.n0.y, 0), N.n0.z = ql(N.n0.z, 1), N.n0.w = qh(N.n0.w, 1), N.n2.x = ql(N.n2.x, 2), N.n2.y = qh(N.n2.y, 2);
      N.n1.x = ql(N.n1.x, 0), N.n1.y = qh(N.n1.y, 0), N.n1.z = ql(N.n1.z, 1), N.n1.w = qh(N.n1.w, 1), N.n2.z = ql(N.n2.z, 2), N.n2.w = qh(N.n2.w, 2);
    }
    printf("grid16: extent %g %g %g, pad %d cells\n", hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2], pad);
  }
  if (getenv("WIDE")) {
    g_wide = atoi(getenv("WIDE")), g_quant = getenv("QUANT") ? atoi(getenv("QUANT")) : 0;
    int ra = collapse(root_all);
    root_opaque = root_opaque == root_all ? ra : collapse(root_opaque);
    root_all = ra;
    for (Cand &c : cands) c.root = collapse(c.root);
    double fill = 0;
    for (const WNode &w : wnodes) fill += w.n;
    printf("wide %d quant %d: %zu wide nodes, %.2f children per node\n", g_wide, g_quant, wnodes.size(), fill / wnodes.size());
  }
  V3 eye{(float)atof(argv[2]), (float)atof(argv[3]), (float)atof(argv[4])};
  V3 at{(float)atof(argv[5]), (float)atof(argv[6]), (float)atof(argv[7])};
  float fovy = (float)atof(argv[8]) * 3.14159265f / 180.0f;
  int W = atoi(argv[9]), H = atoi(argv[10]), step = atoi(argv[11]);
  V3 light = eye;
  if (argc >= 15) light = V3{(float)atof(argv[12]), (float)atof(argv[13]), (float)atof(argv[14])};
  V3 fw = norm(at - eye), rt = norm(cross(fw, V3{0, 1, 0})), up = cross(rt, fw);
  float th = std::tan(fovy / 2), aspect = (float)W / H;
  uint64_t pv = 0, pt = 0, pr = 0, sv = 0, st = 0, sr = 0, hits = 0, cv = 0, ct = 0, uv_ = 0, ut = 0;
  for (int y = 0; y < H; y += step)
    for (int x = 0; x < W; x += step) {
      float nx = ((float)x / W - 0.5f) * 2, ny = -((float)y / H - 0.5f) * 2;
      V3 d = norm(fw + rt * (nx * aspect * th) + up * (ny * th));
      n_visits = n_tests = n_rays = 0;
      float t;
      bool hit = trav(root_all, eye, d, 3.4e38f, false, t);
      pv += n_visits, pt += n_tests, pr += 1;
      if (hit) {
        ++hits;
        V3 p = eye + d * t;
        V3 l = light - p;
        float len = std::sqrt(dot(l, l));
        if (len > 0.002f) {
          l = l * (1.0f / len);
          n_visits = n_tests = 0;
          float ts;
          trav(root_opaque != kEmpty ? root_opaque : root_all, p + l * 0.001f, l, len - 0.001f, true, ts);
          sv += n_visits, st += n_tests, sr += 1;
          // the candidate phases of the shadow query: one closest-hit traversal per candidate sub-root
          n_visits = n_tests = 0;
          for (const Cand &c : cands) trav(c.root, p + l * 0.001f, l, len - 0.001f, false, ts);
          cv += n_visits, ct += n_tests;
          // alternative: ONE pass over the unified tree that never shrinks its interval (every node overlapping the segment)
          n_visits = n_tests = 0;
          g_no_shrink = true;
          trav(root_all, p + l * 0.001f, l, len - 0.001f, false, ts);
          g_no_shrink = false;
          uv_ += n_visits, ut += n_tests;
        }
      }
    }
  if (sr) printf("shadow candidates (%zu): %.2f visits %.2f tests per ray | unified no-shrink pass: %.2f visits %.2f tests\n", cands.size(),
                 (double)cv / sr, (double)ct / sr, (double)uv_ / sr, (double)ut / sr);
  printf("nodes %zu tris %zu | primary: %.2f visits %.2f tests (%llu rays, %.1f%% hit) | shadow: %.2f visits %.2f tests (%llu rays)\n",
         nodes.size(), tris.size(), (double)pv / pr, (double)pt / pr, (unsigned long long)pr, 100.0 * hits / pr,
         sr ? (double)sv / sr : 0.0, sr ? (double)st / sr : 0.0, (unsigned long long)sr);
  return 0;
}
