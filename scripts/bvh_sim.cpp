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
      bool hit = traverse(root_all, eye, d, 3.4e38f, false, t);
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
          traverse(root_opaque != kEmpty ? root_opaque : root_all, p + l * 0.001f, l, len - 0.001f, true, ts);
          sv += n_visits, st += n_tests, sr += 1;
          // the candidate phases of the shadow query: one closest-hit traversal per candidate sub-root
          n_visits = n_tests = 0;
          for (const Cand &c : cands) traverse(c.root, p + l * 0.001f, l, len - 0.001f, false, ts);
          cv += n_visits, ct += n_tests;
          // alternative: ONE pass over the unified tree that never shrinks its interval (every node overlapping the segment)
          n_visits = n_tests = 0;
          g_no_shrink = true;
          traverse(root_all, p + l * 0.001f, l, len - 0.001f, false, ts);
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
