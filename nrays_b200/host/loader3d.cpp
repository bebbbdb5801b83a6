// loader3d.cpp — C++ twin of the reference's only front end, examples/loader3d.rs (+ src/mtl.rs, src/obj.rs):
// parse a .scene file, assemble the scene with the reference's API names (nrays.hpp), render every camera
// through the C-ABI and save a PNG.
//
//   loader3d scene_file [--resolution W H] [--aa SPP WINDOW] [--seed S] [--validate]
//
// --validate: flatten + host-side BVH build only (nrb_scene_validate), print the table sizes, no device needed.
//
// The optional flags override what the scene file says (the BASELINE configs use other resolutions / aa than
// the shipped files: SURVEY F7).  Paths inside the scene file are relative to the current directory, as in
// the reference (its Makefile runs `cd scenes && ../target/release/loader3d X.scene`).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <unordered_map>

#include "nrays.hpp"
#include "png.hpp"

using namespace nrays;

namespace {

[[noreturn]] void error(size_t line, const std::string &err) {  // loader3d.rs:206-208
  throw std::runtime_error("At line " + std::to_string(line) + ": " + err);
}

std::vector<std::string> split_words(const std::string &s) {
  std::istringstream is(s);
  std::vector<std::string> w;
  std::string t;
  while (is >> t) w.push_back(t);
  return w;
}

std::string join(const std::vector<std::string> &w, size_t from) {
  std::string r;
  for (size_t i = from; i < w.size(); ++i) r += (i > from ? " " : "") + w[i];
  return r;
}

double parse_f64(size_t l, const std::string &s) {
  char *end = nullptr;
  double v = std::strtod(s.c_str(), &end);
  if (end == s.c_str() || *end) error(l, "failed to parse `" + s + "' as a f64.");
  return v;
}

std::string read_file(const std::string &path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("Unable to find the file: " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

// ---- src/mtl.rs ------------------------------------------------------------------------------------------
struct MtlMaterial {  // MtlMaterial::new_default (src/mtl.rs:147-162)
  std::string name;
  float shininess = 60.0f, alpha = 1.0f;
  float ambiant[3] = {1, 1, 1}, diffuse[3] = {1, 1, 1}, specular[3] = {1, 1, 1};
  std::string diffuse_texture, opacity_map;
};

std::vector<MtlMaterial> parse_mtl(const std::string &text) {  // src/mtl.rs:29-89
  std::vector<MtlMaterial> res;
  MtlMaterial cur;
  std::istringstream is(text);
  std::string line;
  size_t l = 0;
  for (; std::getline(is, line); ++l) {
    auto w = split_words(line);
    if (w.empty() || w[0][0] == '#' || w.size() < 2) continue;
    auto color = [&](float *dst) {
      if (w.size() < 4) error(l, "3 components were expected, found " + std::to_string(w.size() - 1) + ".");
      for (int k = 0; k < 3; ++k) dst[k] = (float)parse_f64(l, w[1 + k]);
    };
    if (w[0] == "newmtl") {
      if (!cur.name.empty()) res.push_back(cur);
      cur = MtlMaterial();
      cur.name = join(w, 1);
    } else if (w[0] == "Ka") color(cur.ambiant);
    else if (w[0] == "Kd") color(cur.diffuse);
    else if (w[0] == "Ks") color(cur.specular);
    else if (w[0] == "Ns") cur.shininess = (float)parse_f64(l, w[1]);
    else if (w[0] == "d") cur.alpha = (float)parse_f64(l, w[1]);
    else if (w[0] == "map_Kd") cur.diffuse_texture = join(w, 1);
    else if (w[0] == "map_d" || w[0] == "map_opacity") cur.opacity_map = join(w, 1);
  }
  if (!cur.name.empty()) res.push_back(cur);
  return res;
}

// ---- textures: Texture2d::from_png with the per-path cache of src/texture2d.rs:28-48 -----------------------
std::map<std::pair<std::string, bool>, std::shared_ptr<const ImageData>> g_texture_cache;

std::shared_ptr<Texture2d> texture_from_png(const std::string &path, bool opacity) {
  auto key = std::make_pair(path, opacity);
  auto it = g_texture_cache.find(key);
  if (it != g_texture_cache.end()) return std::make_shared<Texture2d>(Texture2d{it->second, Interpolation::Bilinear, Overflow::Wrap});
  png::Decoded d;
  if (!png::load(path, d)) throw std::runtime_error("Image not found: " + path);  // .expect("Image not found."), loader3d.rs:470
  Texture2d t = Texture2d::from_pixels(d.data.data(), d.w, d.h, d.depth, opacity, Interpolation::Bilinear, Overflow::Wrap);
  g_texture_cache[key] = t.data;
  return std::make_shared<Texture2d>(t);
}

std::shared_ptr<Material> phong_from_mtl(const MtlMaterial &m, const std::string &prefix) {
  auto path = [&](const std::string &p) { return prefix.empty() ? p : prefix + "/" + p; };
  std::shared_ptr<Texture2d> t, a;
  if (!m.diffuse_texture.empty()) t = texture_from_png(path(m.diffuse_texture), false);
  if (!m.opacity_map.empty()) a = texture_from_png(path(m.opacity_map), true);
  return std::make_shared<PhongMaterial>(m.ambiant, m.diffuse, m.specular, t, a, m.shininess);
}

// ---- src/obj.rs ------------------------------------------------------------------------------------------
struct ObjGroup {
  std::string name;
  std::vector<uint32_t> faces;
  bool has_mtl = false;
  MtlMaterial mtl;
};
struct ObjData {
  std::shared_ptr<MeshBuffers> mesh;
  std::vector<ObjGroup> groups;
};

ObjData parse_obj(const std::string &text, const std::string &mtl_base_dir, const std::string &basename) {  // src/obj.rs:62-120
  const int64_t MAXI = 2147483647;
  struct P3 {
    int64_t x, y, z;
    bool operator<(const P3 &o) const { return std::tie(x, y, z) < std::tie(o.x, o.y, o.z); }
  };
  std::vector<float> coords, uvs;
  size_t n_normals = 0;
  std::vector<std::string> group_names = {basename};
  std::unordered_map<std::string, size_t> groups = {{basename, 0}};
  std::vector<std::vector<P3>> groups_ids(1);
  size_t curr_group = 0;
  bool ignore_normals = false, ignore_uvs = false;
  std::unordered_map<std::string, MtlMaterial> mtllib;
  std::map<size_t, MtlMaterial> group2mtl;
  bool has_curr_mtl = false;
  MtlMaterial curr_mtl;
  auto parse_g = [&](const std::vector<std::string> &w, size_t from, const std::string &prefix) -> size_t {
    std::string suffix = join(w, from);
    std::string name = suffix.empty() ? prefix : prefix + "/" + suffix;
    auto it = groups.find(name);
    if (it != groups.end()) return it->second;
    groups_ids.emplace_back();
    group_names.push_back(name);
    return groups[name] = groups_ids.size() - 1;
  };
  std::istringstream is(text);
  std::string line;
  size_t l = 0;
  for (; std::getline(is, line); ++l) {
    auto w = split_words(line);
    if (w.empty() || w[0][0] == '#') continue;
    if (w[0] == "v") {
      if (w.size() < 4) error(l, "3 components were expected, found " + std::to_string(w.size() - 1) + ".");
      for (int k = 0; k < 3; ++k) coords.push_back((float)parse_f64(l, w[1 + k]));
    } else if (w[0] == "vn") {
      if (!ignore_normals) ++n_normals;
    } else if (w[0] == "vt") {
      if (!ignore_uvs) {
        if (w.size() < 3) error(l, "at least 2 components were expected, found " + std::to_string(w.size() - 1) + ".");
        uvs.push_back((float)parse_f64(l, w[1]));
        uvs.push_back((float)parse_f64(l, w[2]));
      }
    } else if (w[0] == "f") {
      auto &g = groups_ids[curr_group];
      size_t i = 0;
      for (size_t wi = 1; wi < w.size(); ++wi) {
        int64_t ids[3] = {MAXI, MAXI, MAXI};
        std::stringstream ws(w[wi]);
        std::string part;
        for (int k = 0; k < 3 && std::getline(ws, part, '/'); ++k)
          if (k == 0 || !part.empty()) {
            char *end = nullptr;
            long v = std::strtol(part.c_str(), &end, 10);
            if (end == part.c_str() || *end) error(l, "failed to parse `" + part + "' as a i32");
            ids[k] = v - 1;
          }
        if (i > 2) {  // on-the-fly fan triangulation, pivot g[len - i] (src/obj.rs:232-239)
          P3 p1 = g[g.size() - i], p2 = g[g.size() - 1];
          g.push_back(p1);
          g.push_back(p2);
        }
        if (ids[1] == MAXI) ignore_uvs = true;
        if (ids[2] == MAXI) ignore_normals = true;
        int64_t x = ids[0] < 0 ? (int64_t)(coords.size() / 3) + ids[0] + 1 : ids[0];
        int64_t y = ids[1] < 0 ? (int64_t)(uvs.size() / 2) + ids[1] + 1 : ids[1];
        int64_t z = ids[2] < 0 ? (int64_t)n_normals + ids[2] + 1 : ids[2];
        g.push_back(P3{x, y, z});
        ++i;
      }
      if (i < 2 && !g.empty())
        for (size_t k = 0; k < 3 - i; ++k) g.push_back(g.back());
    } else if (w[0] == "g") {
      curr_group = parse_g(w, 1, basename);
      if (has_curr_mtl) group2mtl[curr_group] = curr_mtl;
    } else if (w[0] == "mtllib") {
      std::ifstream f(mtl_base_dir + "/" + join(w, 1));
      if (f) {
        std::stringstream ss;
        ss << f.rdbuf();
        for (auto &m : parse_mtl(ss.str())) mtllib[m.name] = m;
      }  // missing file: the reference only warns (src/obj.rs:183)
    } else if (w[0] == "usemtl") {
      std::string mname = join(w, 1);
      if (mname != "None") {
        auto it = mtllib.find(mname);
        if (it == mtllib.end()) {
          has_curr_mtl = false;
        } else if (!group2mtl.count(curr_group)) {
          group2mtl[curr_group] = it->second;
          curr_mtl = it->second, has_curr_mtl = true;
        } else {  // several usemtl in one group: auto-generated group (src/obj.rs:149-160)
          auto gw = split_words(std::to_string(curr_group) + mname);
          curr_group = parse_g(gw, 0, "auto_generated_group_");
          group2mtl[curr_group] = it->second;
          curr_mtl = it->second, has_curr_mtl = true;
        }
      } else {
        has_curr_mtl = false;
      }
    }
  }
  // reformat (src/obj.rs:327-397); groups in first-appearance order (the reference: HashMap order)
  ObjData out;
  out.mesh = std::make_shared<MeshBuffers>();
  std::map<P3, uint32_t> vt2id;
  bool keep_uv = !ignore_uvs;
  for (size_t gi = 0; gi < groups_ids.size(); ++gi) {
    ObjGroup og;
    og.name = group_names[gi];
    for (const P3 &p : groups_ids[gi]) {
      auto it = vt2id.find(p);
      uint32_t idx;
      if (it == vt2id.end()) {
        idx = (uint32_t)(out.mesh->coords.size() / 3);
        if (p.x < 0 || (size_t)p.x >= coords.size() / 3) throw std::runtime_error("face references a missing vertex");
        for (int k = 0; k < 3; ++k) out.mesh->coords.push_back(coords[3 * p.x + k]);
        if (keep_uv) {
          if (p.y < 0 || (size_t)p.y >= uvs.size() / 2) throw std::runtime_error("face references a missing texture coordinate");
          out.mesh->uvs.push_back(uvs[2 * p.y]);
          out.mesh->uvs.push_back(uvs[2 * p.y + 1]);
        }
        vt2id[p] = idx;
      } else {
        idx = it->second;
      }
      og.faces.push_back(idx);
    }
    if (og.faces.size() % 3 != 0) throw std::runtime_error("assertion failed: vertex_ids.len() % 3 == 0");  // src/obj.rs:370
    auto m = group2mtl.find(gi);
    if (m != group2mtl.end()) og.has_mtl = true, og.mtl = m->second;
    if (!og.faces.empty()) out.groups.push_back(std::move(og));
  }
  if (!keep_uv) out.mesh->uvs.assign(out.mesh->coords.size() / 3 * 2, 0.0f);
  return out;
}

// ---- examples/loader3d.rs: scene file parser ----------------------------------------------------------------
struct Camera {
  Vec3 eye, at;
  double fovy = 0, res_x = 0, res_y = 0, aa_n = 1, aa_w = 0;
  std::string output;
};

struct Properties {
  size_t superbloc = 0;
  std::vector<std::pair<std::string, std::vector<std::string>>> geom;
  std::map<std::string, std::vector<double>> num;
  std::map<std::string, std::string> str;
  bool solid = false;
};

typedef std::map<std::string, std::pair<float, std::shared_ptr<Material>>> MtlLib;

void need(const Properties &p, bool present, const std::string &what) {
  if (!present) error(p.superbloc, "missing attribute: " + what);
}

void register_geometry(const Properties &p, MtlLib &mtllib, std::vector<std::shared_ptr<SceneNode>> &nodes) {  // :505-792
  need(p, p.num.count("pos"), "pos <x> <y> <z>");
  need(p, p.num.count("angle"), "color <r> <g> <b>");  // sic
  need(p, !p.geom.empty(), "<geom_type> <geom parameters>]");
  need(p, p.str.count("material"), "material <material_name>");
  const std::string &mname = p.str.at("material");
  bool special = mname == "uvs" || mname == "normals";
  auto mit = mtllib.find(mname);
  if (mit == mtllib.end()) throw std::runtime_error("Attempted to use an unknown material: " + mname);
  float alpha = mit->second.first;
  std::shared_ptr<Material> material = mit->second.second;
  const double pi = 3.14159265358979323846;
  const auto &pos = p.num.at("pos"), &ang = p.num.at("angle");
  Isometry3 transform = Isometry3::new_(Vec3(pos[0], pos[1], pos[2]), Vec3(ang[0] * pi / 180.0, ang[1] * pi / 180.0, ang[2] * pi / 180.0));
  float refl_m = 0, refl_a = 0;
  if (p.num.count("refl")) refl_m = (float)p.num.at("refl")[0], refl_a = (float)p.num.at("refl")[1];
  double refr_c = p.num.count("refr") ? p.num.at("refr")[0] : 1.0;
  auto push = [&](Shape g, std::shared_ptr<Material> m, float a) {
    nodes.push_back(std::make_shared<SceneNode>(m, refl_m, refl_a, a, refr_c, transform, std::move(g), nullptr, p.solid));
  };
  const auto &kind = p.geom[0].first;  // only the first shape of a block is used (F11, :593)
  const auto &w = p.geom[0].second;
  auto num = [&](size_t i) { return parse_f64(p.superbloc, w.at(i)); };
  if (kind == "ball") push(Ball(num(0)), material, alpha);
  else if (kind == "box") push(Cuboid(Vec3(num(0), num(1), num(2))), material, alpha);
  else if (kind == "cylinder") push(Cylinder(num(0), num(1)), material, alpha);
  else if (kind == "capsule") push(Capsule(num(0), num(1)), material, alpha);
  else if (kind == "cone") push(Cone(num(0), num(1)), material, alpha);
  else if (kind == "plane") push(Plane(Vec3(num(0), num(1), num(2))), material, alpha);
  else if (kind == "obj") {
    if (w.size() < 2) error(p.superbloc, "2 paths were expected, found " + std::to_string(w.size()) + ".");
    ObjData od = parse_obj(read_file(w[0]), w[1], "");
    for (float &c : od.mesh->coords) c = (float)((double)c / 4.0);  // loader3d.rs:665-670
    std::shared_ptr<const MeshBuffers> mesh = od.mesh;
    for (auto &g : od.groups) {
      Shape tm = TriMesh(mesh, g.faces);
      if (g.has_mtl) push(std::move(tm), special ? material : phong_from_mtl(g.mtl, w[1]), g.mtl.alpha * alpha);
      else push(std::move(tm), material, alpha);
    }
  }
}

void parse_scene(const std::string &text, std::vector<Light> &lights, std::vector<std::shared_ptr<SceneNode>> &nodes,
                 std::vector<Camera> &cameras) {  // :214-346
  const float ka[3] = {0.1f, 0.1f, 0.1f}, one[3] = {1, 1, 1};
  MtlLib mtllib;
  mtllib["normals"] = {1.0f, std::make_shared<NormalMaterial>()};
  mtllib["uvs"] = {1.0f, std::make_shared<UVMaterial>()};
  mtllib["default"] = {1.0f, std::make_shared<PhongMaterial>(ka, one, one, nullptr, nullptr, 100.0f)};
  std::string mode;
  Properties props;
  auto reg = [&]() {
    if (mode == "light") {  // :438-459
      need(props, props.num.count("pos"), "pos <x> <y> <z>");
      need(props, props.num.count("color"), "color <r> <g> <b>");
      double radius = props.num.count("radius") ? props.num["radius"][0] : 0.0;
      double nsample = props.num.count("nsample") ? props.num["nsample"][0] : 1.0;
      const auto &p = props.num["pos"], &c = props.num["color"];
      float col[3] = {(float)c[0], (float)c[1], (float)c[2]};
      lights.emplace_back(Vec3(p[0], p[1], p[2]), radius, (size_t)nsample, col);
    } else if (mode == "geometry") {
      register_geometry(props, mtllib, nodes);
    } else if (mode == "camera") {  // :408-436
      need(props, props.str.count("output"), "output <filename>");
      need(props, props.num.count("resolution"), "resolution <x> <y>");
      need(props, props.num.count("eye"), "eye <x> <y> <z>");
      need(props, props.num.count("at"), "at <x> <y> <z>");
      need(props, props.num.count("fovy"), "fovy <value>");
      Camera c;
      const auto &e = props.num["eye"], &a = props.num["at"], &r = props.num["resolution"];
      c.eye = Vec3(e[0], e[1], e[2]), c.at = Vec3(a[0], a[1], a[2]);
      c.fovy = props.num["fovy"][0], c.res_x = r[0], c.res_y = r[1];
      if (props.num.count("aa")) c.aa_n = props.num["aa"][0], c.aa_w = props.num["aa"][1];
      if (!(c.aa_n >= 1.0)) throw std::runtime_error("The number of ray per pixel must be at least 1.0");  // :146-149
      c.output = props.str["output"];
      cameras.push_back(c);
    }
  };
  std::istringstream is(text);
  std::string line;
  size_t l = 0;
  for (; std::getline(is, line); ++l) {
    auto w = split_words(line);
    if (w.empty() || w[0][0] == '#') continue;
    const std::string &tag = w[0];
    auto nums = [&](size_t n) {
      if (w.size() < n + 1) error(l, std::to_string(n) + " components were expected, found " + std::to_string(w.size() - 1) + ".");
      std::vector<double> v;
      for (size_t i = 0; i < n; ++i) v.push_back(parse_f64(l, w[1 + i]));
      return v;
    };
    if (tag == "mtllib") {  // register_mtllib, :461-503
      for (auto &m : parse_mtl(read_file(join(w, 1)))) mtllib[m.name] = {m.alpha, phong_from_mtl(m, "")};
    } else if (tag == "light" || tag == "geometry" || tag == "camera") {
      reg();
      props = Properties();
      props.superbloc = l;
      mode = tag;
    } else if (tag == "color" || tag == "angle" || tag == "pos" || tag == "eye" || tag == "at") props.num[tag] = nums(3);
    else if (tag == "material" || tag == "output") props.str[tag] = join(w, 1);
    else if (tag == "fovy" || tag == "refr" || tag == "radius" || tag == "nsample") props.num[tag] = nums(1);
    else if (tag == "resolution" || tag == "refl" || tag == "aa") props.num[tag] = nums(2);
    else if (tag == "ball" || tag == "plane" || tag == "box" || tag == "cylinder" || tag == "capsule" || tag == "cone" || tag == "obj")
      props.geom.emplace_back(tag, std::vector<std::string>(w.begin() + 1, w.end()));
    else if (tag == "solid") props.solid = true;
    else std::printf("Warning: unknown line %zu ignored: `%s'\n", l, line.c_str());
  }
  reg();
}

}  // namespace

int main(int argc, char **argv) {  // examples/loader3d.rs:34-101
  try {
    if (argc < 2) throw std::runtime_error(std::string("Usage: ") + argv[0] + " scene_file [--resolution W H] [--aa SPP WINDOW] [--seed S]");
    double ow = 0, oh = 0, oaa_n = 0, oaa_w = -1;
    uint64_t seed = 0;
    bool validate_only = false;
    for (int i = 2; i < argc; ++i) {
      std::string a = argv[i];
      if (a == "--resolution" && i + 2 < argc) ow = std::atof(argv[++i]), oh = std::atof(argv[++i]);
      else if (a == "--aa" && i + 2 < argc) oaa_n = std::atof(argv[++i]), oaa_w = std::atof(argv[++i]);
      else if (a == "--seed" && i + 1 < argc) seed = std::strtoull(argv[++i], nullptr, 10);
      else if (a == "--validate") validate_only = true;
      else throw std::runtime_error("unknown option " + a);
    }
    std::printf("Loading the scene.\n");
    std::vector<Light> lights;
    std::vector<std::shared_ptr<SceneNode>> nodes;
    std::vector<Camera> cameras;
    parse_scene(read_file(argv[1]), lights, nodes, cameras);
    size_t nnodes = nodes.size(), nlights = lights.size();
    const float white[3] = {1.0f, 1.0f, 1.0f};
    Scene scene(std::move(nodes), std::move(lights), white, 0, !validate_only);  // Scene::new(nodes, lights, (1,1,1)) :61
    if (validate_only) {
      NrbBuildInfo info;
      check(nrb_scene_validate(&scene.desc(), &info), "nrb_scene_validate");
      const NrbSceneDesc &d = scene.desc();
      std::printf("VALID nodes=%u lights=%u materials=%u textures=%u texels=%llu vertices=%llu indices=%llu bvh_nodes=%llu triangles=%llu "
                  "shapes=%llu planes=%llu candidates=%llu depth=%u\n",
                  d.n_nodes, d.n_lights, d.n_materials, d.n_textures, (unsigned long long)d.n_texels, (unsigned long long)d.n_vertices,
                  (unsigned long long)d.n_indices, (unsigned long long)info.bvh_nodes, (unsigned long long)info.triangles,
                  (unsigned long long)info.shapes, (unsigned long long)info.planes, (unsigned long long)info.transparent_candidates,
                  info.max_depth);
      return 0;
    }
    std::printf("Scene loaded. %zu lights, %zu objects, %zu cameras.\n", nlights, nnodes, cameras.size());
    for (auto &c : cameras) {
      if (ow > 0) c.res_x = ow, c.res_y = oh;
      if (oaa_n > 0) c.aa_n = oaa_n, c.aa_w = oaa_w;
      Mat4 projection = camera_projection(c.eye, c.at, c.fovy, c.res_x, c.res_y);
      std::printf("Casting %zu rays per pixels (win. %g).\n", (size_t)c.aa_n, c.aa_w);
      NrbStats st;
      Image img = render(scene, (uint32_t)c.res_x, (uint32_t)c.res_y, (size_t)c.aa_n, c.aa_w, c.eye, projection, seed, &st);
      std::printf("Rays cast. (%llu primary, %llu reflect, %llu refract, %llu shadow; %.3f ms on device)\n",
                  (unsigned long long)st.rays_primary, (unsigned long long)st.rays_reflect, (unsigned long long)st.rays_refract,
                  (unsigned long long)st.rays_shadow, st.ms_device);
      std::printf("Saving image to: %s\n", c.output.c_str());
      auto rgb = img.to_rgb8();
      if (!png::save_rgb8(c.output, rgb.data(), img.width, img.height)) throw std::runtime_error("Failed to save the output image.");
      std::printf("Image saved.\n");
    }
    return 0;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "loader3d: %s\n", e.what());  // the reference panics here
    return 101;                                         // Rust's panic exit code
  }
}
