// nrays.hpp — C++ host mirror of the reference crate's scene API, above the C-ABI of include/nrays_b200.h.
//
// The reference host is Rust (no rustc/cargo in this image); this header keeps the same type names,
// constructor argument order and error behaviour (a Rust panic becomes a C++ exception) so that code
// reads like the reference's own:
//     Light::new(pos, radius, nsample, color)                         src/light.rs:16-23
//     PhongMaterial::new(ambiant, diffuse, specular, texture, alpha, shininess)
//                                                                     src/phong_material.rs:19-35
//     SceneNode::new(material, refl_mix, refl_atenuation, alpha, refr_coeff, transform, geometry,
//                    nmap, solid)                                      src/scene_node.rs:22-47
//     Scene::new(nodes, lights, background)                           src/scene.rs:119-133
//     render(&scene, &resolution, ray_per_pixel, window_width, camera_eye, projection) -> Image
//                                                                     src/scene.rs:29-116
// Nothing here computes pixels: render() flattens to the POD tables and calls nrb_render.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nrays_b200.h"

namespace nrays {

using Scalar = double;  // src/lib.rs:33

struct Vec3 {
  double x = 0, y = 0, z = 0;
  Vec3() {}
  Vec3(double a, double b, double c) : x(a), y(b), z(c) {}
  Vec3 operator-(const Vec3 &o) const { return Vec3(x - o.x, y - o.y, z - o.z); }
  Vec3 operator*(double s) const { return Vec3(x * s, y * s, z * s); }
  double norm() const { return std::sqrt(x * x + y * y + z * z); }
  Vec3 normalized() const {
    double n = norm();
    return Vec3(x / n, y / n, z / n);
  }
  Vec3 cross(const Vec3 &b) const { return Vec3(y * b.z - z * b.y, z * b.x - x * b.z, x * b.y - y * b.x); }
};

// nalgebra Matrix4<f64>, column-major storage.
struct Mat4 {
  double m[16];
  static Mat4 zero() {
    Mat4 r;
    std::memset(r.m, 0, sizeof(r.m));
    return r;
  }
  static Mat4 identity() {
    Mat4 r = zero();
    r.at(0, 0) = r.at(1, 1) = r.at(2, 2) = r.at(3, 3) = 1.0;
    return r;
  }
  double &at(int row, int col) { return m[col * 4 + row]; }
  double at(int row, int col) const { return m[col * 4 + row]; }
  Mat4 operator*(const Mat4 &b) const {
    Mat4 r = zero();
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j)
        for (int k = 0; k < 4; ++k) r.at(i, j) += at(i, k) * b.at(k, j);
    return r;
  }
  // try_inverse().unwrap() — examples/loader3d.rs:77-79 (Gauss-Jordan with partial pivoting)
  Mat4 inverse() const {
    double a[4][8];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) a[i][j] = at(i, j), a[i][4 + j] = (i == j) ? 1.0 : 0.0;
    for (int c = 0; c < 4; ++c) {
      int p = c;
      for (int r = c + 1; r < 4; ++r)
        if (std::fabs(a[r][c]) > std::fabs(a[p][c])) p = r;
      if (a[p][c] == 0.0) throw std::runtime_error("called `Option::unwrap()` on a `None` value: matrix is not invertible");
      if (p != c)
        for (int j = 0; j < 8; ++j) std::swap(a[p][j], a[c][j]);
      double inv = 1.0 / a[c][c];
      for (int j = 0; j < 8; ++j) a[c][j] *= inv;
      for (int r = 0; r < 4; ++r)
        if (r != c) {
          double f = a[r][c];
          for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j];
        }
    }
    Mat4 r;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) r.at(i, j) = a[i][4 + j];
    return r;
  }
};

// nalgebra Isometry3<f64>: rotation (kept as a row-major 3x3 matrix) + translation.
struct Isometry3 {
  double rot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double trans[3] = {0, 0, 0};
  static Isometry3 identity() { return Isometry3(); }
  // Isometry3::new(translation, axisangle): rotation = exp map of the axis-angle vector (loader3d.rs:546-552)
  static Isometry3 new_(const Vec3 &t, const Vec3 &axisangle) {
    Isometry3 r;
    r.trans[0] = t.x, r.trans[1] = t.y, r.trans[2] = t.z;
    double ang = axisangle.norm();
    if (ang != 0.0) {
      Vec3 k = axisangle * (1.0 / ang);
      double K[9] = {0, -k.z, k.y, k.z, 0, -k.x, -k.y, k.x, 0}, K2[9];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          K2[3 * i + j] = 0;
          for (int l = 0; l < 3; ++l) K2[3 * i + j] += K[3 * i + l] * K[3 * l + j];
        }
      double s = std::sin(ang), c = 1.0 - std::cos(ang);
      for (int i = 0; i < 9; ++i) r.rot[i] = ((i % 4 == 0) ? 1.0 : 0.0) + s * K[i] + c * K2[i];
    }
    return r;
  }
  // Isometry3::look_at_rh(eye, target, up).to_homogeneous() — SURVEY B.9
  static Mat4 look_at_rh(const Vec3 &eye, const Vec3 &target, const Vec3 &up) {
    Vec3 f = (target - eye).normalized();
    Vec3 r = f.cross(up).normalized();
    Vec3 u = r.cross(f);
    Mat4 M = Mat4::identity();
    double R[3][3] = {{r.x, r.y, r.z}, {u.x, u.y, u.z}, {-f.x, -f.y, -f.z}};
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) M.at(i, j) = R[i][j];
      M.at(i, 3) = -(R[i][0] * eye.x + R[i][1] * eye.y + R[i][2] * eye.z);
    }
    return M;
  }
};

// Perspective3::new(aspect, fovy, znear, zfar).to_homogeneous() — SURVEY B.9
inline Mat4 perspective3(double aspect, double fovy, double znear, double zfar) {
  double t = std::tan(fovy / 2.0);
  Mat4 P = Mat4::zero();
  P.at(0, 0) = 1.0 / (aspect * t);
  P.at(1, 1) = 1.0 / t;
  P.at(2, 2) = (zfar + znear) / (znear - zfar);
  P.at(2, 3) = 2.0 * zfar * znear / (znear - zfar);
  P.at(3, 2) = -1.0;
  return P;
}

// (perspective * view)^-1 — examples/loader3d.rs:68-79
inline Mat4 camera_projection(const Vec3 &eye, const Vec3 &at, double fovy_deg, double width, double height) {
  const double pi = 3.14159265358979323846;
  Mat4 persp = perspective3(width / height, fovy_deg * pi / 180.0, 1.0, 100000.0);
  return (persp * Isometry3::look_at_rh(eye, at, Vec3(0, 1, 0))).inverse();
}

// ---- shapes (ncollide3d::shape::*) ------------------------------------------------------------------
struct MeshBuffers {  // vertex pool shared by the groups of one OBJ file (loader3d.rs:690-695)
  std::vector<float> coords;  // 3 per vertex
  std::vector<float> uvs;     // 2 per vertex (zero filled when the file has none, src/obj.rs:383)
};
struct Shape {
  int kind = NRB_SHAPE_BALL;
  double param[3] = {0, 0, 0};
  std::shared_ptr<const MeshBuffers> mesh;  // TRIMESH only
  std::vector<uint32_t> faces;              // 3 per triangle
};
inline Shape Ball(double radius) {
  Shape s;
  s.kind = NRB_SHAPE_BALL, s.param[0] = radius;
  return s;
}
inline Shape Cuboid(const Vec3 &he) {
  Shape s;
  s.kind = NRB_SHAPE_CUBOID, s.param[0] = he.x, s.param[1] = he.y, s.param[2] = he.z;
  return s;
}
inline Shape Cylinder(double hh, double r) {
  Shape s;
  s.kind = NRB_SHAPE_CYLINDER, s.param[0] = hh, s.param[1] = r;
  return s;
}
inline Shape Capsule(double hh, double r) {
  Shape s;
  s.kind = NRB_SHAPE_CAPSULE, s.param[0] = hh, s.param[1] = r;
  return s;
}
inline Shape Cone(double hh, double r) {
  Shape s;
  s.kind = NRB_SHAPE_CONE, s.param[0] = hh, s.param[1] = r;
  return s;
}
inline Shape Plane(const Vec3 &normal) {  // Unit::new_normalize — loader3d.rs:656
  Shape s;
  Vec3 n = normal.normalized();
  s.kind = NRB_SHAPE_PLANE, s.param[0] = n.x, s.param[1] = n.y, s.param[2] = n.z;
  return s;
}
inline Shape TriMesh(std::shared_ptr<const MeshBuffers> mesh, std::vector<uint32_t> faces) {
  Shape s;
  s.kind = NRB_SHAPE_TRIMESH;
  size_t nv = mesh->coords.size() / 3;
  for (uint32_t f : faces)
    if (f >= nv) throw std::runtime_error("TriMesh face index out of range");
  s.mesh = std::move(mesh);
  s.faces = std::move(faces);
  return s;
}

// ---- textures ------------------------------------------------------------------------------------------
enum class Interpolation { Bilinear = NRB_INTERP_BILINEAR, Nearest = NRB_INTERP_NEAREST };
enum class Overflow { ClampToEdges = NRB_OVERFLOW_CLAMP, Wrap = NRB_OVERFLOW_WRAP };

struct ImageData {  // src/texture2d.rs:10-26
  std::vector<float> pixels;  // RGBA32F, row-major y*W+x
  size_t w, h;
  ImageData(std::vector<float> px, size_t w_, size_t h_) : pixels(std::move(px)), w(w_), h(h_) {
    if (pixels.size() != 4 * w * h || w < 1 || h < 1) throw std::runtime_error("assertion failed: pixels.len() == dims.x * dims.y");
  }
};
struct Texture2d {
  std::shared_ptr<const ImageData> data;
  Interpolation interpol;
  Overflow overflow;
  // Channel expansion of Texture2d::from_png after decoding (src/texture2d.rs:96-173): y flip, then by depth.
  static Texture2d from_pixels(const uint8_t *px, size_t w, size_t h, int depth, bool opacity, Interpolation ip, Overflow ov) {
    std::vector<float> out(4 * w * h, 1.0f);
    for (size_t y = 0; y < h; ++y)
      for (size_t x = 0; x < w; ++x) {
        const uint8_t *p = px + ((h - 1 - y) * w + x) * depth;  // flipped
        float *o = &out[4 * (y * w + x)];
        float c[4] = {0, 0, 0, 0};
        for (int k = 0; k < depth; ++k) c[k] = (float)p[k] / 255.0f;
        if (depth == 1) {
          if (opacity) o[3] = c[0];
          else o[0] = o[1] = o[2] = c[0];
        } else if (depth == 2) {
          if (opacity) o[3] = c[1] * c[0];
          else o[0] = o[1] = o[2] = c[0] * c[1];
        } else if (depth == 3) {
          if (opacity) o[3] = c[0];  // red channel (texture2d.rs:140)
          else o[0] = c[0], o[1] = c[1], o[2] = c[2];
        } else if (depth == 4) {
          if (opacity) o[3] = c[3];
          else o[0] = c[0], o[1] = c[1], o[2] = c[2];  // alpha dropped (texture2d.rs:163)
        } else {
          throw std::runtime_error("Image depth " + std::to_string(depth) + " not suported.");
        }
      }
    return Texture2d{std::make_shared<ImageData>(std::move(out), w, h), ip, ov};
  }
};

// ---- materials (trait Material, src/material.rs:6-17) -----------------------------------------------------
struct Material {
  virtual ~Material() {}
  virtual int kind() const = 0;
};
struct PhongMaterial : Material {
  float ambiant_color[3], diffuse_color[3], specular_color[3];
  std::shared_ptr<Texture2d> texture, alpha;
  float shininess;
  PhongMaterial(const float a[3], const float d[3], const float s[3], std::shared_ptr<Texture2d> tex,
                std::shared_ptr<Texture2d> al, float sh)
      : texture(std::move(tex)), alpha(std::move(al)), shininess(sh) {
    for (int k = 0; k < 3; ++k) ambiant_color[k] = a[k], diffuse_color[k] = d[k], specular_color[k] = s[k];
  }
  int kind() const override { return NRB_MAT_PHONG; }
};
struct NormalMaterial : Material {
  int kind() const override { return NRB_MAT_NORMAL; }
};
struct UVMaterial : Material {
  int kind() const override { return NRB_MAT_UV; }
};

struct Light {  // src/light.rs:8-23
  Vec3 pos;
  double radius;
  size_t racsample;
  float color[3];
  Light(const Vec3 &p, double r, size_t nsample, const float c[3]) : pos(p), radius(r) {
    racsample = (size_t)std::sqrt((float)nsample);  // ((nsample as f32).sqrt()) as usize
    for (int k = 0; k < 3; ++k) color[k] = c[k];
  }
};

struct SceneNode {  // src/scene_node.rs:8-47
  std::shared_ptr<Material> material;
  float refl_mix, refl_atenuation, alpha;
  double refr_coeff;
  Isometry3 transform;
  Shape geometry;
  std::shared_ptr<Texture2d> nmap;
  bool solid;
  SceneNode(std::shared_ptr<Material> m, float refl_mix_, float refl_atenuation_, float alpha_, double refr_coeff_,
            const Isometry3 &t, Shape g, std::shared_ptr<Texture2d> nmap_ = nullptr, bool solid_ = false)
      : material(std::move(m)), refl_mix(refl_mix_), refl_atenuation(refl_atenuation_), alpha(alpha_),
        refr_coeff(refr_coeff_), transform(t), geometry(std::move(g)), nmap(std::move(nmap_)), solid(solid_) {}
};

inline void check(int status, const char *what) {
  if (status != NRB_OK) throw std::runtime_error(std::string(what) + ": " + nrb_last_error());
}

struct Image {  // src/image.rs:12-24
  uint32_t width = 0, height = 0;
  std::vector<float> pixels;  // Vec<Vector3<f32>>, row-major
  // Image::to_png's quantisation (src/image.rs:64-77)
  std::vector<uint8_t> to_rgb8() const {
    std::vector<uint8_t> out(pixels.size());
    for (size_t i = 0; i < pixels.size(); ++i) {
      float v = pixels[i] * 255.0f;
      v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);
      out[i] = (uint8_t)(unsigned)v;
    }
    return out;
  }
};

class Scene {  // src/scene.rs:21-25
 public:
  // Scene::new(nodes, lights, background) — flattens and uploads (the BVH is built by the library)
  // `upload = false` stops after flattening (host-only validation through nrb_scene_validate).
  Scene(std::vector<std::shared_ptr<SceneNode>> nodes, std::vector<Light> lights, const float background[3], int device = 0,
        bool upload = true)
      : nodes_(std::move(nodes)), lights_(std::move(lights)) {
    for (int k = 0; k < 3; ++k) background_[k] = background[k];
    flatten();
    if (upload) check(nrb_scene_create(&desc_, device, &handle_), "nrb_scene_create");
  }
  ~Scene() {
    if (handle_) nrb_scene_destroy(handle_);
  }
  Scene(const Scene &) = delete;
  Scene &operator=(const Scene &) = delete;
  const std::vector<Light> &lights() const { return lights_; }
  void set_background(const float bg[3]) {
    for (int k = 0; k < 3; ++k) background_[k] = bg[k];
    check(nrb_scene_set_background(handle_, background_), "nrb_scene_set_background");
  }
  NrbScene *handle() const { return handle_; }
  const NrbSceneDesc &desc() const { return desc_; }

 private:
  void flatten() {
    std::map<const ImageData *, uint64_t> data_off;
    std::map<std::tuple<const ImageData *, int, int>, int> tex_index;
    auto tex_id = [&](const std::shared_ptr<Texture2d> &t) -> int {
      if (!t) return -1;
      auto key = std::make_tuple(t->data.get(), (int)t->interpol, (int)t->overflow);
      auto it = tex_index.find(key);
      if (it != tex_index.end()) return it->second;
      if (!data_off.count(t->data.get())) {
        data_off[t->data.get()] = texels_.size() / 4;
        texels_.insert(texels_.end(), t->data->pixels.begin(), t->data->pixels.end());
      }
      NrbTextureDesc d{(uint32_t)t->data->w, (uint32_t)t->data->h, (int)t->interpol, (int)t->overflow, data_off[t->data.get()]};
      textures_.push_back(d);
      return tex_index[key] = (int)textures_.size() - 1;
    };
    std::map<const Material *, int> mat_index;
    auto mat_id = [&](const std::shared_ptr<Material> &m) -> int {
      auto it = mat_index.find(m.get());
      if (it != mat_index.end()) return it->second;
      NrbMaterialDesc d;
      std::memset(&d, 0, sizeof(d));
      d.kind = m->kind(), d.texture = -1, d.alpha_texture = -1;
      if (auto p = dynamic_cast<const PhongMaterial *>(m.get())) {
        for (int k = 0; k < 3; ++k) d.ambient[k] = p->ambiant_color[k], d.diffuse[k] = p->diffuse_color[k], d.specular[k] = p->specular_color[k];
        d.shininess = p->shininess;
        d.texture = tex_id(p->texture);
        d.alpha_texture = tex_id(p->alpha);
      }
      materials_.push_back(d);
      return mat_index[m.get()] = (int)materials_.size() - 1;
    };
    std::map<const MeshBuffers *, uint64_t> vbase;
    for (auto &n : nodes_) {
      NrbNodeDesc r;
      std::memset(&r, 0, sizeof(r));
      r.shape = n->geometry.kind;
      r.material = mat_id(n->material);
      for (int k = 0; k < 3; ++k) r.param[k] = n->geometry.param[k], r.trans[k] = n->transform.trans[k];
      for (int k = 0; k < 9; ++k) r.rot[k] = n->transform.rot[k];
      r.refr_coeff = n->refr_coeff;
      r.refl_mix = n->refl_mix, r.refl_atenuation = n->refl_atenuation, r.alpha = n->alpha;
      r.solid = n->solid ? 1 : 0;
      r.nmap_texture = tex_id(n->nmap);
      if (n->geometry.kind == NRB_SHAPE_TRIMESH) {
        const MeshBuffers *mb = n->geometry.mesh.get();
        if (!vbase.count(mb)) {
          vbase[mb] = positions_.size() / 3;
          positions_.insert(positions_.end(), mb->coords.begin(), mb->coords.end());
          if (mb->uvs.size() == mb->coords.size() / 3 * 2)
            uvs_.insert(uvs_.end(), mb->uvs.begin(), mb->uvs.end());
          else
            uvs_.insert(uvs_.end(), mb->coords.size() / 3 * 2, 0.0f);
        }
        r.vertex_base = vbase[mb];
        r.first_index = indices_.size();
        r.tri_count = n->geometry.faces.size() / 3;
        indices_.insert(indices_.end(), n->geometry.faces.begin(), n->geometry.faces.end());
      }
      node_rows_.push_back(r);
    }
    for (auto &l : lights_) {
      NrbLightDesc d;
      std::memset(&d, 0, sizeof(d));
      d.pos[0] = l.pos.x, d.pos[1] = l.pos.y, d.pos[2] = l.pos.z;
      d.radius = l.radius;
      d.racsample = (uint32_t)l.racsample;
      for (int k = 0; k < 3; ++k) d.color[k] = l.color[k];
      light_rows_.push_back(d);
    }
    std::memset(&desc_, 0, sizeof(desc_));
    desc_.struct_size = sizeof(NrbSceneDesc), desc_.abi_version = NRB_ABI_VERSION;
    desc_.n_nodes = (uint32_t)node_rows_.size(), desc_.nodes = node_rows_.data();
    desc_.n_lights = (uint32_t)light_rows_.size(), desc_.lights = light_rows_.data();
    desc_.n_materials = (uint32_t)materials_.size(), desc_.materials = materials_.data();
    desc_.n_textures = (uint32_t)textures_.size(), desc_.textures = textures_.data();
    desc_.n_texels = texels_.size() / 4, desc_.texels = texels_.data();
    desc_.n_vertices = positions_.size() / 3, desc_.positions = positions_.data();
    desc_.uvs = uvs_.empty() ? nullptr : uvs_.data();
    desc_.n_indices = indices_.size(), desc_.indices = indices_.data();
    for (int k = 0; k < 3; ++k) desc_.background[k] = background_[k];
  }

  std::vector<std::shared_ptr<SceneNode>> nodes_;
  std::vector<Light> lights_;
  float background_[3];
  std::vector<NrbNodeDesc> node_rows_;
  std::vector<NrbLightDesc> light_rows_;
  std::vector<NrbMaterialDesc> materials_;
  std::vector<NrbTextureDesc> textures_;
  std::vector<float> texels_, positions_, uvs_;
  std::vector<uint32_t> indices_;
  NrbSceneDesc desc_;
  NrbScene *handle_ = nullptr;
};

// scene::render — src/scene.rs:29-116
inline Image render(const Scene &scene, uint32_t res_x, uint32_t res_y, size_t ray_per_pixel, Scalar window_width,
                    const Vec3 &camera_eye, const Mat4 &projection, uint64_t seed = 0, NrbStats *stats = nullptr) {
  if (ray_per_pixel == 0) throw std::runtime_error("assertion failed: ray_per_pixel > 0");  // src/scene.rs:37
  NrbCamera cam;
  std::memset(&cam, 0, sizeof(cam));
  cam.width = res_x, cam.height = res_y, cam.ray_per_pixel = (uint32_t)ray_per_pixel;
  cam.window_width = window_width;
  cam.eye[0] = camera_eye.x, cam.eye[1] = camera_eye.y, cam.eye[2] = camera_eye.z;
  std::memcpy(cam.projection, projection.m, sizeof(cam.projection));
  cam.seed = seed;
  Image img;
  img.width = res_x, img.height = res_y;
  img.pixels.resize((size_t)res_x * res_y * 3);
  check(nrb_render(scene.handle(), &cam, img.pixels.data(), stats), "nrb_render");
  return img;
}

}  // namespace nrays
