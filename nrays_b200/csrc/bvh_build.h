// bvh_build.h — host BVH builder replacing BVT::new_balanced (src/scene.rs:126) and the inner BVT of
// TriMesh::new (examples/loader3d.rs:695).  The reference's median-split tree shape does not affect
// results, so the device tree is a binned-SAH BVH2 in the 64-byte two-box node layout.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

#include "device_types.cuh"

namespace nrb {

struct Box {
  float lo[3], hi[3];
  void reset() {
    for (int i = 0; i < 3; ++i) lo[i] = 3.402823466e+38f, hi[i] = -3.402823466e+38f;
  }
  void grow(const Box &b) {
    for (int i = 0; i < 3; ++i) {
      lo[i] = b.lo[i] < lo[i] ? b.lo[i] : lo[i];
      hi[i] = b.hi[i] > hi[i] ? b.hi[i] : hi[i];
    }
  }
  void grow(const float p[3]) {
    for (int i = 0; i < 3; ++i) {
      lo[i] = p[i] < lo[i] ? p[i] : lo[i];
      hi[i] = p[i] > hi[i] ? p[i] : hi[i];
    }
  }
  float half_area() const {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
  }
  bool valid() const { return lo[0] <= hi[0]; }
};

// A build item: its box and, for "payload" builds (top level over sub-roots / analytic shapes),
// the child code to emit when the item becomes a leaf on its own.
struct BuildItem {
  Box box;
  int payload;  // child code (payload builds) or original triangle index (triangle builds)
};

struct BvhBuilder {
  std::vector<BvhNode> nodes;      // shared node pool (all builds append here)
  std::vector<uint32_t> tri_order; // leaf-ordered triangle indices (triangle builds append here)
  std::vector<char> dead;          // nodes of the pool that were replaced (tops of device-built trees, api.cu) and must not be laid out
  int max_depth_seen = 0;

  // Builds a tree over triangles; leaves hold up to kMaxLeafTris consecutive entries of tri_order.
  // Returns the child code of the root and its box.
  int build_triangles(std::vector<BuildItem> &items, Box *root_box);
  // Builds a tree over payload items (one item per leaf; a single item returns its payload).
  int build_payloads(std::vector<BuildItem> &items, Box *root_box);

 private:
  struct Task {
    BuildItem *items;
    size_t n;
    int depth;
    int placeholder;  // child code written into the parent until the subtree is merged
  };
  // known_bounds / known_cbounds: box and centroid box of the items when the parent already has them (large nodes)
  int build_rec(BuildItem *items, size_t n, bool payload, int depth, Box *out_box, const Box *known_bounds = nullptr,
                const Box *known_cbounds = nullptr);
  std::vector<Task> *tasks_ = nullptr;  // non-null while the top of a large tree is being split
  size_t task_size_ = 0;
  std::unique_ptr<BuildItem[]> scratch_;  // stable partition of large nodes
  size_t scratch_n_ = 0;
};

// widen a box by a few ulps so f32 slab tests stay conservative w.r.t. the triangle test
void pad_box(Box &b, float scene_extent);

}  // namespace nrb
