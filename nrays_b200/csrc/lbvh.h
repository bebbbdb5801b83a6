// lbvh.h — device LBVH build over a set of (padded) triangle boxes.  See lbvh.cu.
#pragma once
#include <vector>

#include "bvh_build.h"

namespace nrb {

// Builds the tree on the current CUDA device.  Outputs are host vectors in this library's node layout:
// inner child codes index `nodes_out`, leaf codes address positions of `order_out` (which maps a leaf
// position to the index of the input box).  `root_code` may be a leaf code when n <= 4.
cudaError_t lbvh_build(const Box *h_boxes, uint32_t n, std::vector<BvhNode> &nodes_out, std::vector<uint32_t> &order_out,
                       int *root_code, Box *root_box, int *depth, float *gpu_ms);

// Same contract, PLOC build (parallel locally-ordered clustering, search radius `radius`): merges are chosen by merged
// surface area instead of Morton prefixes, so the tree traverses close to the host SAH tree.
cudaError_t ploc_build(const Box *h_boxes, uint32_t n, int radius, std::vector<BvhNode> &nodes_out, std::vector<uint32_t> &order_out,
                       int *root_code, Box *root_box, int *depth, float *gpu_ms);

}  // namespace nrb
