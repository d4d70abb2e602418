// Device kd-tree builder: fills the sibling-pair device layout of KdTree (pair_nodes, pair_leaf_refs, box, height,
// counts) from the triangles; the reference-shaped node array stays empty until reference_shape_from_pairs() derives it.
#pragma once
#include <string>

#include "kdtree_build.h"

namespace trn {

// returns 0, or a negative trn_status (-2 CUDA, -5 limit) with `err` set
int build_kdtree_gpu(const HostTriangles& tris, KdTree& out, int device, std::string& err);

} // namespace trn
