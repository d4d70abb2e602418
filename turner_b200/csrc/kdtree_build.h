// Host-side kd-tree construction for the B200 path tracer.
//
// Produces, node for node, the tree the reference's KDTree(Triangles) constructor
// builds (lib/kdtree.cpp:124-490: SAH sweep of Wald-Havran Alg. 4 with clipped
// triangle boxes, then the DFS flatten), so that Stats' "Kd-Tree Height", the
// leaf visiting order and therefore exact-tie resolution are identical -- but as
// a task-parallel build over allocation-free clipping instead of the reference's
// single-threaded recursion over std::vector<Point3f>.
#pragma once
#include <cstdint>
#include <vector>

namespace trn {

struct HostTriangles {
    // 64-byte intersection record per triangle, the reference's precomputed quantities
    // (lib/triangle.h:33-39): v0.xyz n.xyz u.xyz v.xyz uv vv uu denom
    std::vector<float> isect; // n*16
    // shading record: n0.xyz n1.xyz n2.xyz pad*3 rgba (lib/triangle.h:95-98)
    std::vector<float> shade; // n*16
    std::vector<float> verts; // n*9, as given
    std::vector<float> mirror; // n*4 reflective rgba (raytracer only); reflectivity sits in shade[9]
    uint32_t count = 0;
};

// fills isect/shade from raw arrays (verts n*9, normals n*9, diffuse n*4)
void precompute_triangles(const float* verts, const float* normals, const float* diffuse, uint32_t n, HostTriangles& out,
                          const float* reflective = nullptr, const float* reflectivity = nullptr);

struct KdTree {
    float box[6]; // min xyz, max xyz (KDTree::box())
    // the reference's FlatNode encoding (lib/kdtree.h:62-154), DFS order, left child = next node
    std::vector<uint64_t> nodes;
    uint64_t height = 0;
    uint64_t num_leaf_refs = 0;
    double build_ms = 0;

    // Device layout of the SAME tree plus the empty-space cuts the reference throws away
    // (lib/kdtree.cpp:168-172 returns the non-empty child when the other side of the chosen plane holds no
    // triangle, so its traversal walks the surviving subtree even where the ray only crosses the cut-off
    // void). Nodes are 8 bytes, siblings adjacent and 16-byte aligned so one 16-byte load fetches both
    // children:   inner: x = split bits,       y = (index of the child pair << 2) | axis
    //             leaf:  x = first reference,  y = (count << 2) | 3        (count 0 = cut-off void)
    // Node 0 is the root (node 1 pads the pair). Leaves are visited in the reference's order.
    // The first kTreeletNodes entries are the top of the tree in breadth-first order (the "top treelet" a kernel can
    // stage in shared memory as one contiguous block); below that each subtree is laid out depth-first.
    static constexpr uint32_t kTreeletNodes = 2048;
    std::vector<uint64_t> pair_nodes;      // low 32 bits = x, high 32 bits = y
    std::vector<uint32_t> pair_leaf_refs;  // triangle ids, leaf after leaf; every run starts at a multiple of 4 (padded)
    uint64_t num_pair_refs = 0;            // references without the padding
    uint64_t num_cut_nodes = 0;
    uint64_t expected_nodes = 0;           // device builder: size `nodes` will have once reference_shape_from_pairs() ran
};

// num_threads <= 0: hardware concurrency
void build_kdtree(const HostTriangles& tris, KdTree& out, int num_threads = 0);

// Derives the reference-shaped array `nodes` (FlatNode encoding, DFS, cut nodes dropped as lib/kdtree.cpp:168-172 drops
// them) and `height` from the sibling-pair layout. The host builder fills both itself; the device builder
// (kdtree_build_gpu.h) only makes the pair layout, and this runs when somebody asks for the reference's view of the tree.
void reference_shape_from_pairs(KdTree& tree);

} // namespace trn
