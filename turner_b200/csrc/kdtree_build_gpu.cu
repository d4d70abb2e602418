// Device kd-tree builder (replaces lib/kdtree.cpp:124-467 for scenes whose host build would
// dominate time-to-image). Level-synchronous binned SAH on the GPU:
//
//   per level, over ALL nodes of that level at once
//     bin      every triangle reference adds the first / last bin of its (clipped) box to its node's 3 x 32-bin
//              histograms, taken over the node's TIGHT box (the union of its references' boxes, carried down from the
//              parent's scatter) -- CTAs whose references all belong to one node (the top levels) accumulate in shared
//              memory first
//     select   one warp per node: prefix sums over the bins, the reference's cost function
//              lambda (15 + 20 (SA_l/SA N_l + SA_r/SA N_r)), lambda = 0.8 next to an empty side (lib/kdtree.cpp:178-203)
//              at the 33 bin boundaries of every axis (the two outer ones cut off empty space), the reference's
//              termination rules (<= 3 triangles, zero-area box, 20 N lambda < best cost; lib/kdtree.cpp:133-158)
//     apply    writes the node into the sibling-pair layout (kdtree_build.h) and opens its two children
//     count    references of new leaves go to the leaf pool; the others count into their children
//     scatter  references move to their children (both, when their box reaches over the plane); a straddling
//              triangle is clipped against the child's box (Sutherland-Hodgman) so that the child sees its exact extent
//   scans (cub) give every level's pair slots, child segments and leaf-pool offsets: the result does not depend on
//   the order in which atomics land; a final 64-bit radix sort orders the ids inside every leaf.
//
// What is NOT reproduced is the reference's exact plane choice (it sweeps all clipped-box events, this bins 32 planes
// per axis): the tree has a different shape. Closest-hit ids do not depend on the tree shape apart from
// which of several triangles reports an EXACT tie; tests/test_gpu_parity.py runs the parity suites on this tree too.
// Robustness: a reference whose box only touches the split plane (within 1e-6 of the scene scale) goes to BOTH
// children, and a plane that cuts off empty space is moved 4e-6 x scene scale into the void (as in kdtree_build.cpp).
#include "kdtree_build.h"
#include "kdtree_build_gpu.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace trn {
namespace gpubuild {

constexpr int kBins = 32;
constexpr int kMaxLevels = 56;          // tree height bound (the traversal stack holds 64 entries)
// lib/kdtree.cpp:178 has 15: on the pooled GPU traversal an inner-node step costs more against a triangle pre-test than on
// the CPU; 30 renders the 1M mesh 3 % faster than 15 (8: -6 %, 60: +2.5 %), 40 together with the weaker empty-space bonus
// below another 3 % (profiles/README.md); the tree is as correct
constexpr float kCostTraversal = 40.f;
constexpr float kCostIntersection = 20.f; // lib/kdtree.cpp:179
// lib/kdtree.cpp:183-188 has 0.8. 39 % of the walk steps of the 1M mesh were at empty-space cuts; a cut has to save more than a
// step per ray through the node to pay: 0.8 / 0.85 / 0.9 / 0.95 / 1.0 render at 1691 / 1718 / 1738 / 1728 / 1641 Mrays/s
constexpr float kLambdaEmpty = 0.9f;
constexpr uint32_t kLeafMax = 3;        // lib/kdtree.cpp:133-136

struct SahParams { // the reference's constants; TRN_KD_* override them for experiments (profiles/README.md)
    float kt, ki, lambda;
    uint32_t leaf_max;
};
constexpr float kFltMax = 3.402823466e+38f;

#define GB_TRY(expr)                                                                                     \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            err = std::string(#expr) + ": " + cudaGetErrorString(e__) + " (kdtree_build_gpu.cu:" +       \
                  std::to_string(__LINE__) + ")";                                                        \
            return -2;                                                                                   \
        }                                                                                                \
    } while (0)

// floats as order-preserving unsigned keys (atomicMin / atomicMax on boxes)
__host__ __device__ inline uint32_t f2o(float f) {
#ifdef __CUDA_ARCH__
    const uint32_t u = __float_as_uint(f);
#else
    uint32_t u;
    std::memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ inline float o2f(uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

struct Level { // one level's nodes (structure of arrays, capacity = max nodes per level)
    float* blo;       // [3 * cap] node box
    float* bhi;
    uint32_t* tlo;    // [3 * cap] tight box of the node's references, ordered keys
    uint32_t* thi;
    uint32_t* first;  // segment of the node's references in the level's reference array
    uint32_t* count;
    uint32_t* slot;   // index of the node's word in pair_nodes
};

struct Decision { // per node of the current level
    int* axis;            // 0..2 split, -1 leaf, -2 void
    float* pos;
    uint32_t* is_split;   // scan inputs / outputs
    uint32_t* split_idx;
    uint32_t* leaf_alloc;
    uint32_t* leaf_off;
    uint32_t* leaf_first; // pool offset of a leaf's run
};

struct Totals {
    uint32_t num_split, leaf_alloc, next_refs, cut_nodes, leaf_refs;
    unsigned long long leaf_ref_nodes; // FlatNodes the leaves take in the reference's encoding: floor(count / 2) + 1 each
};

__global__ void init_refs_kernel(const float* __restrict__ verts, uint32_t n, float4* __restrict__ ra, float4* __restrict__ rb) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = verts + size_t(i) * 9;
    float lo[3], hi[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        lo[c] = fminf(p[c], fminf(p[3 + c], p[6 + c]));
        hi[c] = fmaxf(p[c], fmaxf(p[3 + c], p[6 + c]));
    }
    ra[i] = make_float4(lo[0], lo[1], lo[2], __uint_as_float(i));
    rb[i] = make_float4(hi[0], hi[1], hi[2], __uint_as_float(0u));
}

// ---- bin: 6 histogram increments per reference
__global__ void __launch_bounds__(256) bin_kernel(const float4* __restrict__ ra, const float4* __restrict__ rb, uint32_t nrefs, Level lv,
                                                  uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_hist[6 * kBins];
    __shared__ uint32_t s_uniform;
    const uint32_t base = blockIdx.x * 1024u;
    const uint32_t last = min(base + 1024u, nrefs) - 1u;
    if (threadIdx.x == 0) {
        const uint32_t n0 = __float_as_uint(rb[base].w), n1 = __float_as_uint(rb[last].w);
        s_uniform = n0 == n1 ? n0 : 0xFFFFFFFFu;
    }
    for (int k = threadIdx.x; k < 6 * kBins; k += 256) s_hist[k] = 0;
    __syncthreads();
    const uint32_t uni = s_uniform;
    for (int k = 0; k < 4; ++k) {
        const uint32_t i = base + k * 256u + threadIdx.x;
        if (i >= nrefs) break;
        const float4 a = ra[i], b = rb[i];
        const uint32_t node = __float_as_uint(b.w);
        const float lo[3] = {a.x, a.y, a.z}, hi[3] = {b.x, b.y, b.z};
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
            const float tl = o2f(lv.tlo[3 * node + ax]), th = o2f(lv.thi[3 * node + ax]);
            const float ext = th - tl;
            if (!(ext > 0.f)) continue;
            const float sc = static_cast<float>(kBins) / ext;
            const int bl = min(kBins - 1, max(0, static_cast<int>((lo[ax] - tl) * sc)));
            const int bh = min(kBins - 1, max(0, static_cast<int>((hi[ax] - tl) * sc)));
            if (uni != 0xFFFFFFFFu) {
                atomicAdd(&s_hist[(2 * ax) * kBins + bl], 1u);
                atomicAdd(&s_hist[(2 * ax + 1) * kBins + bh], 1u);
            } else {
                atomicAdd(&hist[size_t(node) * 6 * kBins + (2 * ax) * kBins + bl], 1u);
                atomicAdd(&hist[size_t(node) * 6 * kBins + (2 * ax + 1) * kBins + bh], 1u);
            }
        }
    }
    if (uni != 0xFFFFFFFFu) {
        __syncthreads();
        for (int k = threadIdx.x; k < 6 * kBins; k += 256)
            if (s_hist[k]) atomicAdd(&hist[size_t(uni) * 6 * kBins + k], s_hist[k]);
    }
}

__device__ inline float box_area(const float* lo, const float* hi) {
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return 2.f * (dx * dy + dx * dz + dy * dz);
}

// ---- select: one warp per node
__global__ void __launch_bounds__(256) select_kernel(Level lv, uint32_t num_nodes, const uint32_t* __restrict__ hist, Decision dc, int level,
                                                     float scene_scale, float touch, SahParams sah) {
    const uint32_t node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    if (node >= num_nodes) return;
    const uint32_t n = lv.count[node];
    int axis = -1;
    float pos = 0.f;
    if (n == 0) {
        axis = -2;
    } else {
        float blo[3], bhi[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            blo[c] = lv.blo[3 * node + c];
            bhi[c] = lv.bhi[3 * node + c];
        }
        const float area = box_area(blo, bhi);
        if (n > sah.leaf_max && level < kMaxLevels && area > 0.f) { // lib/kdtree.cpp:133-141
            float best = kFltMax, best_pos = 0.f;
            int best_ax = -1;
            uint32_t best_nl = 0, best_nr = 0;
            for (int ax = 0; ax < 3; ++ax) {
                const float tl = o2f(lv.tlo[3 * node + ax]), th = o2f(lv.thi[3 * node + ax]);
                const float ext = th - tl;
                if (!(ext > 0.f)) continue;
                // a box this thin is not split along this axis any more: references within `touch` of a plane go to both
                // children, so planes closer together than that separate nothing (and fp32 cannot resolve them from afar)
                if (!(bhi[ax] - blo[ax] > 64.f * touch)) continue;
                const uint32_t s = hist[size_t(node) * 6 * kBins + (2 * ax) * kBins + lane];
                const uint32_t e = hist[size_t(node) * 6 * kBins + (2 * ax + 1) * kBins + lane];
                uint32_t ps = s, pe = e; // inclusive scans
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t vs = __shfl_up_sync(0xffffffffu, ps, off), ve = __shfl_up_sync(0xffffffffu, pe, off);
                    if (static_cast<int>(lane) >= off) {
                        ps += vs;
                        pe += ve;
                    }
                }
                const uint32_t tot_e = __shfl_sync(0xffffffffu, pe, 31);
                // plane k = lane (k = 0..31) and, on lane 31 as a second candidate, plane 32
                for (int extra = 0; extra < 2; ++extra) {
                    if (extra == 1 && lane != 31u) break;
                    const int k = extra ? kBins : static_cast<int>(lane);
                    const uint32_t nl = extra ? ps : ps - s;           // references that start below plane k
                    const uint32_t nr = extra ? 0u : tot_e - (pe - e); // references that end at or above plane k
                    float p = k == kBins ? th : tl + static_cast<float>(k) * (ext / static_cast<float>(kBins));
                    const float void_shift = 4e-6f * fmaxf(fabsf(p), scene_scale);
                    if (nl == 0u) p -= void_shift; // cut: plane into the void below the references
                    if (nr == 0u) p += void_shift;
                    if (!(p > blo[ax] + 8.f * touch && p < bhi[ax] - 8.f * touch)) continue; // the plane must split the node's box
                    if (nl == 0u && nr == 0u) continue;
                    float llo[3] = {blo[0], blo[1], blo[2]}, lhi[3] = {bhi[0], bhi[1], bhi[2]};
                    float rlo[3] = {blo[0], blo[1], blo[2]}, rhi[3] = {bhi[0], bhi[1], bhi[2]};
                    lhi[ax] = p;
                    rlo[ax] = p;
                    const float lam = (nl == 0u || nr == 0u) ? sah.lambda : 1.f;
                    const float cost = lam * (sah.kt + sah.ki * (box_area(llo, lhi) / area * static_cast<float>(nl) +
                                                                 box_area(rlo, rhi) / area * static_cast<float>(nr)));
                    if (cost < best) {
                        best = cost;
                        best_pos = p;
                        best_ax = ax;
                        best_nl = nl;
                        best_nr = nr;
                    }
                }
            }
            // warp arg-min (ties: lowest lane, then lowest axis by evaluation order)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float oc = __shfl_down_sync(0xffffffffu, best, off);
                const float op = __shfl_down_sync(0xffffffffu, best_pos, off);
                const int oa = __shfl_down_sync(0xffffffffu, best_ax, off);
                const uint32_t onl = __shfl_down_sync(0xffffffffu, best_nl, off), onr = __shfl_down_sync(0xffffffffu, best_nr, off);
                if (oc < best) {
                    best = oc;
                    best_pos = op;
                    best_ax = oa;
                    best_nl = onl;
                    best_nr = onr;
                }
            }
            best = __shfl_sync(0xffffffffu, best, 0);
            best_pos = __shfl_sync(0xffffffffu, best_pos, 0);
            best_ax = __shfl_sync(0xffffffffu, best_ax, 0);
            best_nl = __shfl_sync(0xffffffffu, best_nl, 0);
            best_nr = __shfl_sync(0xffffffffu, best_nr, 0);
            if (best_ax >= 0) {
                const float lam = (best_nl == 0u || best_nr == 0u) ? sah.lambda : 1.f;
                // automatic termination, lib/kdtree.cpp:153-158; and a split that separates nothing is no split
                const bool no_progress = best_nl >= n && best_nr >= n;
                if (!(sah.ki * static_cast<float>(n) * lam < best) && !no_progress) {
                    axis = best_ax;
                    pos = best_pos;
                }
            }
        }
    }
    if (lane == 0) {
        dc.axis[node] = axis;
        dc.pos[node] = pos;
        dc.is_split[node] = axis >= 0 ? 1u : 0u;
        dc.leaf_alloc[node] = axis == -1 ? ((n + 3u) & ~3u) : 0u;
    }
}

// ---- apply: the node's word in the sibling-pair layout; open the children
__global__ void apply_kernel(Level lv, uint32_t num_nodes, Decision dc, Level next, uint2* __restrict__ pair_nodes, uint32_t pair_base,
                             uint32_t pool_base, Totals* totals) {
    const uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = node < num_nodes;
    const int axis = valid ? dc.axis[node] : -2;
    {
        // references that end in leaves: one atomic per warp
        uint32_t sum = (valid && axis == -1) ? lv.count[node] : 0u;
        uint32_t fn = (valid && axis == -1) ? lv.count[node] / 2u + 1u : 0u;
        for (int off = 16; off > 0; off >>= 1) {
            sum += __shfl_down_sync(0xffffffffu, sum, off);
            fn += __shfl_down_sync(0xffffffffu, fn, off);
        }
        if ((threadIdx.x & 31u) == 0u && sum) {
            atomicAdd(&totals->leaf_refs, sum);
            atomicAdd(&totals->leaf_ref_nodes, static_cast<unsigned long long>(fn));
        }
    }
    if (!valid) return;
    const uint32_t slot = lv.slot[node];
    if (axis == -2) {
        pair_nodes[slot] = make_uint2(0u, 3u); // void
        return;
    }
    if (axis == -1) {
        const uint32_t first = pool_base + dc.leaf_off[node];
        dc.leaf_first[node] = first;
        pair_nodes[slot] = make_uint2(first, (lv.count[node] << 2) | 3u);
        return;
    }
    const uint32_t c0 = 2u * dc.split_idx[node];
    const uint32_t pair = pair_base + c0;
    const float pos = dc.pos[node];
    pair_nodes[slot] = make_uint2(__float_as_uint(pos), (pair << 2) | static_cast<uint32_t>(axis));
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const uint32_t c = c0 + side;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float lo = lv.blo[3 * node + k], hi = lv.bhi[3 * node + k];
            if (k == axis) {
                if (side == 0) hi = pos;
                else lo = pos;
            }
            next.blo[3 * c + k] = lo;
            next.bhi[3 * c + k] = hi;
            next.tlo[3 * c + k] = f2o(kFltMax);
            next.thi[3 * c + k] = f2o(-kFltMax);
        }
        next.count[c] = 0;
        next.slot[c] = pair + side;
    }
}

__device__ inline void classify(const float4& a, const float4& b, int axis, float pos, float eps, bool& left, bool& right) {
    const float lo = axis == 0 ? a.x : (axis == 1 ? a.y : a.z), hi = axis == 0 ? b.x : (axis == 1 ? b.y : b.z);
    // a box that only touches the plane (within eps) goes to both sides: whichever side a ray arrives from, the cell it
    // is in holds the triangle
    left = lo <= pos + eps;
    right = hi >= pos - eps;
    if (!left && !right) left = true; // cannot happen for lo <= hi; keeps every reference somewhere
}

// ---- count: leaf references go to the pool; the others count into their children
__global__ void __launch_bounds__(256) count_kernel(const float4* __restrict__ ra, const float4* __restrict__ rb, uint32_t nrefs, Level lv,
                                                    Decision dc, Level next, uint32_t* __restrict__ pool_ids, uint32_t* __restrict__ pool_key,
                                                    float eps) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrefs) return;
    const float4 a = ra[i], b = rb[i];
    const uint32_t node = __float_as_uint(b.w);
    const int axis = dc.axis[node];
    if (axis == -1) {
        const uint32_t first = dc.leaf_first[node], local = i - lv.first[node], cnt = lv.count[node];
        pool_ids[first + local] = __float_as_uint(a.w);
        pool_key[first + local] = first;
        if (local == 0) // the padding of the run (a run starts at a multiple of 4 ids): sorts behind the run's ids
            for (uint32_t k = cnt; k < ((cnt + 3u) & ~3u); ++k) {
                pool_ids[first + k] = 0xFFFFFFFFu;
                pool_key[first + k] = first;
            }
        return;
    }
    if (axis < 0) return;
    bool left, right;
    classify(a, b, axis, dc.pos[node], eps, left, right);
    const uint32_t c0 = 2u * dc.split_idx[node];
    if (left) atomicAdd(&next.count[c0], 1u);
    if (right) atomicAdd(&next.count[c0 + 1], 1u);
}

// Sutherland-Hodgman clip of a triangle against an axis-aligned box; returns the bounds of what is left (false: nothing)
__device__ inline bool clip_triangle_bounds(const float* __restrict__ tri9, const float* blo, const float* bhi, float* olo, float* ohi) {
    float px[10][3], qx[10][3];
    int n = 3;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) px[k][c] = tri9[3 * k + c];
    for (int ax = 0; ax < 3; ++ax) {
        for (int side = 0; side < 2; ++side) {
            const float bound = side == 0 ? blo[ax] : bhi[ax];
            int m = 0;
            for (int i = 0; i < n; ++i) {
                const float* A = px[(i + n - 1) % n];
                const float* B = px[i];
                const float da = side == 0 ? A[ax] - bound : bound - A[ax];
                const float db = side == 0 ? B[ax] - bound : bound - B[ax];
                if ((da < 0.f) != (db < 0.f) && m < 9) {
                    const float t = da / (da - db);
#pragma unroll
                    for (int c = 0; c < 3; ++c) qx[m][c] = c == ax ? bound : A[c] + t * (B[c] - A[c]);
                    ++m;
                }
                if (db >= 0.f && m < 9) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) qx[m][c] = B[c];
                    ++m;
                }
            }
            n = m;
            if (n == 0) return false;
            for (int i = 0; i < n; ++i)
#pragma unroll
                for (int c = 0; c < 3; ++c) px[i][c] = qx[i][c];
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float mn = kFltMax, mx = -kFltMax;
        for (int i = 0; i < n; ++i) {
            mn = fminf(mn, px[i][c]);
            mx = fmaxf(mx, px[i][c]);
        }
        olo[c] = mn;
        ohi[c] = mx;
    }
    return true;
}

// ---- scatter: references move into their children's segments
__global__ void __launch_bounds__(128) scatter_kernel(const float4* __restrict__ ra, const float4* __restrict__ rb, uint32_t nrefs, Decision dc,
                                                      Level next, uint32_t* __restrict__ cursor, const float* __restrict__ verts,
                                                      float4* __restrict__ oa, float4* __restrict__ ob, float eps) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrefs) return;
    const float4 a = ra[i], b = rb[i];
    const uint32_t node = __float_as_uint(b.w);
    const int axis = dc.axis[node];
    if (axis < 0) return;
    const float pos = dc.pos[node];
    bool left, right;
    classify(a, b, axis, pos, eps, left, right);
    const uint32_t c0 = 2u * dc.split_idx[node];
    const uint32_t tri = __float_as_uint(a.w);
    for (int side = 0; side < 2; ++side) {
        if (side == 0 ? !left : !right) continue;
        const uint32_t c = c0 + side;
        float lo[3] = {a.x, a.y, a.z}, hi[3] = {b.x, b.y, b.z};
        float clo[3], chi[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            clo[k] = next.blo[3 * c + k];
            chi[k] = next.bhi[3 * c + k];
        }
        if (left && right) {
            // straddler: its exact extent inside the child's box ("perfect split"); never larger than the box-clipped
            // reference box, never empty (a touching triangle keeps a degenerate box at the plane)
            float tlo[3], thi[3];
            const bool any = clip_triangle_bounds(verts + size_t(tri) * 9, clo, chi, tlo, thi);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float l = fmaxf(lo[k], clo[k]), h = fminf(hi[k], chi[k]);
                if (any) {
                    l = fmaxf(l, tlo[k]);
                    h = fminf(h, thi[k]);
                }
                if (h < l) { // touching / rounding: collapse onto the child's side of the reference box
                    l = fminf(fmaxf(lo[k], clo[k]), chi[k]);
                    h = l;
                }
                lo[k] = l;
                hi[k] = h;
            }
        }
        const uint32_t dst = next.first[c] + atomicAdd(&cursor[c], 1u);
        oa[dst] = make_float4(lo[0], lo[1], lo[2], __uint_as_float(tri));
        ob[dst] = make_float4(hi[0], hi[1], hi[2], __uint_as_float(c));
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(&next.tlo[3 * c + k], f2o(lo[k]));
            atomicMax(&next.thi[3 * c + k], f2o(hi[k]));
        }
    }
}

__global__ void totals_kernel(const Decision dc, uint32_t num_nodes, Totals* t) {
    // exclusive scans + the last element = totals (one thread; three loads)
    const uint32_t last = num_nodes - 1;
    t->num_split = dc.split_idx[last] + dc.is_split[last];
    t->leaf_alloc = dc.leaf_off[last] + dc.leaf_alloc[last];
}
__global__ void next_totals_kernel(const Level next, uint32_t num_next, Totals* t) {
    t->next_refs = num_next ? next.first[num_next - 1] + next.count[num_next - 1] : 0u;
}
__global__ void count_cuts_kernel(const Level next, uint32_t num_next, Totals* t) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= num_next) return;
    if ((c & 1u) == 0u && (next.count[c] == 0u) != (next.count[c + 1] == 0u)) atomicAdd(&t->cut_nodes, 1u);
}

__global__ void make_keys_kernel(const uint32_t* __restrict__ pool_key, const uint32_t* __restrict__ pool_ids, uint32_t n,
                                 unsigned long long* __restrict__ keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = (static_cast<unsigned long long>(pool_key[i]) << 32) | pool_ids[i];
}
__global__ void finish_pool_kernel(const unsigned long long* __restrict__ keys, uint32_t n, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t id = static_cast<uint32_t>(keys[i]);
    for (uint32_t back = 1; id == 0xFFFFFFFFu && back <= 3 && back <= i; ++back) id = static_cast<uint32_t>(keys[i - back]); // padding repeats the run's last id
    out[i] = id == 0xFFFFFFFFu ? 0u : id;
}

// ---- depth-first re-layout of the finished tree. The levels are built breadth-first (a level's pairs are contiguous);
// a walk then jumps across the whole level array at every step. In depth-first order the children pair of a pair's first
// inner node follows it directly (same 128-byte line), and a subtree is contiguous: L1 / L2 lines are shared along a path.
__global__ void subtree_count_kernel(const uint2* __restrict__ nodes, uint32_t begin, uint32_t end, uint32_t* __restrict__ cnt) {
    const uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    const uint2 n = nodes[i];
    uint32_t c = 0;
    if ((n.y & 3u) != 3u) {
        const uint32_t child = n.y >> 2;
        c = 1u + cnt[child] + cnt[child + 1];
    }
    cnt[i] = c; // pairs in the subtree below this node
}
__global__ void assign_pos_kernel(const uint2* __restrict__ nodes, uint32_t begin, uint32_t end, const uint32_t* __restrict__ cnt,
                                  uint32_t* __restrict__ pos) {
    const uint32_t q = begin / 2u + blockIdx.x * blockDim.x + threadIdx.x; // pair index
    if (2u * q >= end) return;
    const uint2 a = nodes[2u * q], b = nodes[2u * q + 1];
    const uint32_t P = pos[q];
    if ((a.y & 3u) != 3u) pos[(a.y >> 2) / 2u] = P + 1u;
    if ((b.y & 3u) != 3u) pos[(b.y >> 2) / 2u] = P + 1u + cnt[2u * q];
}
__global__ void relayout_kernel(const uint2* __restrict__ nodes, uint32_t npairs, const uint32_t* __restrict__ pos, uint2* __restrict__ out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= npairs) return;
    const uint32_t P = pos[q];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        uint2 n = nodes[2u * q + k];
        if ((n.y & 3u) != 3u) n.y = ((2u * pos[(n.y >> 2) / 2u]) << 2) | (n.y & 3u);
        out[2u * P + k] = n;
    }
}

// One device allocation for the whole build: ~35 separate cudaMalloc / cudaFree pairs cost more than the build's kernels
// (and their cost varies with the state of the process). Buffers are carved from the arena; one that does not fit any more
// (a retry with a larger budget estimates generously, so this is rare) falls back to its own allocation.
struct Arena {
    char* base = nullptr;
    size_t off = 0, cap = 0;
    ~Arena() { cudaFree(base); }
    cudaError_t reserve(size_t bytes) {
        cap = bytes;
        return cudaMalloc(reinterpret_cast<void**>(&base), bytes);
    }
    void* take(size_t bytes) {
        const size_t at = (off + 255) & ~size_t(255);
        if (!base || at + bytes > cap) return nullptr;
        off = at + bytes;
        return base + at;
    }
};
static thread_local Arena* t_arena = nullptr;

struct Buf {
    void* p = nullptr;
    bool owned = false;
    ~Buf() {
        if (owned) cudaFree(p);
    }
    cudaError_t alloc(size_t bytes) {
        if (bytes == 0) bytes = 16;
        if (t_arena && (p = t_arena->take(bytes)) != nullptr) return cudaSuccess;
        owned = true;
        return cudaMalloc(&p, bytes);
    }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

struct LevelMem {
    Buf blo, bhi, tlo, thi, first, count, slot;
    cudaError_t alloc(size_t cap) {
        cudaError_t e;
        if ((e = blo.alloc(cap * 12)) != cudaSuccess) return e;
        if ((e = bhi.alloc(cap * 12)) != cudaSuccess) return e;
        if ((e = tlo.alloc(cap * 12)) != cudaSuccess) return e;
        if ((e = thi.alloc(cap * 12)) != cudaSuccess) return e;
        if ((e = first.alloc(cap * 4)) != cudaSuccess) return e;
        if ((e = count.alloc(cap * 4)) != cudaSuccess) return e;
        return slot.alloc(cap * 4);
    }
    Level view() const {
        return Level{blo.as<float>(), bhi.as<float>(), tlo.as<uint32_t>(), thi.as<uint32_t>(), first.as<uint32_t>(), count.as<uint32_t>(),
                     slot.as<uint32_t>()};
    }
};

static inline unsigned blocks(uint64_t n, unsigned bs) { return static_cast<unsigned>((n + bs - 1) / bs); }

} // namespace gpubuild

static int build_impl(const HostTriangles& tris, KdTree& out, int device, std::string& err, uint64_t budget,
                      std::chrono::steady_clock::time_point t0) {
    using namespace gpubuild;
    const uint32_t n = tris.count;
    GB_TRY(cudaSetDevice(device));
    // scene box exactly as the host builder / KDTree::KDTree compute it (lib/kdtree.cpp:474-490)
    const float* v = tris.verts.data();
    float box[6];
    for (uint32_t i = 0; i < n; ++i) {
        const float* p = v + size_t(i) * 9;
        for (int c = 0; c < 3; ++c) {
            const float mn = std::fmin(p[c], std::min(p[3 + c], p[6 + c])), mx = std::fmax(p[c], std::max(p[3 + c], p[6 + c]));
            const float lo = mx < mn ? mx : mn, hi = mn < mx ? mx : mn;
            if (i == 0) {
                box[c] = lo;
                box[3 + c] = hi;
            } else {
                box[c] = lo < box[c] ? lo : box[c];
                box[3 + c] = box[3 + c] < hi ? hi : box[3 + c];
            }
        }
    }
    float scale = 0.f;
    for (int c = 0; c < 6; ++c) scale = std::fmax(scale, std::fabs(box[c]));
    const float eps = 1e-6f * scale;

    // capacities: references per level, nodes per level, nodes of the whole tree. Sized for meshes (the 1M mesh peaks at
    // 2.1 references per triangle on a level, 0.47 nodes per triangle on a level, 4.8 nodes per triangle in total);
    // overlapping soups need more: the caller retries with a 4x / 16x budget. Allocation is a third of the build time.
    const uint64_t ref_cap = std::max<uint64_t>(uint64_t(n) * 6u, 1u << 18) * budget;
    const uint64_t node_cap = std::max<uint64_t>(uint64_t(n), 1u << 16) * budget;
    const uint64_t pair_cap = std::max<uint64_t>(uint64_t(n) * 8u, 1u << 18) * budget;
    if (ref_cap >= (1ull << 32) || pair_cap >= (1ull << 30)) {
        err = "scene too large for the device kd builder";
        return -5;
    }
    Arena arena;
    {
        const size_t total = size_t(n) * 36 + 4 * ref_cap * 16 + 2 * node_cap * 60 + node_cap * 6 * kBins * 4 + pair_cap * 8 + 2 * ref_cap * 4 +
                             node_cap * 4 * 8 + pair_cap * 14 + ref_cap * 28 + (size_t(8) << 20);
        if (arena.reserve(total) != cudaSuccess) { // not enough memory in one piece: separate allocations may still fit
            cudaGetLastError();
            arena.base = nullptr;
        }
    }
    struct ArenaScope {
        explicit ArenaScope(Arena* a) { t_arena = a; }
        ~ArenaScope() { t_arena = nullptr; }
    } arena_scope(&arena);
    Buf d_verts, ra[2], rb[2], hist, pair_nodes, pool_ids, pool_key, cursor, totals, scan_tmp, keys, keys_alt, pool_out;
    Buf d_axis, d_pos, d_is_split, d_split_idx, d_leaf_alloc, d_leaf_off, d_leaf_first;
    LevelMem lm[2];
    GB_TRY(d_verts.alloc(size_t(n) * 36));
    GB_TRY(cudaMemcpy(d_verts.p, v, size_t(n) * 36, cudaMemcpyHostToDevice));
    for (int k = 0; k < 2; ++k) {
        GB_TRY(ra[k].alloc(ref_cap * 16));
        GB_TRY(rb[k].alloc(ref_cap * 16));
        GB_TRY(lm[k].alloc(node_cap));
    }
    GB_TRY(hist.alloc(node_cap * 6 * kBins * 4));
    GB_TRY(pair_nodes.alloc(pair_cap * 8));
    GB_TRY(pool_ids.alloc(ref_cap * 4));
    GB_TRY(pool_key.alloc(ref_cap * 4));
    GB_TRY(cursor.alloc(node_cap * 4));
    GB_TRY(totals.alloc(sizeof(Totals)));
    GB_TRY(d_axis.alloc(node_cap * 4));
    GB_TRY(d_pos.alloc(node_cap * 4));
    GB_TRY(d_is_split.alloc(node_cap * 4));
    GB_TRY(d_split_idx.alloc(node_cap * 4));
    GB_TRY(d_leaf_alloc.alloc(node_cap * 4));
    GB_TRY(d_leaf_off.alloc(node_cap * 4));
    GB_TRY(d_leaf_first.alloc(node_cap * 4));
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_is_split.as<uint32_t>(), d_split_idx.as<uint32_t>(), static_cast<int>(node_cap));
    GB_TRY(scan_tmp.alloc(scan_bytes));
    Decision dc{d_axis.as<int>(), d_pos.as<float>(), d_is_split.as<uint32_t>(), d_split_idx.as<uint32_t>(), d_leaf_alloc.as<uint32_t>(),
                d_leaf_off.as<uint32_t>(), d_leaf_first.as<uint32_t>()};
    GB_TRY(cudaMemset(totals.p, 0, sizeof(Totals)));

    const auto t_alloc = std::chrono::steady_clock::now();
    // root
    {
        Level l0 = lm[0].view();
        uint32_t tl[3], th[3], zero = 0;
        for (int c = 0; c < 3; ++c) {
            tl[c] = f2o(box[c]);
            th[c] = f2o(box[3 + c]);
        }
        GB_TRY(cudaMemcpy(l0.blo, box, 12, cudaMemcpyHostToDevice));
        GB_TRY(cudaMemcpy(l0.bhi, box + 3, 12, cudaMemcpyHostToDevice));
        GB_TRY(cudaMemcpy(l0.tlo, tl, 12, cudaMemcpyHostToDevice));
        GB_TRY(cudaMemcpy(l0.thi, th, 12, cudaMemcpyHostToDevice));
        GB_TRY(cudaMemcpy(l0.first, &zero, 4, cudaMemcpyHostToDevice));
        GB_TRY(cudaMemcpy(l0.count, &n, 4, cudaMemcpyHostToDevice));
        GB_TRY(cudaMemcpy(l0.slot, &zero, 4, cudaMemcpyHostToDevice));
        const uint2 pad = make_uint2(0u, 3u); // node 1 pads the root's pair: an empty leaf
        GB_TRY(cudaMemcpy(pair_nodes.as<uint2>() + 1, &pad, 8, cudaMemcpyHostToDevice));
        init_refs_kernel<<<blocks(n, 256), 256>>>(d_verts.as<float>(), n, ra[0].as<float4>(), rb[0].as<float4>());
    }
    const bool debug = std::getenv("TRN_KD_DEBUG") != nullptr;
    auto envf = [](const char* name, float dflt) {
        const char* v = std::getenv(name);
        return (v && *v) ? static_cast<float>(std::atof(v)) : dflt;
    };
    const SahParams sah{envf("TRN_KD_KT", kCostTraversal), envf("TRN_KD_KI", kCostIntersection), envf("TRN_KD_LAMBDA", kLambdaEmpty),
                        static_cast<uint32_t>(envf("TRN_KD_LEAF", static_cast<float>(kLeafMax)))};
    uint32_t num_nodes = 1, nrefs = n, pair_count = 2, pool_count = 0;
    uint64_t splits_total = 0;
    std::vector<uint32_t> level_begin{0u}; // first node of every level in pair_nodes (level 0 = the root pair)
    int cur = 0, level = 0;
    uint64_t height = 0;
    while (num_nodes > 0) {
        Level lv = lm[cur].view(), nx = lm[cur ^ 1].view();
        GB_TRY(cudaMemsetAsync(hist.p, 0, size_t(num_nodes) * 6 * kBins * 4));
        if (nrefs) bin_kernel<<<blocks(nrefs, 1024), 256>>>(ra[cur].as<float4>(), rb[cur].as<float4>(), nrefs, lv, hist.as<uint32_t>());
        select_kernel<<<blocks(uint64_t(num_nodes) * 32, 256), 256>>>(lv, num_nodes, hist.as<uint32_t>(), dc, level, scale, eps, sah);
        size_t tmp = scan_bytes;
        cub::DeviceScan::ExclusiveSum(scan_tmp.p, tmp, dc.is_split, dc.split_idx, static_cast<int>(num_nodes));
        tmp = scan_bytes;
        cub::DeviceScan::ExclusiveSum(scan_tmp.p, tmp, dc.leaf_alloc, dc.leaf_off, static_cast<int>(num_nodes));
        totals_kernel<<<1, 1>>>(dc, num_nodes, totals.as<Totals>());
        Totals t;
        GB_TRY(cudaMemcpy(&t, totals.p, sizeof t, cudaMemcpyDeviceToHost));
        const uint32_t num_next = 2 * t.num_split;
        if (uint64_t(pair_count) + num_next > pair_cap || uint64_t(pool_count) + t.leaf_alloc + 4 > ref_cap || num_next > node_cap) {
            err = "device kd builder: tree exceeds its node / reference budget (level " + std::to_string(level) + ": " +
                  std::to_string(num_nodes) + " nodes, " + std::to_string(t.num_split) + " splits, " + std::to_string(t.leaf_alloc) +
                  " leaf slots, " + std::to_string(pair_count) + " pairs so far)";
            return -5;
        }
        apply_kernel<<<blocks(num_nodes, 256), 256>>>(lv, num_nodes, dc, nx, pair_nodes.as<uint2>(), pair_count, pool_count, totals.as<Totals>());
        if (nrefs)
            count_kernel<<<blocks(nrefs, 256), 256>>>(ra[cur].as<float4>(), rb[cur].as<float4>(), nrefs, lv, dc, nx, pool_ids.as<uint32_t>(),
                                                       pool_key.as<uint32_t>(), eps);
        uint32_t next_refs = 0;
        if (num_next) {
            tmp = scan_bytes;
            cub::DeviceScan::ExclusiveSum(scan_tmp.p, tmp, nx.count, nx.first, static_cast<int>(num_next));
            next_totals_kernel<<<1, 1>>>(nx, num_next, totals.as<Totals>());
            count_cuts_kernel<<<blocks(num_next, 256), 256>>>(nx, num_next, totals.as<Totals>());
            GB_TRY(cudaMemcpy(&t, totals.p, sizeof t, cudaMemcpyDeviceToHost));
            next_refs = t.next_refs;
            if (next_refs > ref_cap) {
                err = "device kd builder: a level outgrew the reference buffer";
                return -5;
            }
            GB_TRY(cudaMemsetAsync(cursor.p, 0, size_t(num_next) * 4));
            scatter_kernel<<<blocks(nrefs, 128), 128>>>(ra[cur].as<float4>(), rb[cur].as<float4>(), nrefs, dc, nx, cursor.as<uint32_t>(),
                                                        d_verts.as<float>(), ra[cur ^ 1].as<float4>(), rb[cur ^ 1].as<float4>(), eps);
        }
        GB_TRY(cudaGetLastError());
        if (debug)
            std::fprintf(stderr, "[kd-gpu] level %d: nodes %u refs %u -> split %u leaf_alloc %u next_refs %u (pairs %u pool %u)\n", level,
                         num_nodes, nrefs, t.num_split, t.leaf_alloc, next_refs, pair_count, pool_count);
        if (num_next) level_begin.push_back(pair_count);
        pair_count += num_next;
        pool_count += t.leaf_alloc;
        splits_total += t.num_split;
        if (num_next) height = static_cast<uint64_t>(level) + 1;
        num_nodes = num_next;
        nrefs = next_refs;
        cur ^= 1;
        ++level;
        if (level > kMaxLevels + 2) {
            err = "device kd builder: did not terminate";
            return -5;
        }
    }
    // depth-first re-layout (TRN_KD_LAYOUT=bfs keeps the level order, for A/B)
    const char* layout_env = std::getenv("TRN_KD_LAYOUT");
    Buf relaid;
    uint2* final_nodes = pair_nodes.as<uint2>();
    if (!(layout_env && std::strcmp(layout_env, "bfs") == 0) && pair_count > 2) {
        Buf cnt, pos;
        GB_TRY(cnt.alloc(size_t(pair_count) * 4));
        GB_TRY(pos.alloc(size_t(pair_count / 2) * 4));
        GB_TRY(relaid.alloc(size_t(pair_count) * 8));
        level_begin.push_back(pair_count);
        const int nl = static_cast<int>(level_begin.size()) - 1;
        for (int l = nl - 1; l >= 0; --l) {
            const uint32_t b = level_begin[l], e = level_begin[l + 1];
            subtree_count_kernel<<<blocks(e - b, 256), 256>>>(pair_nodes.as<uint2>(), b, e, cnt.as<uint32_t>());
        }
        GB_TRY(cudaMemsetAsync(pos.p, 0, 4)); // the root pair stays first
        for (int l = 0; l < nl; ++l) {
            const uint32_t b = level_begin[l], e = level_begin[l + 1];
            assign_pos_kernel<<<blocks((e - b) / 2, 256), 256>>>(pair_nodes.as<uint2>(), b, e, cnt.as<uint32_t>(), pos.as<uint32_t>());
        }
        relayout_kernel<<<blocks(pair_count / 2, 256), 256>>>(pair_nodes.as<uint2>(), pair_count / 2, pos.as<uint32_t>(), relaid.as<uint2>());
        GB_TRY(cudaGetLastError());
        GB_TRY(cudaDeviceSynchronize()); // cnt / pos go out of scope
        final_nodes = relaid.as<uint2>();
    }
    const auto t_levels = std::chrono::steady_clock::now();
    // ids inside every leaf in ascending order (the reference's leaves hold ascending ids as well), padding last
    const uint32_t pool_total = pool_count + 4; // a 4-wide read of the last chunk stays inside
    std::vector<uint32_t> host_pool(pool_total, 0u);
    if (pool_count) {
        GB_TRY(keys.alloc(size_t(pool_count) * 8));
        GB_TRY(keys_alt.alloc(size_t(pool_count) * 8));
        GB_TRY(pool_out.alloc(size_t(pool_count) * 4));
        make_keys_kernel<<<blocks(pool_count, 256), 256>>>(pool_key.as<uint32_t>(), pool_ids.as<uint32_t>(), pool_count, keys.as<unsigned long long>());
        cub::DoubleBuffer<unsigned long long> kb(keys.as<unsigned long long>(), keys_alt.as<unsigned long long>());
        size_t sort_bytes = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, kb, static_cast<int>(pool_count), 0, 64);
        Buf sort_tmp;
        GB_TRY(sort_tmp.alloc(sort_bytes));
        cub::DeviceRadixSort::SortKeys(sort_tmp.p, sort_bytes, kb, static_cast<int>(pool_count), 0, 64);
        finish_pool_kernel<<<blocks(pool_count, 256), 256>>>(kb.Current(), pool_count, pool_out.as<uint32_t>());
        GB_TRY(cudaGetLastError());
        GB_TRY(cudaMemcpy(host_pool.data(), pool_out.p, size_t(pool_count) * 4, cudaMemcpyDeviceToHost));
        for (uint32_t k = 0; k < 4; ++k) host_pool[pool_count + k] = host_pool[pool_count - 1];
    }
    Totals t;
    GB_TRY(cudaMemcpy(&t, totals.p, sizeof t, cudaMemcpyDeviceToHost));
    out.pair_nodes.resize(pair_count);
    static_assert(sizeof(uint2) == sizeof(uint64_t), "pair node = 8 bytes");
    GB_TRY(cudaMemcpy(out.pair_nodes.data(), final_nodes, size_t(pair_count) * 8, cudaMemcpyDeviceToHost)); // little-endian: x low, y high
    out.pair_leaf_refs.swap(host_pool);
    for (int c = 0; c < 6; ++c) out.box[c] = box[c];
    out.height = height;
    out.num_pair_refs = t.leaf_refs;
    out.num_leaf_refs = t.leaf_refs;
    out.num_cut_nodes = t.cut_nodes;
    out.nodes.clear(); // the reference-shaped array is derived on demand (reference_shape_from_pairs)
    out.expected_nodes = splits_total - t.cut_nodes + t.leaf_ref_nodes; // its size: inner nodes that are no cuts + leaf runs
    out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (debug) { // chains of empty-space cuts: how many cut nodes sit directly below another cut
        const auto& pn = out.pair_nodes;
        auto y_of = [&](uint32_t i) { return static_cast<uint32_t>(pn[i] >> 32); };
        auto is_cut = [&](uint32_t i) {
            const uint32_t y = y_of(i);
            if ((y & 3u) == 3u) return false;
            const uint32_t c = y >> 2;
            return (y_of(c) == 3u) != (y_of(c + 1) == 3u);
        };
        uint64_t hist[8] = {0};
        std::vector<uint8_t> below_cut(pn.size(), 0);
        for (uint32_t i = 0; i < pn.size(); ++i)
            if (is_cut(i)) {
                const uint32_t c = y_of(i) >> 2;
                below_cut[y_of(c) == 3u ? c + 1 : c] = 1;
            }
        for (uint32_t i = 0; i < pn.size(); ++i)
            if (is_cut(i) && !below_cut[i]) { // head of a chain
                uint32_t len = 0, j = i;
                while (is_cut(j)) {
                    ++len;
                    const uint32_t c = y_of(j) >> 2;
                    j = y_of(c) == 3u ? c + 1 : c;
                }
                hist[std::min<uint32_t>(len, 7)]++;
            }
        std::fprintf(stderr, "[kd-gpu] cut chains by length 1..7+: %llu %llu %llu %llu %llu %llu %llu\n", (unsigned long long)hist[1],
                     (unsigned long long)hist[2], (unsigned long long)hist[3], (unsigned long long)hist[4], (unsigned long long)hist[5],
                     (unsigned long long)hist[6], (unsigned long long)hist[7]);
    }
    if (debug) {
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double, std::milli>(b - a).count();
        };
        std::fprintf(stderr, "[kd-gpu] box + alloc + upload %.1f ms, %d levels %.1f ms, leaf sort + download %.1f ms, total %.1f ms\n",
                     ms(t0, t_alloc), level, ms(t_alloc, t_levels), ms(t_levels, std::chrono::steady_clock::now()), out.build_ms);
    }
    return 0;
}

int build_kdtree_gpu(const HostTriangles& tris, KdTree& out, int device, std::string& err) {
    const auto t0 = std::chrono::steady_clock::now();
    int rc = 0;
    for (uint64_t budget : {1u, 4u, 16u}) { // -5 = a level outgrew its buffers: once more with larger ones
        rc = build_impl(tris, out, device, err, budget, t0);
        if (rc != -5) break;
    }
    // development aid (tools/pooled_emul.cpp replays the production tree on the CPU): TRN_KD_DUMP=<file> writes the pair
    // layout -- u64 pair-node count, u64 reference count, the nodes, the references
    if (rc == 0) {
        if (const char* path = std::getenv("TRN_KD_DUMP")) {
            if (FILE* f = std::fopen(path, "wb")) {
                const uint64_t hdr[2] = {out.pair_nodes.size(), out.pair_leaf_refs.size()};
                std::fwrite(hdr, 8, 2, f);
                std::fwrite(out.pair_nodes.data(), 8, out.pair_nodes.size(), f);
                std::fwrite(out.pair_leaf_refs.data(), 4, out.pair_leaf_refs.size(), f);
                std::fclose(f);
            }
        }
    }
    return rc;
}

} // namespace trn
