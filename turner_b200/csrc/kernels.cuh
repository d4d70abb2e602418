// sm_100a wavefront kernel set of the path tracer: raygen, closest-hit kd traversal,
// shade/bounce (+ warp-ballot compaction of the next wave), any-hit shadow traversal.
//
// Arithmetic contract (tests/test_gpu_parity.py): everything that decides WHICH
// triangle is hit -- primary ray setup, slab test, split-plane distances, the
// ray/triangle test -- is the reference's fp32 operation sequence
// (lib/types.h:119-123, lib/intersection.h:40-128, lib/kdtree.cpp:515-607) with
// IEEE division and no FMA contraction: this file is compiled with -fmad=false
// (-prec-div/-prec-sqrt default to true). Hit triangle ids and (r,s,t) are
// therefore bit-identical to the reference's.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace trn {

constexpr float kEpsDir = 0.00001f;            // EPS, lib/types.h:13 (fix_direction, lib/kdtree.cpp:503-511)
constexpr uint32_t kMiss = 0x40000000u;         // OptionalId miss, lib/kdtree.h:156-161
constexpr int kStackDepth = 64;                 // >= tree height + 1 (checked at scene creation)
constexpr float kFltMax = 3.402823466e+38f;
constexpr float kCellSlack = 1e-4f;             // relative slack of the per-cell hit range (see traverse_pairs)
// lower end of a cell's hit range: a plane hit up to this far in front of the cell's entry is still accepted in the cell
// (rounding of r against the split-plane distances). Any rule that SKIPS a cell because it starts beyond a limit (the
// light of an any-hit query) must compare this bound, not tenter itself.
__host__ __device__ __forceinline__ float cell_lo(float tenter) { return tenter - kCellSlack * (fabsf(tenter) + 1.f); }

// ------------------------------------------------------------------ device scene
struct DevScene {
    const uint2* nodes;        // inner: (split bits, right<<2 | axis)   leaf-run head: (first ref, count<<2 | 3)
    const uint32_t* leaf_refs; // triangle ids of all leaf runs, in the reference's visiting order
    // intersection records (the reference's precomputed quantities, lib/triangle.h:33-39), split so that the part
    // every test reads stays L2-resident: hot = 2 x float4 per triangle: v0.xyz n.x | n.yz u.xy  (plane test + first
    // edge words), cold = 2 x float4: u.z v.xyz | uv vv uu denom (only for candidates that pass 0 <= r < best)
    const float4* isect_hot;
    const float4* isect_cold;
    // bounding box of each triangle, grown by 1e-4 x scene scale: 2 x float4 (lo.xyz -, hi.xyz -). Used by the one-pass
    // leaf schedule (big leaves): a plane hit outside the box cannot pass the barycentric test.
    const float4* tri_box;
    const float4* shade;       // 4 x float4 per triangle: n0.xyz n1.x | n1.yz n2.xy | n2.z reflectivity - - | rgba
    const float4* mirror;      // reflective rgba per triangle (raytracer integrator only)
    float lo[3], hi[3];        // KDTree::box()
    // production layout: sibling pairs + empty-space cuts (kdtree_build.h); `nodes`/`leaf_refs` above hold the
    // reference-shaped tree and are uploaded only for the instrumented reference-schedule twin
    const uint2* pnodes;
    const uint32_t* prefs;
    uint32_t treelet_pairs; // node pairs of the breadth-first top of pnodes (<= KdTree::kTreeletNodes / 2)
    // 1: every ray takes the reference's schedule verbatim (negative tenter, solid side of every cut, no early exit, no
    // per-cell hit range). Set for scenes smaller than 0.1 units: the reference's builder clips with an ABSOLUTE thickness
    // (EPS = 1e-5, lib/clipping.h:135-187, lib/types.h:13), which at that size is no longer small against the relative 1e-4
    // slack the shortcuts rely on -- its tree then holds triangles in cells they do not overlap and misses them in cells they
    // do, and what the reference finds depends on every cell it visits (tools/diag_tiny.py: differences appear below ~0.02).
    uint32_t verbatim;
};

// ray wave, SoA of float4 (fully coalesced 16-byte lanes)
struct RayWave {
    float4* a; // o.x o.y o.z d.x
    float4* b; // d.y d.z  rel(u32: primary-sample index relative to the batch)  node(u32: index in the m-ary ray tree)
    float4* T; // rgba throughput
};
struct ShadowWave {
    float4* a; // o.x o.y o.z d.x
    float4* b; // d.y d.z tmax pixel(u32)
    float4* c; // rgba contribution to add when unoccluded
};

struct CameraDev {
    float pos[3];
    float rot[9];
    float delta_x, delta_y;
};

struct FrameParams {
    CameraDev cam;
    int32_t width, height;
    int32_t pps;           // --pixel-samples of the whole job (RNG key + jitter table stride)
    int32_t n_local;       // pixel samples this call renders per pixel
    int32_t sample_begin, sample_stride;
    int32_t mc_samples, max_depth;
    float bg[4];
    int32_t has_light;
    float light_pos[3];
    float light_rgba[4];
    float max_visibility;
    float shadow_intensity;
    uint64_t seed;
    int32_t child_major; // next-wave layout inside a warp's block: 1 = the m children of a hit adjacent, 0 = k-major
};

// -------------------------------------------------------------------------- RNG
// xorshift64* (lib/xorshift.h:36-41) and its float conversion (lib/xorshift.h:57-59)
__device__ __forceinline__ uint64_t xs64star_next(uint64_t& s) {
    s ^= s >> 12;
    s ^= s << 25;
    s ^= s >> 27;
    return s * 2685821657736338717ULL;
}
__device__ __forceinline__ float xs64star_float(uint64_t& s) {
    return static_cast<float>(xs64star_next(s) & 0xFFFFFFull) * 5.9604644775390625e-8f; // ldexp(x, -24), exact
}
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// counter-seeded stream of one ray-tree node (DESIGN.md "RNG"); replaces sampling.h:12's sequential stream
__host__ __device__ __forceinline__ uint64_t node_state(uint64_t seed_mix, uint64_t sample_index, uint64_t node) {
    uint64_t s = splitmix64(splitmix64(seed_mix ^ splitmix64(sample_index)) + node);
    return s ? s : 0x9E3779B97F4A7C15ull;
}

// ------------------------------------------------------------------- traversal
struct HitRec {
    uint32_t id;
    float r, s, t;
};

__device__ __forceinline__ float sel3(int ax, float x, float y, float z) { return ax == 0 ? x : (ax == 1 ? y : z); }

// Front-to-back kd traversal (Hapala-Havran Alg. 2 as in lib/kdtree.cpp:515-578) with the
// stack in local memory. Differences from the reference, none of which can change the
// result (pinned in tests): tenter is clamped to 0 and the walk stops once the
// best hit lies inside the current cell instead of draining the stack.
// ANY_HIT: stop at the first accepted triangle with r <= tmax (shadow predicate,
// pathtracer.cpp:53: lit <=> !hit || r_closest > dist_to_light).
// COUNT: also tally inner-node visits, reference leaf nodes (8-byte id pairs, lib/kdtree.h:62-154) and
// triangle tests -- the n_* of SURVEY 8(d)'s algorithmic-bytes formula, same schedule as the oracle's
// instrumented early-exit mode (tests/test_gpu_parity.py::test_traversal_counters_equal_oracle).
struct VisitCounts {
    uint32_t inner, leaf_nodes, tri_tests;
};

template <bool ANY_HIT, bool COUNT = false>
__device__ __forceinline__ bool traverse(const DevScene& sc, float ox, float oy, float oz, float dx, float dy, float dz,
                                         float tmax_any, HitRec& out, VisitCounts* vc = nullptr) {
    const float fdx = dx == 0.f ? kEpsDir : dx;
    const float fdy = dy == 0.f ? kEpsDir : dy;
    const float fdz = dz == 0.f ? kEpsDir : dz;
    const float ix = 1 / fdx, iy = 1 / fdy, iz = 1 / fdz;

    // intersect_ray_box, lib/intersection.h:105-128
    float tx1 = (sc.lo[0] - ox) * ix, tx2 = (sc.hi[0] - ox) * ix;
    float tenter = fminf(tx1, tx2), texit = fmaxf(tx1, tx2);
    float ty1 = (sc.lo[1] - oy) * iy, ty2 = (sc.hi[1] - oy) * iy;
    tenter = fmaxf(tenter, fminf(ty1, ty2));
    texit = fminf(texit, fmaxf(ty1, ty2));
    float tz1 = (sc.lo[2] - oz) * iz, tz2 = (sc.hi[2] - oz) * iz;
    tenter = fmaxf(tenter, fminf(tz1, tz2));
    texit = fminf(texit, fmaxf(tz1, tz2));

    out.id = kMiss;
    out.r = kFltMax;
    out.s = 0.f;
    out.t = 0.f;
    if (texit < tenter) return false;
    if (tenter < 0.f) tenter = 0.f;

    uint32_t stk_node[kStackDepth];
    float stk_tmin[kStackDepth];
    float stk_tmax[kStackDepth];
    int sp = 0;
    uint32_t node = 0;

    for (;;) {
        uint2 n = __ldg(&sc.nodes[node]);
        while ((n.y & 3u) != 3u) {
            if (COUNT) vc->inner += 1;
            const int ax = static_cast<int>(n.y & 3u);
            const float split = __uint_as_float(n.x);
            const float o_ax = sel3(ax, ox, oy, oz);
            const float i_ax = sel3(ax, ix, iy, iz);
            const float d_ax = sel3(ax, fdx, fdy, fdz);
            const float t = (split - o_ax) * i_ax;
            uint32_t near = node + 1, far = n.y >> 2;
            if (d_ax <= 0.f) {
                const uint32_t tmp = near;
                near = far;
                far = tmp;
            }
            if (texit < t) {
                node = near;
            } else if (t < tenter) {
                node = far;
            } else {
                stk_node[sp] = far;
                stk_tmin[sp] = t;
                stk_tmax[sp] = texit;
                ++sp;
                node = near;
                texit = t;
            }
            n = __ldg(&sc.nodes[node]);
        }

        // leaf run (lib/kdtree.cpp:580-607): strict '<' keeps the first-visited triangle on ties
        const uint32_t first = n.x, count = n.y >> 2;
        if (COUNT) vc->leaf_nodes += (count + 1) >> 1;
        for (uint32_t i = 0; i < count; ++i) {
            if (COUNT) vc->tri_tests += 1;
            const uint32_t id = __ldg(&sc.leaf_refs[first + i]);
            const float4* rec = sc.isect_hot + 2 * static_cast<size_t>(id);
            const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1);
            // intersect_ray_plane, lib/intersection.h:40-49
            const float nx = q0.w, ny = q1.x, nz = q1.y;
            const float denom = nx * dx + ny * dy + nz * dz;
            if (denom == 0.f) continue;
            const float nom = nx * (q0.x - ox) + ny * (q0.y - oy) + nz * (q0.z - oz);
            const float r = nom / denom;
            // r < 0 rejects (intersection.h:66); a hit only matters if it beats the running
            // minimum (kdtree.cpp:591,569), resp. lies within tmax for the shadow predicate
            if (!(r >= 0.f)) continue;
            if (ANY_HIT ? !(r <= tmax_any) : !(r < out.r)) continue;
            const float4* rec2 = sc.isect_cold + 2 * static_cast<size_t>(id);
            const float4 q2 = __ldg(rec2), q3 = __ldg(rec2 + 1);
            // lib/intersection.h:70-86
            const float wx = (ox + r * dx) - q0.x, wy = (oy + r * dy) - q0.y, wz = (oz + r * dz) - q0.z;
            const float ux = q1.z, uy = q1.w, uz = q2.x, vx = q2.y, vy = q2.z, vz = q2.w;
            const float wv = wx * vx + wy * vy + wz * vz;
            const float wu = wx * ux + wy * uy + wz * uz;
            const float s = (q3.x * wv - q3.y * wu) / q3.w;
            if (s < 0.f) continue;
            const float t = (q3.x * wu - q3.z * wv) / q3.w;
            if (t < 0.f || 1.f < s + t) continue;
            out.id = id;
            out.r = r;
            out.s = s;
            out.t = t;
            if (ANY_HIT) return true;
        }

        if (out.id != kMiss && out.r <= texit) break;
        if (sp == 0) break;
        --sp;
        node = stk_node[sp];
        tenter = stk_tmin[sp];
        texit = stk_tmax[sp];
        // the stack is ordered front to back: once the nearest pending cell starts beyond the
        // light, nothing behind it can hold an occluder with r <= tmax
        if (ANY_HIT && tenter > tmax_any) break;
    }
    return out.id != kMiss;
}

// Production traversal: same visiting order and per-triangle arithmetic as traverse<> above, on the
// sibling-pair layout with the empty-space cuts kept (kdtree_build.h). One 16-byte load brings both children;
// the far child's node word travels on the stack, so a pop needs no node load; cut-off voids are never pushed.
// A void only removes leaves the ray segment does not pass through, which cannot hold the closest hit
// (tests/test_gpu_parity.py compares against the reference's exhaustive schedule).
template <bool ANY_HIT, bool COUNT = false>
__device__ __forceinline__ bool traverse_pairs(const DevScene& sc, float ox, float oy, float oz, float dx, float dy, float dz,
                                               float tmax_any, HitRec& out, VisitCounts* vc = nullptr) {
    const float fdx = dx == 0.f ? kEpsDir : dx;
    const float fdy = dy == 0.f ? kEpsDir : dy;
    const float fdz = dz == 0.f ? kEpsDir : dz;
    const float ix = 1 / fdx, iy = 1 / fdy, iz = 1 / fdz;
    const bool axis_parallel = dx == 0.f || dy == 0.f || dz == 0.f || sc.verbatim != 0u;

    float tx1 = (sc.lo[0] - ox) * ix, tx2 = (sc.hi[0] - ox) * ix;
    float tenter = fminf(tx1, tx2), texit = fmaxf(tx1, tx2);
    float ty1 = (sc.lo[1] - oy) * iy, ty2 = (sc.hi[1] - oy) * iy;
    tenter = fmaxf(tenter, fminf(ty1, ty2));
    texit = fminf(texit, fmaxf(ty1, ty2));
    float tz1 = (sc.lo[2] - oz) * iz, tz2 = (sc.hi[2] - oz) * iz;
    tenter = fmaxf(tenter, fminf(tz1, tz2));
    texit = fminf(texit, fmaxf(tz1, tz2));

    out.id = kMiss;
    out.r = kFltMax;
    out.s = 0.f;
    out.t = 0.f;
    if (texit < tenter) return false;
    // Rays with an exact-zero direction component (and every ray of a `verbatim` scene) are traversed with the "fixed" direction of lib/kdtree.cpp:503-511 but
    // tested with the real one: their interval is not the real ray's, and what the reference finds depends on every cell
    // it happens to visit. They take the reference's schedule verbatim: negative tenter, solid side of every cut, no
    // early exit, no per-cell hit range.
    if (tenter < 0.f && !axis_parallel) tenter = 0.f;

    uint4 stack[kStackDepth]; // (node.x, node.y, tmin, tmax)
    int sp = 0;
    uint2 n = __ldg(&sc.pnodes[0]);

    for (;;) {
        while ((n.y & 3u) != 3u) {
            if (COUNT) vc->inner += 1;
            const int ax = static_cast<int>(n.y & 3u);
            const float split = __uint_as_float(n.x);
            const uint4 pair = __ldg(reinterpret_cast<const uint4*>(sc.pnodes + (n.y >> 2)));
            const float o_ax = sel3(ax, ox, oy, oz);
            const float i_ax = sel3(ax, ix, iy, iz);
            const float t = (split - o_ax) * i_ax;
            if (axis_parallel && (pair.y == 3u || pair.w == 3u)) { // cut node: the reference has no plane here
                n = pair.y == 3u ? make_uint2(pair.z, pair.w) : make_uint2(pair.x, pair.y);
                continue;
            }
            // left is near unless fixed_ray.d[ax] <= 0 (lib/kdtree.cpp:549-553); sign(d) == sign(1/d), d != 0
            const bool flip = (__float_as_uint(i_ax) >> 31) != 0u;
            const uint2 near = flip ? make_uint2(pair.z, pair.w) : make_uint2(pair.x, pair.y);
            const uint2 far = flip ? make_uint2(pair.x, pair.y) : make_uint2(pair.z, pair.w);
            // lib/kdtree.cpp:555-563 without divergent branches: near only | far only | both (far pushed); cut-off voids
            // (count-0 leaves, y == 3) are neither entered nor pushed
            const bool near_only = texit < t;
            const bool far_only = !near_only && (t < tenter);
            const bool both = !near_only && !far_only;
            const bool go_far = far_only || (both && near.y == 3u);
            if (both && near.y != 3u && far.y != 3u) stack[sp++] = make_uint4(far.x, far.y, __float_as_uint(t), __float_as_uint(texit));
            n = go_far ? far : near;
            tenter = (both && go_far) ? t : tenter;
            texit = (both && !go_far) ? t : texit;
        }

        const uint32_t first = n.x, count = n.y >> 2;
        if (COUNT) vc->leaf_nodes += (count + 1) >> 1;
        // Only plane hits inside this cell's parameter range (plus slack) go on to the barycentric part: a hit of
        // this triangle outside the cell is found again in the cell that contains it (the triangle is referenced
        // there too), so nothing is lost and ~4 of 5 barycentric evaluations are saved.
        // Not for rays with an exact-zero direction component: their traversal interval belongs to the "fixed"
        // direction (lib/kdtree.cpp:503-511), not to the ray the triangles are tested with.
        const float r_lo = axis_parallel ? -kFltMax : tenter - kCellSlack * (fabsf(tenter) + 1.f);
        const float r_hi = axis_parallel ? kFltMax : texit + kCellSlack * (fabsf(texit) + 1.f);
        for (uint32_t i = 0; i < count; ++i) {
            if (COUNT) vc->tri_tests += 1;
            const uint32_t id = __ldg(&sc.prefs[first + i]);
            const float4* rec = sc.isect_hot + 2 * static_cast<size_t>(id);
            const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1);
            const float nx = q0.w, ny = q1.x, nz = q1.y;
            const float denom = nx * dx + ny * dy + nz * dz; // lib/intersection.h:40-49
            if (denom == 0.f) continue;
            const float nom = nx * (q0.x - ox) + ny * (q0.y - oy) + nz * (q0.z - oz);
            const float r = nom / denom;
            if (!(r >= 0.f)) continue;
            if (ANY_HIT ? !(r <= tmax_any) : !(r < out.r)) continue;
            if (!(r >= r_lo && r <= r_hi)) continue;
            if (count > 8u) { // big leaf (a scene that is one leaf): most surviving plane hits lie far outside their triangle
                const float4 blo = __ldg(sc.tri_box + 2 * static_cast<size_t>(id)), bhi = __ldg(sc.tri_box + 2 * static_cast<size_t>(id) + 1);
                const float hx = ox + r * dx, hy = oy + r * dy, hz = oz + r * dz;
                if (hx < blo.x || hy < blo.y || hz < blo.z || hx > bhi.x || hy > bhi.y || hz > bhi.z) continue;
            }
            const float4* rec2 = sc.isect_cold + 2 * static_cast<size_t>(id);
            const float4 q2 = __ldg(rec2), q3 = __ldg(rec2 + 1);
            const float wx = (ox + r * dx) - q0.x, wy = (oy + r * dy) - q0.y, wz = (oz + r * dz) - q0.z; // :70-71
            const float ux = q1.z, uy = q1.w, uz = q2.x, vx = q2.y, vy = q2.z, vz = q2.w;
            const float wv = wx * vx + wy * vy + wz * vz;
            const float wu = wx * ux + wy * uy + wz * uz;
            const float s = (q3.x * wv - q3.y * wu) / q3.w; // :78-86
            if (s < 0.f) continue;
            const float t = (q3.x * wu - q3.z * wv) / q3.w;
            if (t < 0.f || 1.f < s + t) continue;
            out.id = id;
            out.r = r;
            out.s = s;
            out.t = t;
            if (ANY_HIT) return true;
        }

        if (!axis_parallel && out.id != kMiss && out.r <= texit) break;
        if (sp == 0) break;
        const uint4 e = stack[--sp];
        n = make_uint2(e.x, e.y);
        tenter = __uint_as_float(e.z);
        texit = __uint_as_float(e.w);
        // (pruned with the same slack the cell's hit range has: a hit within rounding noise in front of the cell's entry is
        // accepted there, so the cell may only be skipped when even that lower bound lies beyond the light)
        if (ANY_HIT && !axis_parallel && cell_lo(tenter) > tmax_any) break;
    }
    return out.id != kMiss;
}

// ---------------------------------------------------------------------- kernels

__device__ __forceinline__ uint32_t pixel_of(const FrameParams& fp, uint64_t first_local_index, uint32_t rel,
                                             uint32_t& sample_i) {
    const uint64_t gl = first_local_index + rel;
    uint32_t pixel, j;
    if ((gl >> 32) == 0) { // (a 64-bit division costs ~100 instructions: 9 % of the last wave's shading)
        const uint32_t g = static_cast<uint32_t>(gl), nl = static_cast<uint32_t>(fp.n_local);
        pixel = g / nl;
        j = g - pixel * nl;
    } else {
        pixel = static_cast<uint32_t>(gl / static_cast<uint32_t>(fp.n_local));
        j = static_cast<uint32_t>(gl % static_cast<uint32_t>(fp.n_local));
    }
    sample_i = fp.sample_begin + j * fp.sample_stride;
    return pixel;
}

// raygen: one thread per primary sample of the batch. Batch-relative index rel -> (pixel, local sample j) ->
// jittered raster position -> Camera::raster2cam (lib/types.h:119-123). Jitter = the reference's per-row
// xorshift64star<float>(42) stream (main.cpp:201-206), precomputed on the host as a [width][pps][2] table
// because every row restarts it.
__global__ void __launch_bounds__(256) raygen_kernel(FrameParams fp, const float2* __restrict__ jitter,
                                                     uint64_t first_local_index, uint32_t count, RayWave out) {
    const uint32_t rel = blockIdx.x * blockDim.x + threadIdx.x;
    if (rel >= count) return;
    uint32_t i;
    const uint32_t pixel = pixel_of(fp, first_local_index, rel, i);
    const int x = static_cast<int>(pixel % static_cast<uint32_t>(fp.width));
    const int y = static_cast<int>(pixel / static_cast<uint32_t>(fp.width));
    const float2 jit = __ldg(&jitter[static_cast<size_t>(x) * fp.pps + i]);
    const float px = x + jit.x, py = y + jit.y;
    const float w = static_cast<float>(fp.width), h = static_cast<float>(fp.height);
    const float vx = -fp.cam.delta_x * (1 - 2 * px / w);
    const float vy = fp.cam.delta_y * (1 - 2 * py / h);
    const float vz = -1.f;
    const float* m = fp.cam.rot;
    const float dx = m[0] * vx + m[1] * vy + m[2] * vz;
    const float dy = m[3] * vx + m[4] * vy + m[5] * vz;
    const float dz = m[6] * vx + m[7] * vy + m[8] * vz;
    // waves are streamed once: evict-first stores/loads (.cs) keep the scene arrays resident in L2
    __stcs(&out.a[rel], make_float4(fp.cam.pos[0], fp.cam.pos[1], fp.cam.pos[2], dx));
    __stcs(&out.b[rel], make_float4(dy, dz, __uint_as_float(rel), __uint_as_float(0u)));
    __stcs(&out.T[rel], make_float4(1.f, 1.f, 1.f, 1.f));
}

// closest hit for a wave of rays; one thread per ray
__global__ void __launch_bounds__(128) trace_closest_kernel(DevScene sc, const float4* __restrict__ ra,
                                                            const float4* __restrict__ rb, uint32_t count,
                                                            uint4* __restrict__ hits) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    const float4 a = __ldcs(&ra[idx]);
    const float4 b = __ldcs(&rb[idx]);
    HitRec h;
    traverse_pairs<false>(sc, a.x, a.y, a.z, a.w, b.x, b.y, 0.f, h);
    __stcs(&hits[idx], make_uint4(h.id, __float_as_uint(h.r), __float_as_uint(h.s), __float_as_uint(h.t)));
}

// plain (o, d) arrays in, for trn_intersect
__global__ void __launch_bounds__(128) trace_closest_plain_kernel(DevScene sc, const float* __restrict__ o,
                                                                  const float* __restrict__ d, uint32_t count,
                                                                  uint4* __restrict__ hits) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    HitRec h;
    traverse_pairs<false>(sc, o[3 * idx], o[3 * idx + 1], o[3 * idx + 2], d[3 * idx], d[3 * idx + 1], d[3 * idx + 2], 0.f, h);
    hits[idx] = make_uint4(h.id, __float_as_uint(h.r), __float_as_uint(h.s), __float_as_uint(h.t));
}

// instrumented twins of the two traversal kernels (bench.py's roofline leg and the counter parity test only)
__device__ __forceinline__ void flush_counts(const VisitCounts& vc, unsigned long long* g) {
    unsigned a = vc.inner, b = vc.leaf_nodes, c = vc.tri_tests;
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, off);
        b += __shfl_down_sync(0xffffffffu, b, off);
        c += __shfl_down_sync(0xffffffffu, c, off);
    }
    if ((threadIdx.x & 31u) == 0) {
        atomicAdd(g + 0, static_cast<unsigned long long>(a));
        atomicAdd(g + 1, static_cast<unsigned long long>(b));
        atomicAdd(g + 2, static_cast<unsigned long long>(c));
    }
}

__global__ void __launch_bounds__(128) trace_closest_count_kernel(DevScene sc, const float4* __restrict__ ra,
                                                                  const float4* __restrict__ rb, uint32_t count,
                                                                  uint4* __restrict__ hits, unsigned long long* g) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    VisitCounts vc{0, 0, 0}, vp{0, 0, 0};
    if (idx < count) {
        const float4 a = ra[idx];
        const float4 b = rb[idx];
        HitRec h;
        traverse<false, true>(sc, a.x, a.y, a.z, a.w, b.x, b.y, 0.f, h, &vc);       // reference-shaped schedule: defines B_alg
        traverse_pairs<false, true>(sc, a.x, a.y, a.z, a.w, b.x, b.y, 0.f, h, &vp); // what the production kernel really visits
        hits[idx] = make_uint4(h.id, __float_as_uint(h.r), __float_as_uint(h.s), __float_as_uint(h.t));
    }
    flush_counts(vc, g);
    flush_counts(vp, g + 6);
}

__global__ void __launch_bounds__(128) trace_closest_plain_count_kernel(DevScene sc, const float* __restrict__ o,
                                                                        const float* __restrict__ d, uint32_t count,
                                                                        uint4* __restrict__ hits, unsigned long long* g) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    VisitCounts vc{0, 0, 0}, vp{0, 0, 0};
    if (idx < count) {
        HitRec h;
        traverse<false, true>(sc, o[3 * idx], o[3 * idx + 1], o[3 * idx + 2], d[3 * idx], d[3 * idx + 1], d[3 * idx + 2], 0.f, h, &vc);
        traverse_pairs<false, true>(sc, o[3 * idx], o[3 * idx + 1], o[3 * idx + 2], d[3 * idx], d[3 * idx + 1], d[3 * idx + 2], 0.f, h, &vp);
        hits[idx] = make_uint4(h.id, __float_as_uint(h.r), __float_as_uint(h.s), __float_as_uint(h.t));
    }
    flush_counts(vc, g);
    flush_counts(vp, g + 6);
}

struct WaveCounters {
    uint32_t next_count;   // rays appended to the next wave
    uint32_t shadow_count; // shadow rays appended
    uint32_t trace_cursor;  // work cursors of the persistent traversal kernels working on this chunk
    uint32_t shadow_cursor;
};

__device__ __forceinline__ void accumulate(float4* acc, uint32_t pixel, float4 v) {
    if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) return;
    atomicAdd(acc + pixel, v); // red.global.add.v4.f32 (sm_90+)
}

// shade/bounce (pathtracer.cpp:26-101 in throughput form): per ray of the wave
//   miss            -> acc[pixel] += T * bg
//   hit             -> shadow ray with pre-weighted contribution T * rho * max(0, n.l) * light / pi
//                      and, below max depth, m uniform-hemisphere children with T' = T * rho * (2 cos / m)
// The next wave is built compacted: per warp one ballot/popc and one atomicAdd reserve the slots
// (children laid out k-major inside the warp's block so that every store is a coalesced run).
__global__ void __launch_bounds__(256) shade_bounce_kernel(DevScene sc, FrameParams fp, uint64_t first_local_index,
                                                           RayWave cur, const uint4* __restrict__ hits, uint32_t count,
                                                           int depth, RayWave next, ShadowWave shadow,
                                                           WaveCounters* __restrict__ counters,
                                                           float4* __restrict__ acc) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const bool valid = idx < count;
    bool is_hit = false;
    float4 a, b, T;
    uint4 h;
    uint32_t rel = 0, node = 0, pixel = 0, sample_i = 0;
    if (valid) {
        a = __ldcs(&cur.a[idx]);
        b = __ldcs(&cur.b[idx]);
        T = __ldcs(&cur.T[idx]);
        h = __ldcs(&hits[idx]);
        rel = __float_as_uint(b.z);
        node = __float_as_uint(b.w);
        pixel = pixel_of(fp, first_local_index, rel, sample_i);
        is_hit = h.x != kMiss;
        if (!is_hit) accumulate(acc, pixel, make_float4(T.x * fp.bg[0], T.y * fp.bg[1], T.z * fp.bg[2], T.w * fp.bg[3]));
    }

    float nx = 0, ny = 0, nz = 0, p2x = 0, p2y = 0, p2z = 0;
    float4 rho = make_float4(0, 0, 0, 0);
    bool want_shadow = false;
    float ldx = 0, ldy = 0, ldz = 0, ldist = 0;
    float4 contrib = make_float4(0, 0, 0, 0);
    if (is_hit) {
        const float r = __uint_as_float(h.y), s = __uint_as_float(h.z), t = __uint_as_float(h.w);
        const float dx = a.w, dy = b.x, dz = b.y;
        const float px = a.x + r * dx, py = a.y + r * dy, pz = a.z + r * dz; // pathtracer.cpp:30
        const float4* srec = sc.shade + 4 * static_cast<size_t>(h.x);
        const float4 s0 = __ldg(srec), s1 = __ldg(srec + 1), s2 = __ldg(srec + 2);
        rho = __ldg(srec + 3);
        // Triangle::interpolate_normal(1-s-t, s, t), lib/triangle.h:54-56
        const float br = 1.f - s - t;
        float mx = (br * s0.x + s * s0.w) + t * s1.z;
        float my = (br * s0.y + s * s1.x) + t * s1.w;
        float mz = (br * s0.z + s * s1.y) + t * s2.x;
        const float inv = 1 / sqrtf(mx * mx + my * my + mz * mz);
        nx = inv * mx;
        ny = inv * my;
        nz = inv * mz;
        p2x = px + 0.0001f * nx; // pathtracer.cpp:36
        p2y = py + 0.0001f * ny;
        p2z = pz + 0.0001f * nz;
        if (fp.has_light) { // pathtracer.cpp:44-58
            float lx = fp.light_pos[0] - px, ly = fp.light_pos[1] - py, lz = fp.light_pos[2] - pz;
            const float linv = 1 / sqrtf(lx * lx + ly * ly + lz * lz);
            ldx = linv * lx;
            ldy = linv * ly;
            ldz = linv * lz;
            const float qx = fp.light_pos[0] - p2x, qy = fp.light_pos[1] - p2y, qz = fp.light_pos[2] - p2z;
            ldist = sqrtf(qx * qx + qy * qy + qz * qz);
            const float wgt = fmaxf(0.f, ldx * nx + ldy * ny + ldz * nz);
            const float k = 0.318309886183790671538f; // M_1_PI as float
            contrib = make_float4(T.x * (rho.x * (k * (wgt * fp.light_rgba[0]))), T.y * (rho.y * (k * (wgt * fp.light_rgba[1]))),
                                  T.z * (rho.z * (k * (wgt * fp.light_rgba[2]))), T.w * (rho.w * (k * (wgt * fp.light_rgba[3]))));
            want_shadow = !(contrib.x == 0.f && contrib.y == 0.f && contrib.z == 0.f && contrib.w == 0.f);
        }
    }

    // ---- compaction: slots of the shadow rays and of the children in the next wave. Ballot/popc inside the warp, one
    // shared-memory exchange per CTA, then ONE atomicAdd per queue per CTA (8x fewer same-address atomics than per warp:
    // on one-leaf scenes the shade kernel runs 20 M warps per step and the queue counters serialise in L2).
    const bool spawn = is_hit && depth < fp.max_depth;
    const int m = fp.mc_samples;
    const unsigned smask = __ballot_sync(0xffffffffu, want_shadow);
    const unsigned hmask = __ballot_sync(0xffffffffu, spawn);
    const uint32_t nh = __popc(hmask);
    __shared__ uint32_t s_shadow[8], s_child[8], s_base[2];
    const unsigned warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_shadow[warp] = __popc(smask);
        s_child[warp] = nh * static_cast<uint32_t>(m);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t ts = 0, tc = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const uint32_t a_ = s_shadow[w], b_ = s_child[w];
            s_shadow[w] = ts; // exclusive prefix inside the CTA
            s_child[w] = tc;
            ts += a_;
            tc += b_;
        }
        s_base[0] = ts ? atomicAdd(&counters->shadow_count, ts) : 0u;
        s_base[1] = tc ? atomicAdd(&counters->next_count, tc) : 0u;
    }
    __syncthreads();
    if (want_shadow) {
        const uint32_t slot = s_base[0] + s_shadow[warp] + __popc(smask & ((1u << lane) - 1u));
        __stcs(&shadow.a[slot], make_float4(p2x, p2y, p2z, ldx));
        __stcs(&shadow.b[slot], make_float4(ldy, ldz, ldist, __uint_as_float(pixel)));
        __stcs(&shadow.c[slot], contrib);
    }
    if (!spawn) return;
    const uint32_t cbase = s_base[1] + s_child[warp];
    const uint32_t rank = __popc(hmask & ((1u << lane) - 1u));

    // aiMatrix3x3::FromToMatrix((0,0,1) -> normal), assimp matrix3x3.inl (Moeller-Hughes), pathtracer.cpp:68-70
    float m00, m01, m02, m10, m11, m12, m20, m21, m22;
    {
        const float fx = 0.f, fy = 0.f, fz = 1.f;
        const float e = fx * nx + fy * ny + fz * nz;
        const float f = e < 0.f ? -e : e;
        if (f > 1.0f - 0.00001f) {
            // x = axis "most nearly orthogonal" to from=(0,0,1) by assimp's tie rules -> (0,1,0)
            const float xx = 0.f, xy = 1.f, xz = 0.f;
            const float ux = xx - fx, uy = xy - fy, uz = xz - fz;
            const float vx = xx - nx, vy = xy - ny, vz = xz - nz;
            const float c1 = 2.0f / (ux * ux + uy * uy + uz * uz);
            const float c2 = 2.0f / (vx * vx + vy * vy + vz * vz);
            const float c3 = c1 * c2 * (ux * vx + uy * vy + uz * vz);
            const float u[3] = {ux, uy, uz}, v[3] = {vx, vy, vz};
            float mt[3][3];
#pragma unroll
            for (int i = 0; i < 3; i++) {
#pragma unroll
                for (int j = 0; j < 3; j++) mt[i][j] = -c1 * u[i] * u[j] - c2 * v[i] * v[j] + c3 * v[i] * u[j];
                mt[i][i] += 1.0f;
            }
            m00 = mt[0][0]; m01 = mt[0][1]; m02 = mt[0][2];
            m10 = mt[1][0]; m11 = mt[1][1]; m12 = mt[1][2];
            m20 = mt[2][0]; m21 = mt[2][1]; m22 = mt[2][2];
        } else {
            const float vx = fy * nz - fz * ny, vy = fz * nx - fx * nz, vz = fx * ny - fy * nx;
            const float hh = 1.0f / (1.0f + e);
            const float hvx = hh * vx, hvz = hh * vz;
            const float hvxy = hvx * vy, hvxz = hvx * vz, hvyz = hvz * vy;
            m00 = e + hvx * vx; m01 = hvxy - vz;        m02 = hvxz + vy;
            m10 = hvxy + vz;    m11 = e + hh * vy * vy; m12 = hvyz - vx;
            m20 = hvxz - vy;    m21 = hvyz + vx;        m22 = e + hvz * vz;
        }
    }

    const uint64_t sample_index = static_cast<uint64_t>(pixel) * static_cast<uint64_t>(fp.pps) + sample_i;
    const float fm = static_cast<float>(m);
    // x / m == x * (1 / m) bit for bit when m is a power of two (-m 4 of every BASELINE config); an IEEE division is ~12 instructions
    const bool m_pow2 = (m & (m - 1)) == 0;
    const float inv_m = 1.f / fm;
    for (int k = 0; k < m; ++k) {
        const uint32_t child = node * static_cast<uint32_t>(m) + static_cast<uint32_t>(k) + 1u;
        uint64_t st = node_state(fp.seed, sample_index, child);
        const float u1 = xs64star_float(st);
        const float u2 = xs64star_float(st);
        // sampling::hemisphere, lib/sampling.h:20-32
        const float z = u1;
        const float rr = sqrtf(fmaxf(0.f, 1.f - z * z));
        const float phi = 6.28318530717958647692f * u2;
        float sphi, cphi;
        sincosf(phi, &sphi, &cphi); // one shared range reduction; same accuracy class as sinf/cosf (not the fast intrinsic)
        const float lx = rr * cphi, ly = rr * sphi, lz = z;
        const float dx = m00 * lx + m01 * ly + m02 * lz;
        const float dy = m10 * lx + m11 * ly + m12 * lz;
        const float dz = m20 * lx + m21 * ly + m22 * lz;
        const float wgt = m_pow2 ? (2.f * u1) * inv_m : (2.f * u1) / fm; // pathtracer.cpp:84,88,100-101: rho * 2 * (cos / m)
        // k-major: every store of the loop is one contiguous run; child-major: the m rays that share an origin sit in
        // adjacent lanes of the next wave, so their walks down the tree touch the same lines (profiles/README.md)
        const uint32_t slot = fp.child_major ? cbase + rank * static_cast<uint32_t>(m) + static_cast<uint32_t>(k)
                                             : cbase + static_cast<uint32_t>(k) * nh + rank;
        __stcs(&next.a[slot], make_float4(p2x, p2y, p2z, dx));
        __stcs(&next.b[slot], make_float4(dy, dz, __uint_as_float(rel), __uint_as_float(child)));
        __stcs(&next.T[slot], make_float4(T.x * (rho.x * wgt), T.y * (rho.y * wgt), T.z * (rho.z * wgt), T.w * (rho.w * wgt)));
    }
}

// raycaster.cpp:7-24: colour = diffuse rgb, alpha = clamp(1 - r / max_visibility); bg on a miss
__global__ void __launch_bounds__(256) shade_raycast_kernel(DevScene sc, FrameParams fp, uint64_t first_local_index,
                                                            RayWave cur, const uint4* __restrict__ hits, uint32_t count,
                                                            unsigned long long* __restrict__ hit_counter,
                                                            float4* __restrict__ acc) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = idx < count;
    bool is_hit = false;
    if (valid) {
        const uint4 h = hits[idx];
        const float4 b = cur.b[idx];
        uint32_t sample_i;
        const uint32_t pixel = pixel_of(fp, first_local_index, __float_as_uint(b.z), sample_i);
        is_hit = h.x != kMiss;
        float4 c;
        if (is_hit) {
            c = __ldg(sc.shade + 4 * static_cast<size_t>(h.x) + 3);
            const float al = 1.f - (__uint_as_float(h.y) / fp.max_visibility);
            c.w = al < 0.f ? 0.f : (1.f < al ? 1.f : al);
        } else {
            c = make_float4(fp.bg[0], fp.bg[1], fp.bg[2], fp.bg[3]);
        }
        accumulate(acc, pixel, c);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, is_hit);
    if ((threadIdx.x & 31u) == 0 && mask) atomicAdd(hit_counter, static_cast<unsigned long long>(__popc(mask)));
}

// any-hit shadow traversal; count is read from device memory (the wave was built by shade_bounce_kernel in the
// same stream), the grid covers the upper bound
__global__ void __launch_bounds__(128) trace_shadow_kernel(DevScene sc, ShadowWave sw,
                                                           const WaveCounters* __restrict__ counters,
                                                           float4* __restrict__ acc) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= counters->shadow_count) return;
    const float4 a = __ldcs(&sw.a[idx]);
    const float4 b = __ldcs(&sw.b[idx]);
    HitRec h;
    const bool occluded = traverse_pairs<true>(sc, a.x, a.y, a.z, a.w, b.x, b.y, b.z, h);
    if (!occluded) accumulate(acc, __float_as_uint(b.w), __ldcs(&sw.c[idx]));
}

__global__ void __launch_bounds__(128) trace_shadow_count_kernel(DevScene sc, ShadowWave sw,
                                                                 const WaveCounters* __restrict__ counters,
                                                                 float4* __restrict__ acc, unsigned long long* g) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    VisitCounts vc{0, 0, 0}, vp{0, 0, 0};
    if (idx < counters->shadow_count) {
        const float4 a = sw.a[idx];
        const float4 b = sw.b[idx];
        HitRec h;
        traverse<true, true>(sc, a.x, a.y, a.z, a.w, b.x, b.y, b.z, h, &vc);
        const bool occluded = traverse_pairs<true, true>(sc, a.x, a.y, a.z, a.w, b.x, b.y, b.z, h, &vp);
        if (!occluded && acc) accumulate(acc, __float_as_uint(b.w), sw.c[idx]); // acc == nullptr: the pooled twin adds it
    }
    flush_counts(vc, g);
    flush_counts(vp, g + 6);
}

// ------------------------------------------------------------------ raytracer integrator (raytracer.cpp:6-67)
// Throughput form of the recursion  color = F * ((1-k) * direct + k * R * trace(reflected)),  F = shadowed ? (x - x*sh) : x:
//   pixel += T * F * base,  base = k > 0 ? (1-k)*direct : direct;   child throughput T' = T * F * (k*R).
// The shade kernel emits the shadow query (carrying T*base and the slot of the child ray); the shadow kernel applies F
// to the contribution AND to the child's throughput before the child's own shading reads it (stream order).
__global__ void __launch_bounds__(256) shade_raytrace_kernel(DevScene sc, FrameParams fp, uint64_t first_local_index,
                                                             RayWave cur, const uint4* __restrict__ hits, uint32_t count,
                                                             int depth, RayWave next, ShadowWave shadow,
                                                             uint32_t* __restrict__ child_slot,
                                                             WaveCounters* __restrict__ counters,
                                                             unsigned long long* __restrict__ cut_rays,
                                                             float4* __restrict__ acc) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    bool is_hit = false, spawn = false, cut = false;
    float4 a, b, T = make_float4(0, 0, 0, 0), contrib = make_float4(0, 0, 0, 0), Tc = make_float4(0, 0, 0, 0);
    float p2x = 0, p2y = 0, p2z = 0, ldx = 0, ldy = 0, ldz = 0, ldist = 0, rx = 0, ry = 0, rz = 0;
    uint32_t rel = 0, pixel = 0;
    if (idx < count) {
        a = __ldcs(&cur.a[idx]);
        b = __ldcs(&cur.b[idx]);
        T = __ldcs(&cur.T[idx]);
        const uint4 h = __ldcs(&hits[idx]);
        rel = __float_as_uint(b.z);
        uint32_t sample_i;
        pixel = pixel_of(fp, first_local_index, rel, sample_i);
        is_hit = h.x != kMiss;
        if (!is_hit) {
            accumulate(acc, pixel, make_float4(T.x * fp.bg[0], T.y * fp.bg[1], T.z * fp.bg[2], T.w * fp.bg[3]));
        } else {
            const float r = __uint_as_float(h.y), s = __uint_as_float(h.z), t = __uint_as_float(h.w);
            const float dx = a.w, dy = b.x, dz = b.y;
            const float px = a.x + r * dx, py = a.y + r * dy, pz = a.z + r * dz;
            float lx = fp.light_pos[0] - px, ly = fp.light_pos[1] - py, lz = fp.light_pos[2] - pz;
            const float linv = 1 / sqrtf(lx * lx + ly * ly + lz * lz);
            lx = linv * lx; ly = linv * ly; lz = linv * lz;
            const float4* srec = sc.shade + 4 * static_cast<size_t>(h.x);
            const float4 s0 = __ldg(srec), s1 = __ldg(srec + 1), s2 = __ldg(srec + 2), rho = __ldg(srec + 3);
            const float br = 1.f - s - t;
            const float mx = (br * s0.x + s * s0.w) + t * s1.z;
            const float my = (br * s0.y + s * s1.x) + t * s1.w;
            const float mz = (br * s0.z + s * s1.y) + t * s2.x;
            const float ninv = 1 / sqrtf(mx * mx + my * my + mz * mz);
            const float nx = ninv * mx, ny = ninv * my, nz = ninv * mz;
            // lambertian(L, N, C, I), lib/lambertian.h:15-19
            const float ca = fmaxf(0.f, lx * nx + ly * ny + lz * nz);
            const float4 direct = make_float4((ca * rho.x) * fp.light_rgba[0], (ca * rho.y) * fp.light_rgba[1],
                                              (ca * rho.z) * fp.light_rgba[2], (ca * rho.w) * fp.light_rgba[3]);
            p2x = px + 0.0001f * nx; p2y = py + 0.0001f * ny; p2z = pz + 0.0001f * nz;
            const float k = s2.y;
            float4 base = direct;
            if (k > 0.f) { // raytracer.cpp:44-54
                const float4 R = __ldg(sc.mirror + h.x);
                const float two_nd = 2.f * (nx * dx + ny * dy + nz * dz);
                rx = dx - two_nd * nx; ry = dy - two_nd * ny; rz = dz - two_nd * nz;
                const float omk = 1.f - k;
                base = make_float4(omk * direct.x, omk * direct.y, omk * direct.z, omk * direct.w);
                Tc = make_float4(T.x * (k * R.x), T.y * (k * R.y), T.z * (k * R.z), T.w * (k * R.w));
                spawn = depth < fp.max_depth;
                cut = !spawn; // the reference still counts the call that the depth check rejects (raytracer.cpp:9-13)
            }
            contrib = make_float4(T.x * base.x, T.y * base.y, T.z * base.z, T.w * base.w);
            float qx = fp.light_pos[0] - p2x, qy = fp.light_pos[1] - p2y, qz = fp.light_pos[2] - p2z; // :57-59
            ldist = sqrtf(qx * qx + qy * qy + qz * qz);
            const float qinv = 1 / ldist;
            ldx = qinv * qx; ldy = qinv * qy; ldz = qinv * qz;
        }
    }
    // slots for the mirror children and the shadow queries: warp ballot/popc, one atomicAdd per queue per CTA
    const unsigned cmask = __ballot_sync(0xffffffffu, spawn);
    const unsigned kmask = __ballot_sync(0xffffffffu, cut);
    const unsigned smask = __ballot_sync(0xffffffffu, is_hit);
    __shared__ uint32_t s_child[8], s_shadow[8], s_cut[8], s_base[2];
    const unsigned warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_child[warp] = __popc(cmask);
        s_shadow[warp] = __popc(smask);
        s_cut[warp] = __popc(kmask);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tc = 0, ts = 0, tk = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const uint32_t c_ = s_child[w], s_ = s_shadow[w];
            s_child[w] = tc;
            s_shadow[w] = ts;
            tc += c_;
            ts += s_;
            tk += s_cut[w];
        }
        s_base[0] = tc ? atomicAdd(&counters->next_count, tc) : 0u;
        s_base[1] = ts ? atomicAdd(&counters->shadow_count, ts) : 0u;
        if (tk) atomicAdd(cut_rays, static_cast<unsigned long long>(tk));
    }
    __syncthreads();
    uint32_t cslot = 0xFFFFFFFFu;
    if (spawn) {
        cslot = s_base[0] + s_child[warp] + __popc(cmask & ((1u << lane) - 1u));
        __stcs(&next.a[cslot], make_float4(p2x, p2y, p2z, rx));
        __stcs(&next.b[cslot], make_float4(ry, rz, __uint_as_float(rel), __uint_as_float(0u)));
        next.T[cslot] = Tc; // re-read (and possibly scaled) by the shadow kernel: keep it cacheable
    }
    {
        if (is_hit) {
            const uint32_t slot = s_base[1] + s_shadow[warp] + __popc(smask & ((1u << lane) - 1u));
            // shadowed <=> closest hit strictly nearer than the light (raytracer.cpp:63): any-hit with r <= pred(dist)
            const float tmax = ldist > 0.f ? __uint_as_float(__float_as_uint(ldist) - 1u) : -1.f;
            __stcs(&shadow.a[slot], make_float4(p2x, p2y, p2z, ldx));
            __stcs(&shadow.b[slot], make_float4(ldy, ldz, tmax, __uint_as_float(pixel)));
            __stcs(&shadow.c[slot], contrib);
            child_slot[slot] = cslot;
        }
    }
}

__global__ void __launch_bounds__(128) trace_shadow_raytrace_kernel(DevScene sc, ShadowWave sw,
                                                                    const uint32_t* __restrict__ child_slot,
                                                                    const WaveCounters* __restrict__ counters,
                                                                    float shadow_intensity, float4* __restrict__ nextT,
                                                                    float4* __restrict__ acc) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= counters->shadow_count) return;
    const float4 a = __ldcs(&sw.a[idx]);
    const float4 b = __ldcs(&sw.b[idx]);
    HitRec h;
    const bool shadowed = traverse_pairs<true>(sc, a.x, a.y, a.z, a.w, b.x, b.y, b.z, h);
    float4 c = __ldcs(&sw.c[idx]);
    const uint32_t cs = child_slot[idx];
    if (shadowed) { // color -= color * shadow_intensity (raytracer.cpp:64-66), applied to both addends of color
        c = make_float4(c.x - shadow_intensity * c.x, c.y - shadow_intensity * c.y, c.z - shadow_intensity * c.z, c.w - shadow_intensity * c.w);
        if (cs != 0xFFFFFFFFu) {
            const float4 t = nextT[cs];
            nextT[cs] = make_float4(t.x - shadow_intensity * t.x, t.y - shadow_intensity * t.y, t.z - shadow_intensity * t.z,
                                    t.w - shadow_intensity * t.w);
        }
    }
    accumulate(acc, __float_as_uint(b.w), c);
}

// trn_occluded: plain (o, d, tmax) arrays -> a shadow wave whose "pixel" is the ray index and whose contribution is 1
__global__ void pack_shadow_queries_kernel(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ tmax,
                                           uint32_t count, ShadowWave sw, WaveCounters* __restrict__ counters) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) counters->shadow_count = count;
    if (idx >= count) return;
    sw.a[idx] = make_float4(o[3 * idx], o[3 * idx + 1], o[3 * idx + 2], d[3 * idx]);
    sw.b[idx] = make_float4(d[3 * idx + 1], d[3 * idx + 2], tmax[idx], __uint_as_float(idx));
    sw.c[idx] = make_float4(1.f, 0.f, 0.f, 0.f);
}
__global__ void unpack_occlusion_kernel(const float4* __restrict__ acc, uint32_t count, uint8_t* __restrict__ occluded) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < count) occluded[idx] = acc[idx].x == 0.f ? 1 : 0; // the shadow kernel adds the contribution iff unoccluded
}

// Gather micro-benchmark (trn_measure_gather_peak): every lane issues independent 16-byte loads at pseudo-random
// element indices of buf[0..elems) -- the access shape of the traversal kernels (node pairs, plane records, id vectors:
// one 16-byte gather per lane, a different line per lane). What the set size decides is only where the lines live.
__global__ void __launch_bounds__(256) gather_peak_kernel(const uint4* __restrict__ buf, uint32_t elems, int iters, int mode,
                                                          uint4* __restrict__ sink) {
    (void)mode;
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            s = s * 1664525u + 1013904223u;
            const uint4 v = __ldg(buf + __umulhi(s, elems));
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    }
    if (acc.x == 0x12345678u) *sink = acc; // never true for the memset pattern: keeps the loads alive
}

__global__ void unpack_hits_kernel(const uint4* __restrict__ hits, uint32_t count, uint32_t* __restrict__ ids,
                                   float* __restrict__ rst) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    const uint4 h = hits[idx];
    ids[idx] = h.x;
    const bool miss = h.x == kMiss;
    rst[3 * idx] = miss ? 0.f : __uint_as_float(h.y);
    rst[3 * idx + 1] = miss ? 0.f : __uint_as_float(h.z);
    rst[3 * idx + 2] = miss ? 0.f : __uint_as_float(h.w);
}

} // namespace trn
