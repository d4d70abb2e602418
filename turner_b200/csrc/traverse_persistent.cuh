// Persistent-warp kd traversal with lane refill -- the closest-hit kernel of trees that are a few big leaves (shadow
// waves of those trees keep the one-thread-per-ray kernel of kernels.cuh). Trees with >= 1024 leaves run the pooled
// kernel of traverse_pooled.cuh, which grew out of this one; A/B numbers in profiles/README.md.
//
// Why: ncu on the one-thread-per-ray kernel (profiles/r1_prof_trace_v2.txt) shows the issue slots 60-70 % busy but only
// 5-12 of 32 lanes active per instruction. Ray cost is heavy-tailed (a miss leaves at once, a grazing ray visits
// hundreds of leaves), so a warp of 32 fixed rays idles most of its lanes waiting for the slowest.
//
// Here a warp is a long-lived worker and a lane is a slot holding one ray (its traversal stack lives in local memory):
//   * one scheduling quantum of a lane = walk down to the next leaf, test its triangles, pop ("while-while");
//   * after every two quanta the warp counts its busy lanes; when fewer than 28 are busy the idle lanes take fresh
//     rays from a per-warp pool that is replenished with one atomicAdd per 32 rays on the wave's work cursor
//     (bigger chunks leave a static, unbalanced tail; smaller ones make the cursor a hot spot);
//   * grid = resident CTAs per SM x number of SMs (9 x 148 on B200): no tail of half-empty CTAs.
//
// Per-ray arithmetic, visiting order and tie rule are exactly those of traverse_pairs<> in kernels.cuh (the bit-exact
// contract); only the scheduling of rays onto lanes differs.
#pragma once
#include "kernels.cuh"

namespace trn {

// MODE 0: closest hit, rays from a RayWave (a,b); result -> hits[idx]
// MODE 1: any-hit shadow rays from a ShadowWave (a,b,c); unoccluded -> acc[pixel] += c
// MODE 2: closest hit, rays from plain (o,d) float arrays; result -> hits[idx]

// Barycentric part of intersect_ray_triangle (lib/intersection.h:70-86) for a triangle whose plane distance r already
// passed 0 <= r; re-checks r < best (strict: the first-visited triangle keeps a tie). Returns true only in ANY mode
// when the triangle is hit (occluder found).
template <bool ANY>
__device__ __forceinline__ bool test_candidate(const DevScene& sc, uint32_t id, float r, float ox, float oy, float oz, float dx,
                                               float dy, float dz, uint32_t& best_id, float& best_r, float& best_s, float& best_t) {
    if (!ANY && !(r < best_r)) return false;
    const float4* rec = sc.isect_hot + 2 * static_cast<size_t>(id);
    const float4* rec2 = sc.isect_cold + 2 * static_cast<size_t>(id);
    const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1), q2 = __ldg(rec2), q3 = __ldg(rec2 + 1);
    const float wx = (ox + r * dx) - q0.x, wy = (oy + r * dy) - q0.y, wz = (oz + r * dz) - q0.z;
    const float ux = q1.z, uy = q1.w, uz = q2.x, vx = q2.y, vy = q2.z, vz = q2.w;
    const float wv = wx * vx + wy * vy + wz * vz;
    const float wu = wx * ux + wy * uy + wz * uz;
    const float s = (q3.x * wv - q3.y * wu) / q3.w;
    if (s < 0.f) return false;
    const float t = (q3.x * wu - q3.z * wv) / q3.w;
    if (t < 0.f || 1.f < s + t) return false;
    best_id = id;
    best_r = r;
    best_s = s;
    best_t = t;
    return ANY;
}


// Variant 2 ("while-while" quantum): same persistent warps + lane refill, but one scheduling quantum of a lane is
// "walk down to the next leaf, test its triangles, pop" -- the loop body of traverse_pairs<> -- instead of a
// single state-machine step. Fewer bookkeeping instructions per unit of work, coarser balancing.
// TWO_PASS: leaves are evaluated as "all plane tests, then the recorded survivors" (best for trees with small leaves);
// otherwise in one pass (best when a leaf holds many triangles, e.g. a scene that is a single leaf). Chosen per scene.
#ifndef TRN_TREELET_NODES
#define TRN_TREELET_NODES 0 // nodes of the top treelet staged in shared memory by the persistent kernels (0 = off)
#endif
#ifndef TRN_WW_MINBLOCKS
#define TRN_WW_MINBLOCKS 9
#endif
template <int MODE, bool TWO_PASS>
__global__ void __launch_bounds__(128, TRN_WW_MINBLOCKS) trace_persistent_ww_kernel(DevScene sc, const float4* __restrict__ ra,
                                                                  const float4* __restrict__ rb,
                                                                  const float4* __restrict__ rc, const float* __restrict__ po,
                                                                  const float* __restrict__ pd, uint32_t count_arg,
                                                                  const uint32_t* __restrict__ count_ptr,
                                                                  uint32_t* __restrict__ cursor, uint4* __restrict__ hits,
                                                                  float4* __restrict__ acc, int refill_below, int quanta,
                                                                  const uint32_t* __restrict__ order, uint32_t treelet_pairs,
                                                                  uint32_t pool_chunk) {
    constexpr bool ANY = MODE == 1;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t count = count_ptr ? *count_ptr : count_arg;
#if TRN_TREELET_NODES > 0
    // top treelet staged in shared memory: the first TRN_TREELET_NODES nodes of the breadth-first-on-top layout
    __shared__ uint4 s_pairs[TRN_TREELET_NODES / 2];
    for (uint32_t k = threadIdx.x; k < TRN_TREELET_NODES / 2; k += blockDim.x)
        s_pairs[k] = k < treelet_pairs ? __ldg(reinterpret_cast<const uint4*>(sc.pnodes) + k) : make_uint4(0u, 3u, 0u, 3u);
    __syncthreads();
#endif

    uint4 stack[kStackDepth];
    int sp = 0;
    float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0, ix = 0, iy = 0, iz = 0;
    float tenter = 0, texit = 0, tmax_any = 0;
    uint2 n = make_uint2(0u, 3u);
    uint32_t best_id = kMiss, idx = 0;
    float best_r = kFltMax, best_s = 0.f, best_t = 0.f;
    bool busy = false, axis_parallel = false;
    bool exhausted = false;
    uint32_t pool_next = 0, pool_end = 0;

    for (;;) {
        int nbusy = __popc(__ballot_sync(0xffffffffu, busy));
        if (nbusy < refill_below && !exhausted) {
            if (pool_next == pool_end) {
                uint32_t b = 0;
                if (lane == 0) b = atomicAdd(cursor, pool_chunk);
                b = __shfl_sync(0xffffffffu, b, 0);
                if (b >= count) {
                    exhausted = true;
                } else {
                    pool_next = b;
                    pool_end = min(b + pool_chunk, count);
                }
            }
            if (!exhausted) {
                const unsigned need = __ballot_sync(0xffffffffu, !busy);
                const uint32_t take = min(static_cast<uint32_t>(__popc(need)), pool_end - pool_next);
                const uint32_t rank = __popc(need & lt_mask);
                if (!busy && rank < take) {
                    idx = pool_next + rank;
                    if (order) idx = __ldcs(&order[idx]); // rays are consumed in sorted order, results go to their own slot
                    if (MODE == 2) {
                        ox = po[3 * idx]; oy = po[3 * idx + 1]; oz = po[3 * idx + 2];
                        dx = pd[3 * idx]; dy = pd[3 * idx + 1]; dz = pd[3 * idx + 2];
                    } else {
                        const float4 a = __ldcs(&ra[idx]);
                        const float4 b = __ldcs(&rb[idx]);
                        ox = a.x; oy = a.y; oz = a.z; dx = a.w; dy = b.x; dz = b.y;
                        if (ANY) tmax_any = b.z;
                    }
                    const float fdx = dx == 0.f ? kEpsDir : dx;
                    const float fdy = dy == 0.f ? kEpsDir : dy;
                    const float fdz = dz == 0.f ? kEpsDir : dz;
                    ix = 1 / fdx;
                    iy = 1 / fdy;
                    iz = 1 / fdz;
                    float tx1 = (sc.lo[0] - ox) * ix, tx2 = (sc.hi[0] - ox) * ix;
                    float t0 = fminf(tx1, tx2), t1 = fmaxf(tx1, tx2);
                    float ty1 = (sc.lo[1] - oy) * iy, ty2 = (sc.hi[1] - oy) * iy;
                    t0 = fmaxf(t0, fminf(ty1, ty2));
                    t1 = fminf(t1, fmaxf(ty1, ty2));
                    float tz1 = (sc.lo[2] - oz) * iz, tz2 = (sc.hi[2] - oz) * iz;
                    t0 = fmaxf(t0, fminf(tz1, tz2));
                    t1 = fminf(t1, fmaxf(tz1, tz2));
                    best_id = kMiss;
                    best_r = kFltMax;
                    best_s = 0.f;
                    best_t = 0.f;
                    if (t1 < t0) {
                        if (ANY) accumulate(acc, __float_as_uint(__ldcs(&rb[idx]).w), __ldcs(&rc[idx]));
                        else __stcs(&hits[idx], make_uint4(kMiss, __float_as_uint(kFltMax), 0u, 0u));
                    } else {
                        axis_parallel = dx == 0.f || dy == 0.f || dz == 0.f || sc.verbatim != 0u; // reference schedule verbatim, see traverse_pairs<>
                        tenter = (t0 < 0.f && !axis_parallel) ? 0.f : t0;
                        texit = t1;
                        sp = 0;
                        n = __ldg(&sc.pnodes[0]);
                        busy = true;
                    }
                }
                pool_next += take;
            }
            nbusy = __popc(__ballot_sync(0xffffffffu, busy));
        }
        if (nbusy == 0) {
            if (exhausted) break;
            continue;
        }

#pragma unroll 1
        for (int q = 0; q < quanta; ++q) {
            if (!busy) continue;
            while ((n.y & 3u) != 3u) {
                const int ax = static_cast<int>(n.y & 3u);
                const float split = __uint_as_float(n.x);
#if TRN_TREELET_NODES > 0
                const uint32_t child = n.y >> 2;
                const uint4 pair = child < TRN_TREELET_NODES ? s_pairs[child >> 1]
                                                             : __ldg(reinterpret_cast<const uint4*>(sc.pnodes + child));
#else
                const uint4 pair = __ldg(reinterpret_cast<const uint4*>(sc.pnodes + (n.y >> 2)));
#endif
                const float o_ax = sel3(ax, ox, oy, oz);
                const float i_ax = sel3(ax, ix, iy, iz);
                const float t = (split - o_ax) * i_ax;
                if (axis_parallel && (pair.y == 3u || pair.w == 3u)) { // cut node: the reference has no plane here
                    n = pair.y == 3u ? make_uint2(pair.z, pair.w) : make_uint2(pair.x, pair.y);
                    continue;
                }
                const bool flip = (__float_as_uint(i_ax) >> 31) != 0u;
                const uint2 near = flip ? make_uint2(pair.z, pair.w) : make_uint2(pair.x, pair.y);
                const uint2 far = flip ? make_uint2(pair.x, pair.y) : make_uint2(pair.z, pair.w);
                // lib/kdtree.cpp:555-563 without divergent branches: near only | far only | both (far pushed), with the
                // cut-off voids (count-0 leaves, y == 3) never entered nor pushed
                const bool near_only = texit < t;
                const bool far_only = !near_only && (t < tenter);
                const bool both = !near_only && !far_only;
                const bool go_far = far_only || (both && near.y == 3u);
                if (both && near.y != 3u && far.y != 3u) stack[sp++] = make_uint4(far.x, far.y, __float_as_uint(t), __float_as_uint(texit));
                n = go_far ? far : near;
                tenter = (both && go_far) ? t : tenter;
                texit = (both && !go_far) ? t : texit;
            }
            const uint32_t first = n.x, cnt = n.y >> 2;
            bool occluded = false;
            const float r_lo = axis_parallel ? -kFltMax : tenter - kCellSlack * (fabsf(tenter) + 1.f);
            const float r_hi = axis_parallel ? kFltMax : texit + kCellSlack * (fabsf(texit) + 1.f);
            // Two passes over the leaf so that the lanes of the warp stay together: first the plane test of every
            // triangle (lib/intersection.h:40-49), which only RECORDS the few survivors (0 <= r < best, r inside the
            // cell); then the barycentric part (intersection.h:70-86) for the recorded ones, in visiting order and
            // re-checking r < best, i.e. with exactly the accept/replace decisions of the one-pass loop.
            uint32_t c0_id = 0, c1_id = 0;
            float c0_r = 0.f, c1_r = 0.f;
            int ncand = 0;
            const uint32_t cnt_two_pass = TWO_PASS ? cnt : 0u;
            // two triangles per trip: both ids, then both records are requested before either is used, so a leaf costs
            // one id latency + one record latency per PAIR of triangles instead of per triangle
            for (uint32_t i = 0; i < cnt_two_pass; i += 2) {
                const bool two = i + 1 < cnt_two_pass;
                const uint32_t ida = __ldg(&sc.prefs[first + i]);
                const uint32_t idb = two ? __ldg(&sc.prefs[first + i + 1]) : ida;
                const float4* reca = sc.isect_hot + 2 * static_cast<size_t>(ida);
                const float4* recb = sc.isect_hot + 2 * static_cast<size_t>(idb);
                const float4 a0 = __ldg(reca), a1 = __ldg(reca + 1), b0 = __ldg(recb), b1 = __ldg(recb + 1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float4 q0 = h ? b0 : a0, q1 = h ? b1 : a1;
                    const uint32_t id = h ? idb : ida;
                    const float nx = q0.w, ny = q1.x, nz = q1.y;
                    const float denom = nx * dx + ny * dy + nz * dz;
                    const float nom = nx * (q0.x - ox) + ny * (q0.y - oy) + nz * (q0.z - oz);
                    const float r = nom / denom;
                    const bool cand = (h == 0 || two) && denom != 0.f && r >= 0.f && (ANY ? (r <= tmax_any) : (r < best_r)) &&
                                      r >= r_lo && r <= r_hi;
                    if (cand) {
                        if (ncand == 2) { // rare: drain the two recorded survivors first (keeps visiting order)
                            if (test_candidate<ANY>(sc, c0_id, c0_r, ox, oy, oz, dx, dy, dz, best_id, best_r, best_s, best_t)) occluded = true;
                            if (!occluded && test_candidate<ANY>(sc, c1_id, c1_r, ox, oy, oz, dx, dy, dz, best_id, best_r, best_s, best_t)) occluded = true;
                            ncand = 0;
                        }
                        if (ncand == 0) {
                            c0_id = id;
                            c0_r = r;
                        } else {
                            c1_id = id;
                            c1_r = r;
                        }
                        ++ncand;
                    }
                }
            }
            if (ncand > 0 && !occluded && test_candidate<ANY>(sc, c0_id, c0_r, ox, oy, oz, dx, dy, dz, best_id, best_r, best_s, best_t)) occluded = true;
            if (ncand > 1 && !occluded && test_candidate<ANY>(sc, c1_id, c1_r, ox, oy, oz, dx, dy, dz, best_id, best_r, best_s, best_t)) occluded = true;
            for (uint32_t i = cnt_two_pass; !TWO_PASS && i < cnt && !occluded; ++i) {
                const uint32_t id = __ldg(&sc.prefs[first + i]);
                const float4* rec = sc.isect_hot + 2 * static_cast<size_t>(id);
                const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1);
                const float nx = q0.w, ny = q1.x, nz = q1.y;
                const float denom = nx * dx + ny * dy + nz * dz;
                if (denom == 0.f) continue;
                const float nom = nx * (q0.x - ox) + ny * (q0.y - oy) + nz * (q0.z - oz);
                const float r = nom / denom;
                if (!(r >= 0.f)) continue;
                if (ANY ? !(r <= tmax_any) : !(r < best_r)) continue;
                if (!(r >= r_lo && r <= r_hi)) continue;
                // In a big leaf most surviving plane hits lie far outside their triangle (the infinite planes of the two
                // cornell boxes are crossed by almost every ray): reject those on the triangle's (grown) bounding box,
                // 15 instructions instead of the 50 of the barycentric part with its two divisions. A point outside the
                // grown box is outside the triangle, so the reference rejects it as well.
                const float4 blo = __ldg(sc.tri_box + 2 * static_cast<size_t>(id)), bhi = __ldg(sc.tri_box + 2 * static_cast<size_t>(id) + 1);
                const float hx = ox + r * dx, hy = oy + r * dy, hz = oz + r * dz;
                if (hx < blo.x || hy < blo.y || hz < blo.z || hx > bhi.x || hy > bhi.y || hz > bhi.z) continue;
                if (test_candidate<ANY>(sc, id, r, ox, oy, oz, dx, dy, dz, best_id, best_r, best_s, best_t)) occluded = true;
            }
            bool finished = occluded || (!axis_parallel && best_id != kMiss && best_r <= texit) || sp == 0;
            if (!finished) {
                const uint4 e = stack[--sp];
                n = make_uint2(e.x, e.y);
                tenter = __uint_as_float(e.z);
                texit = __uint_as_float(e.w);
                if (ANY && !axis_parallel && cell_lo(tenter) > tmax_any) finished = true;
            }
            if (finished) {
                busy = false;
                if (ANY) {
                    if (!occluded) accumulate(acc, __float_as_uint(__ldcs(&rb[idx]).w), __ldcs(&rc[idx]));
                } else {
                    __stcs(&hits[idx], make_uint4(best_id, __float_as_uint(best_r), __float_as_uint(best_s), __float_as_uint(best_t)));
                }
            }
        }
    }
}


// Sort key of a secondary ray: direction octant (3 bits, major) and the Morton code of its origin on a 32^3 grid over
// the scene box (15 bits). Rays that are neighbours in this order start close together and make the same near/far
// decisions, so the lanes of a warp walk the same nodes (profiles/README.md: lanes per instruction, L1 hit rate).
__device__ __forceinline__ uint32_t spread5(uint32_t v) { // 5 bits -> every third bit
    v = (v | (v << 8)) & 0x0000F00Fu;
    v = (v | (v << 4)) & 0x000C30C3u;
    v = (v | (v << 2)) & 0x00249249u;
    return v;
}
__global__ void __launch_bounds__(256) ray_sort_keys_kernel(DevScene sc, const float4* __restrict__ ra,
                                                            const float4* __restrict__ rb, uint32_t count,
                                                            uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    const float4 a = ra[idx];
    const float4 b = rb[idx];
    uint32_t q[3];
    const float o[3] = {a.x, a.y, a.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float ext = sc.hi[c] - sc.lo[c];
        float u = ext > 0.f ? (o[c] - sc.lo[c]) / ext : 0.f;
        u = fminf(fmaxf(u, 0.f), 0.999999f);
        q[c] = static_cast<uint32_t>(u * 32.f);
    }
    const uint32_t morton = spread5(q[0]) | (spread5(q[1]) << 1) | (spread5(q[2]) << 2);
    const uint32_t octant = (a.w < 0.f ? 1u : 0u) | (b.x < 0.f ? 2u : 0u) | (b.y < 0.f ? 4u : 0u);
    keys[idx] = (octant << 15) | morton;
    vals[idx] = idx;
}

} // namespace trn
