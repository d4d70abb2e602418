// Persistent-warp kd traversal with lane refill.
//
// ncu on the one-thread-per-ray kernel (profiles/r1_trace_closest_v1_mesh1m.txt) showed the issue slots
// half busy but only 5-11 of 32 lanes active per instruction: rays of one warp finish at very different
// times (misses leave at once, grazing rays visit dozens of leaves). Here a warp is a long-lived
// worker: each lane owns one ray (its traversal stack lives in local memory); whenever fewer than
// kRefillBelow lanes are still busy the idle lanes pull fresh rays from a per-warp pool that is
// replenished with one atomicAdd per kPoolChunk rays. Grid = resident CTAs per SM x 148 SMs.
//
// The per-ray arithmetic and visiting order are exactly those of traverse<> in kernels.cuh (same
// bit-exact contract); only the scheduling of rays onto lanes differs.
#pragma once
#include "kernels.cuh"

namespace trn {

constexpr int kRefillBelow = 22;       // refill when fewer lanes than this are busy
constexpr uint32_t kPoolChunk = 256;   // rays a warp reserves per global atomic

struct LaneState {
    float ox, oy, oz, dx, dy, dz;
    float ix, iy, iz;
    float tenter, texit;
    float tmax_any;
    uint32_t node;
    int sp;
    uint32_t idx;
    HitRec best;
};

// set up one ray; returns false when it misses the scene box (intersect_ray_box, lib/intersection.h:105-128)
__device__ __forceinline__ bool lane_begin(const DevScene& sc, LaneState& L) {
    const float fdx = L.dx == 0.f ? kEpsDir : L.dx;
    const float fdy = L.dy == 0.f ? kEpsDir : L.dy;
    const float fdz = L.dz == 0.f ? kEpsDir : L.dz;
    L.ix = 1 / fdx;
    L.iy = 1 / fdy;
    L.iz = 1 / fdz;
    float tx1 = (sc.lo[0] - L.ox) * L.ix, tx2 = (sc.hi[0] - L.ox) * L.ix;
    float tenter = fminf(tx1, tx2), texit = fmaxf(tx1, tx2);
    float ty1 = (sc.lo[1] - L.oy) * L.iy, ty2 = (sc.hi[1] - L.oy) * L.iy;
    tenter = fmaxf(tenter, fminf(ty1, ty2));
    texit = fminf(texit, fmaxf(ty1, ty2));
    float tz1 = (sc.lo[2] - L.oz) * L.iz, tz2 = (sc.hi[2] - L.oz) * L.iz;
    tenter = fmaxf(tenter, fminf(tz1, tz2));
    texit = fminf(texit, fmaxf(tz1, tz2));
    L.best.id = kMiss;
    L.best.r = kFltMax;
    L.best.s = 0.f;
    L.best.t = 0.f;
    L.node = 0;
    L.sp = 0;
    if (texit < tenter) return false;
    L.tenter = tenter < 0.f ? 0.f : tenter;
    L.texit = texit;
    return true;
}

// One scheduling quantum of one lane: walk down to a leaf, test its triangles, then either finish or pop the
// next cell. Returns true when the ray is finished (result in L.best).
template <bool ANY_HIT>
__device__ __forceinline__ bool lane_step(const DevScene& sc, LaneState& L, uint32_t* stk_node, float* stk_tmin, float* stk_tmax) {
    uint32_t node = L.node;
    float tenter = L.tenter, texit = L.texit;
    uint2 n = __ldg(&sc.nodes[node]);
    while ((n.y & 3u) != 3u) {
        const int ax = static_cast<int>(n.y & 3u);
        const float split = __uint_as_float(n.x);
        const float o_ax = sel3(ax, L.ox, L.oy, L.oz);
        const float i_ax = sel3(ax, L.ix, L.iy, L.iz);
        const float t = (split - o_ax) * i_ax;
        uint32_t near = node + 1, far = n.y >> 2;
        // fixed_ray.d[ax] <= 0 (lib/kdtree.cpp:551): the fixed direction is never 0, so its sign is the sign of 1/d
        if (__float_as_uint(i_ax) >> 31) {
            const uint32_t tmp = near;
            near = far;
            far = tmp;
        }
        if (texit < t) {
            node = near;
        } else if (t < tenter) {
            node = far;
        } else {
            stk_node[L.sp] = far;
            stk_tmin[L.sp] = t;
            stk_tmax[L.sp] = texit;
            ++L.sp;
            node = near;
            texit = t;
        }
        n = __ldg(&sc.nodes[node]);
    }

    const uint32_t first = n.x, count = n.y >> 2;
    for (uint32_t i = 0; i < count; ++i) {
        const uint32_t id = __ldg(&sc.leaf_refs[first + i]);
        const float4* rec = sc.isect + 4 * static_cast<size_t>(id);
        const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1);
        const float nx = q0.w, ny = q1.x, nz = q1.y;
        const float denom = nx * L.dx + ny * L.dy + nz * L.dz; // lib/intersection.h:40-49
        if (denom == 0.f) continue;
        const float nom = nx * (q0.x - L.ox) + ny * (q0.y - L.oy) + nz * (q0.z - L.oz);
        const float r = nom / denom;
        if (!(r >= 0.f)) continue;
        if (ANY_HIT ? !(r <= L.tmax_any) : !(r < L.best.r)) continue;
        const float4 q2 = __ldg(rec + 2), q3 = __ldg(rec + 3);
        const float wx = (L.ox + r * L.dx) - q0.x, wy = (L.oy + r * L.dy) - q0.y, wz = (L.oz + r * L.dz) - q0.z;
        const float ux = q1.z, uy = q1.w, uz = q2.x, vx = q2.y, vy = q2.z, vz = q2.w;
        const float wv = wx * vx + wy * vy + wz * vz;
        const float wu = wx * ux + wy * uy + wz * uz;
        const float s = (q3.x * wv - q3.y * wu) / q3.w; // lib/intersection.h:78-86
        if (s < 0.f) continue;
        const float t = (q3.x * wu - q3.z * wv) / q3.w;
        if (t < 0.f || 1.f < s + t) continue;
        L.best.id = id;
        L.best.r = r;
        L.best.s = s;
        L.best.t = t;
        if (ANY_HIT) return true;
    }

    if (L.best.id != kMiss && L.best.r <= texit) return true;
    if (L.sp == 0) return true;
    --L.sp;
    L.node = stk_node[L.sp];
    L.tenter = stk_tmin[L.sp];
    L.texit = stk_tmax[L.sp];
    if (ANY_HIT && L.tenter > L.tmax_any) return true;
    return false;
}

// MODE 0: closest hit, rays from a RayWave (a,b); result -> hits[idx]
// MODE 1: any-hit shadow rays from a ShadowWave; unoccluded -> acc[pixel] += c
// MODE 2: closest hit, rays from plain (o,d) float arrays
template <int MODE>
__global__ void __launch_bounds__(128) trace_persistent_kernel(DevScene sc, const float4* __restrict__ ra,
                                                               const float4* __restrict__ rb,
                                                               const float4* __restrict__ rc, const float* __restrict__ po,
                                                               const float* __restrict__ pd, uint32_t count_arg,
                                                               const uint32_t* __restrict__ count_ptr,
                                                               uint32_t* __restrict__ cursor, uint4* __restrict__ hits,
                                                               float4* __restrict__ acc) {
    constexpr bool ANY = MODE == 1;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t count = count_ptr ? *count_ptr : count_arg;

    uint32_t stk_node[kStackDepth];
    float stk_tmin[kStackDepth];
    float stk_tmax[kStackDepth];

    LaneState L;
    L.tmax_any = 0.f;
    bool busy = false;
    bool exhausted = false;           // warp-uniform
    uint32_t pool_next = 0, pool_end = 0; // warp-uniform

    for (;;) {
        // ---- refill idle lanes
        unsigned need = __ballot_sync(0xffffffffu, !busy);
        while (need != 0u && !exhausted) {
            if (pool_next == pool_end) {
                uint32_t b = 0;
                if (lane == 0) b = atomicAdd(cursor, kPoolChunk);
                b = __shfl_sync(0xffffffffu, b, 0);
                if (b >= count) {
                    exhausted = true;
                    break;
                }
                pool_next = b;
                pool_end = min(b + kPoolChunk, count);
            }
            const uint32_t avail = pool_end - pool_next;
            const uint32_t want = __popc(need);
            const uint32_t take = min(want, avail);
            const uint32_t rank = __popc(need & lt_mask);
            if (!busy && rank < take) {
                const uint32_t idx = pool_next + rank;
                L.idx = idx;
                if (MODE == 2) {
                    L.ox = po[3 * idx]; L.oy = po[3 * idx + 1]; L.oz = po[3 * idx + 2];
                    L.dx = pd[3 * idx]; L.dy = pd[3 * idx + 1]; L.dz = pd[3 * idx + 2];
                } else {
                    const float4 a = ra[idx];
                    const float4 b = rb[idx];
                    L.ox = a.x; L.oy = a.y; L.oz = a.z; L.dx = a.w; L.dy = b.x; L.dz = b.y;
                    if (ANY) L.tmax_any = b.z;
                }
                if (lane_begin(sc, L)) {
                    busy = true;
                } else if (ANY) {
                    accumulate(acc, __float_as_uint(rb[idx].w), rc[idx]); // nothing in the way: lit
                } else {
                    hits[idx] = make_uint4(kMiss, __float_as_uint(kFltMax), 0u, 0u);
                }
            }
            pool_next += take;
            need = __ballot_sync(0xffffffffu, !busy);
        }
        if (!__any_sync(0xffffffffu, busy)) break;

        // ---- traverse until too few lanes are busy
        int nbusy;
        do {
            if (busy) {
                if (lane_step<ANY>(sc, L, stk_node, stk_tmin, stk_tmax)) {
                    busy = false;
                    if (ANY) {
                        if (L.best.id == kMiss) accumulate(acc, __float_as_uint(rb[L.idx].w), rc[L.idx]);
                    } else {
                        hits[L.idx] = make_uint4(L.best.id, __float_as_uint(L.best.r), __float_as_uint(L.best.s), __float_as_uint(L.best.t));
                    }
                }
            }
            nbusy = __popc(__ballot_sync(0xffffffffu, busy));
        } while (nbusy >= kRefillBelow || (exhausted && nbusy > 0));
    }
}

} // namespace trn
