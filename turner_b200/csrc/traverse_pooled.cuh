// Pooled kd traversal -- persistent warps whose lanes WALK their own rays but whose triangle tests are pooled over the
// warp. Production kernel for closest-hit and shadow waves (A/B numbers and the SIMT model behind the design:
// profiles/README.md, tools/simt_sim.cpp).
//
// Why: the while-while kernel (traverse_persistent.cuh) issues 930 warp-instructions per secondary ray with 8.5 of 32
// lanes active: the lanes of a warp need different numbers of inner steps to reach their next leaf (mean 3, max ~10),
// their leaves hold different numbers of triangles (1..20), and the few plane hits that need the full ray/triangle
// test (6 % of the tests) are evaluated by 2-3 lanes while the rest wait. Here every warp runs a three-phase cycle:
//
//   WALK   a fixed number of warp-wide iterations; in each, a lane takes ONE inner-node step of its own ray or, at a
//          leaf, appends the leaf to the warp's queue in shared memory as chunks of <= 4 triangle references and pops
//          its stack at once. The walk is speculative: it does not wait for the leaf's outcome (98.6 % of the leaf
//          visits of the benchmark's secondary rays do not end the ray). ~25 lanes active.
//   TEST   the queued chunks of ALL rays are dealt out 32 at a time, one chunk per lane; the lane reads the owner ray
//          from a table in shared memory and runs a division-free, conservative plane pre-filter on the chunk's
//          triangles (16-byte plane records, FMA arithmetic with explicit error bounds). ~28 lanes active.
//   EXACT  pre-filter survivors of all rays are pooled too and get, 32 at a time, the reference's exact operation
//          sequence (lib/intersection.h:40-89: plane distance with IEEE division, barycentric part); accepted hits are
//          handed to the owner lane in visiting order.
//
// Bit-exact contract (tests/test_gpu_parity.py): a triangle is accepted, and its (r, s, t) computed, ONLY by the exact
// sequence in EXACT; the pre-filter can only discard triangles whose exact plane distance lies outside the current
// cell's parameter range (a hit there is found again in the cell that contains it) -- it never decides a hit.
// Visiting order, first-visited-wins on exact ties and the early exit are those of traverse_pairs<> in kernels.cuh;
// rays with an exact-zero direction component take traverse_pairs<> itself (reference schedule verbatim).
#pragma once
#include "kernels.cuh"

namespace trn {

#ifndef TRN_PQ_CHUNKS
#define TRN_PQ_CHUNKS 128 // chunk descriptors per warp queue
#endif
#ifndef TRN_PQ_MINBLOCKS
#define TRN_PQ_MINBLOCKS 8
#endif
constexpr int kPqChunks = TRN_PQ_CHUNKS;
constexpr int kPqChunkTris = 4;                  // triangle references per chunk
constexpr int kPqMaxAppend = 16;                 // chunks a lane may queue per walk iteration (bigger leaves: instalments)
constexpr int kPqSurv = 32 + 32 * kPqChunkTris; // survivors: < 32 left over + one TEST round

struct PooledWarpSmem {
    float4 ray[64];          // [2*lane] = o.xyz, E   [2*lane+1] = d.xyz, F   (E, F: error bounds of the pre-filter)
    uint4 chunk[kPqChunks];  // first ref, count | owner << 8, lo bits, hi bits (parameter range of the cell)
    uint2 surv[kPqSurv];     // triangle id, owner | seq << 5
    uint32_t nchunk, nvalid, nsurv, pad;
};

// exact-zero direction component: the reference's schedule verbatim (see traverse_pairs<>)
template <bool ANY>
__device__ __noinline__ bool trace_axis_parallel(const DevScene& sc, float ox, float oy, float oz, float dx, float dy, float dz,
                                                 float tmax_any, HitRec& h) {
    return traverse_pairs<ANY>(sc, ox, oy, oz, dx, dy, dz, tmax_any, h);
}

// MODE 0: closest hit, rays from a RayWave (a,b); result -> hits[idx]
// MODE 1: any-hit shadow rays from a ShadowWave (a,b,c); unoccluded -> acc[pixel] += c
// MODE 2: closest hit, rays from plain (o,d) float arrays; result -> hits[idx]
template <int MODE>
__global__ void __launch_bounds__(128, TRN_PQ_MINBLOCKS) trace_pooled_kernel(
    DevScene sc, const float4* __restrict__ planes, const float4* __restrict__ ra, const float4* __restrict__ rb,
    const float4* __restrict__ rc, const float* __restrict__ po, const float* __restrict__ pd, uint32_t count_arg,
    const uint32_t* __restrict__ count_ptr, uint32_t* __restrict__ cursor, uint4* __restrict__ hits, float4* __restrict__ acc,
    int refill_below, int walk_iters, uint32_t pool_chunk) {
    constexpr bool ANY = MODE == 1;
    constexpr unsigned kFull = 0xffffffffu;
    __shared__ PooledWarpSmem smem[4];
    PooledWarpSmem& sm = smem[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t count = count_ptr ? *count_ptr : count_arg;
    if (lane == 0) {
        sm.nchunk = 0;
        sm.nvalid = kPqChunks;
        sm.nsurv = 0;
    }
    __syncwarp();
    float scale = 0.f; // largest |coordinate| of the scene box
#pragma unroll
    for (int c = 0; c < 3; ++c) scale = fmaxf(scale, fmaxf(fabsf(sc.lo[c]), fabsf(sc.hi[c])));

    uint4 stack[kStackDepth];
    int sp = 0;
    float ox = 0, oy = 0, oz = 0, ix = 0, iy = 0, iz = 0;
    float tenter = 0, texit = 0, tmax_any = 0, last_texit = -kFltMax;
    uint2 n = make_uint2(0u, 3u);
    uint32_t best_id = kMiss, best_seq = 0, idx = 0, leaf_off = 0;
    float best_r = kFltMax, best_s = 0.f, best_t = 0.f;
    bool busy = false, walking = false, occluded = false, exhausted = false;
    uint32_t pool_next = 0, pool_end = 0;

    for (;;) {
        // ------------------------------------------------------------------ refill idle lanes from the wave
        int nbusy = __popc(__ballot_sync(kFull, busy));
        if (nbusy < refill_below && !exhausted) {
            if (pool_next == pool_end) {
                uint32_t b = 0;
                if (lane == 0) b = atomicAdd(cursor, pool_chunk);
                b = __shfl_sync(kFull, b, 0);
                if (b >= count) {
                    exhausted = true;
                } else {
                    pool_next = b;
                    pool_end = min(b + pool_chunk, count);
                }
            }
            if (!exhausted) {
                const unsigned need = __ballot_sync(kFull, !busy);
                const uint32_t take = min(static_cast<uint32_t>(__popc(need)), pool_end - pool_next);
                const uint32_t rank = __popc(need & lt_mask);
                if (!busy && rank < take) {
                    idx = pool_next + rank;
                    float dx, dy, dz;
                    if (MODE == 2) {
                        ox = po[3 * idx]; oy = po[3 * idx + 1]; oz = po[3 * idx + 2];
                        dx = pd[3 * idx]; dy = pd[3 * idx + 1]; dz = pd[3 * idx + 2];
                    } else {
                        const float4 a = __ldcs(&ra[idx]);
                        const float4 b = __ldcs(&rb[idx]);
                        ox = a.x; oy = a.y; oz = a.z; dx = a.w; dy = b.x; dz = b.y;
                        if (ANY) tmax_any = b.z;
                    }
                    const bool axis_parallel = dx == 0.f || dy == 0.f || dz == 0.f;
                    bool done = false, hit = false;
                    HitRec h;
                    h.id = kMiss; h.r = kFltMax; h.s = 0.f; h.t = 0.f;
                    if (axis_parallel) {
                        hit = trace_axis_parallel<ANY>(sc, ox, oy, oz, dx, dy, dz, tmax_any, h);
                        done = true;
                    } else {
                        ix = 1 / dx; // fix_direction (lib/kdtree.cpp:503-511) changes nothing: no component is zero
                        iy = 1 / dy;
                        iz = 1 / dz;
                        // intersect_ray_box, lib/intersection.h:105-128
                        float tx1 = (sc.lo[0] - ox) * ix, tx2 = (sc.hi[0] - ox) * ix;
                        float t0 = fminf(tx1, tx2), t1 = fmaxf(tx1, tx2);
                        float ty1 = (sc.lo[1] - oy) * iy, ty2 = (sc.hi[1] - oy) * iy;
                        t0 = fmaxf(t0, fminf(ty1, ty2));
                        t1 = fminf(t1, fmaxf(ty1, ty2));
                        float tz1 = (sc.lo[2] - oz) * iz, tz2 = (sc.hi[2] - oz) * iz;
                        t0 = fmaxf(t0, fminf(tz1, tz2));
                        t1 = fminf(t1, fmaxf(tz1, tz2));
                        if (t1 < t0) {
                            done = true;
                        } else {
                            tenter = t0 < 0.f ? 0.f : t0;
                            texit = t1;
                            sp = 0;
                            n = __ldg(&sc.pnodes[0]);
                            best_id = kMiss;
                            best_r = kFltMax;
                            best_s = 0.f;
                            best_t = 0.f;
                            best_seq = 0;
                            last_texit = -kFltMax;
                            leaf_off = 0;
                            occluded = false;
                            busy = true;
                            walking = !(ANY && tenter > tmax_any);
                            // Error bounds of the pre-filter (u = 2^-24): its plane numerator b = dp - n.o (dp = n.v0 rounded
                            // once, three FMAs) and the reference's n.(v0 - o) (lib/intersection.h:47) both lie within
                            // 11 u (3 S + |o|_1) of each other, S = largest |coordinate| of the scene; the denominators
                            // n.d within 6 u |d|_1. E and F carry a safety factor of ~3.
                            const float E = 1.9073486e-6f * (3.f * scale + (fabsf(ox) + fabsf(oy) + fabsf(oz)));
                            const float F = 9.5367432e-7f * (fabsf(dx) + fabsf(dy) + fabsf(dz));
                            sm.ray[2 * lane] = make_float4(ox, oy, oz, E);
                            sm.ray[2 * lane + 1] = make_float4(dx, dy, dz, F);
                        }
                    }
                    if (done) {
                        if (ANY) {
                            if (!hit) accumulate(acc, __float_as_uint(__ldcs(&rb[idx]).w), __ldcs(&rc[idx]));
                        } else {
                            __stcs(&hits[idx], make_uint4(h.id, __float_as_uint(h.r), __float_as_uint(h.s), __float_as_uint(h.t)));
                        }
                    }
                }
                pool_next += take;
            }
            nbusy = __popc(__ballot_sync(kFull, busy));
        }
        if (nbusy == 0) {
            if (exhausted) break;
            continue;
        }
        __syncwarp();

        // ------------------------------------------------------------------ WALK
        bool blocked = false;
#pragma unroll 1
        for (int it = 0; it < walk_iters; ++it) {
            const bool can = busy && walking && !blocked;
            if (!__any_sync(kFull, can)) break;
            if (can) {
                if ((n.y & 3u) != 3u) {
                    // one inner-node step, lib/kdtree.cpp:540-563 on the sibling-pair layout (see traverse_pairs<>)
                    const uint32_t ax = n.y & 3u;
                    const float split = __uint_as_float(n.x);
                    const uint4 pair = __ldg(reinterpret_cast<const uint4*>(sc.pnodes + (n.y >> 2)));
                    float o_ax = oz, i_ax = iz;
                    if (ax == 0u) { o_ax = ox; i_ax = ix; }
                    if (ax == 1u) { o_ax = oy; i_ax = iy; }
                    const float t = (split - o_ax) * i_ax;
                    const bool flip = (__float_as_uint(i_ax) >> 31) != 0u;
                    const uint2 near = flip ? make_uint2(pair.z, pair.w) : make_uint2(pair.x, pair.y);
                    const uint2 far = flip ? make_uint2(pair.x, pair.y) : make_uint2(pair.z, pair.w);
                    const bool near_only = texit < t;
                    const bool far_only = !near_only && (t < tenter);
                    const bool both = !near_only && !far_only;
                    const bool go_far = far_only || (both && near.y == 3u);
                    if (both && near.y != 3u && far.y != 3u) stack[sp++] = make_uint4(far.x, far.y, __float_as_uint(t), __float_as_uint(texit));
                    n = go_far ? far : near;
                    tenter = (both && go_far) ? t : tenter;
                    texit = (both && !go_far) ? t : texit;
                } else {
                    // leaf: queue its triangle references as chunks; the parameter range of the cell (plus slack) travels
                    // with them. A count-0 leaf (cut-off void) can only be the root of an empty tree.
                    const uint32_t cnt = n.y >> 2;
                    const uint32_t rem = cnt - leaf_off;
                    const uint32_t ch = min((rem + kPqChunkTris - 1) / kPqChunkTris, static_cast<uint32_t>(kPqMaxAppend));
                    bool leaf_done = true;
                    if (ch > 0) {
                        const uint32_t slot = atomicAdd(&sm.nchunk, ch);
                        if (slot + ch <= kPqChunks) {
                            float lo = tenter - kCellSlack * (fabsf(tenter) + 1.f);
                            float hi = texit + kCellSlack * (fabsf(texit) + 1.f);
                            lo = fmaxf(lo, 0.f);
                            hi = ANY ? fminf(hi, tmax_any) : fminf(hi, best_r);
                            const uint32_t first = n.x + leaf_off;
                            for (uint32_t k = 0; k < ch; ++k) {
                                const uint32_t c = min(static_cast<uint32_t>(kPqChunkTris), rem - k * kPqChunkTris);
                                sm.chunk[slot + k] = make_uint4(first + k * kPqChunkTris, c | (lane << 8), __float_as_uint(lo), __float_as_uint(hi));
                            }
                            leaf_off += ch * kPqChunkTris;
                            leaf_done = leaf_off >= cnt;
                        } else {
                            atomicMin(&sm.nvalid, slot); // queue full: everything from this slot on is unwritten
                            blocked = true;
                            leaf_done = false;
                        }
                    }
                    if (leaf_done) {
                        leaf_off = 0;
                        last_texit = texit;
                        if (sp == 0) {
                            walking = false;
                        } else {
                            const uint4 e = stack[--sp];
                            n = make_uint2(e.x, e.y);
                            tenter = __uint_as_float(e.z);
                            texit = __uint_as_float(e.w);
                            // front to back: nothing at or behind a cell that starts beyond the light / the best hit matters
                            if (ANY ? tenter > tmax_any : tenter > best_r) walking = false;
                        }
                    }
                }
            }
        }
        __syncwarp();

        // ------------------------------------------------------------------ TEST (pre-filter) and EXACT rounds
        const uint32_t nch = min(*reinterpret_cast<volatile uint32_t*>(&sm.nchunk), *reinterpret_cast<volatile uint32_t*>(&sm.nvalid));
        uint32_t base = 0;
        for (;;) {
            const uint32_t ns = *reinterpret_cast<volatile uint32_t*>(&sm.nsurv);
            __syncwarp();
            if (ns >= 32u || (base >= nch && ns > 0u)) {
                // EXACT: the reference's operation sequence for up to 32 pooled survivors (taken from the tail)
                const uint32_t take = min(32u, ns), sbase = ns - take;
                bool pass = false;
                uint32_t id = 0, owner = 0, seq = 0;
                float r = 0.f, s = 0.f, t = 0.f;
                if (lane < take) {
                    const uint2 e = sm.surv[sbase + lane];
                    id = e.x;
                    owner = e.y & 31u;
                    seq = e.y >> 5;
                }
                const float lim = __shfl_sync(kFull, ANY ? tmax_any : best_r, owner);
                if (lane < take) {
                    const float4 ro = sm.ray[2 * owner], rd = sm.ray[2 * owner + 1];
                    const float4* rec = sc.isect_hot + 2 * static_cast<size_t>(id);
                    const float4* rec2 = sc.isect_cold + 2 * static_cast<size_t>(id);
                    const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1);
                    const float nx = q0.w, ny = q1.x, nz = q1.y;
                    const float denom = nx * rd.x + ny * rd.y + nz * rd.z; // intersect_ray_plane, lib/intersection.h:40-49
                    const float nom = nx * (q0.x - ro.x) + ny * (q0.y - ro.y) + nz * (q0.z - ro.z);
                    r = nom / denom;
                    // r < 0 rejects (intersection.h:66); only a hit nearer than the owner's best (strictly: ties are
                    // settled below), resp. within the light distance, matters
                    if (denom != 0.f && r >= 0.f && r <= lim) {
                        const float4 q2 = __ldg(rec2), q3 = __ldg(rec2 + 1);
                        const float wx = (ro.x + r * rd.x) - q0.x, wy = (ro.y + r * rd.y) - q0.y, wz = (ro.z + r * rd.z) - q0.z; // :70-71
                        const float ux = q1.z, uy = q1.w, uz = q2.x, vx = q2.y, vy = q2.z, vz = q2.w;
                        const float wv = wx * vx + wy * vy + wz * vz;
                        const float wu = wx * ux + wy * uy + wz * uz;
                        s = (q3.x * wv - q3.y * wu) / q3.w; // :78-86
                        if (!(s < 0.f)) {
                            t = (q3.x * wu - q3.z * wv) / q3.w;
                            pass = !(t < 0.f || 1.f < s + t);
                        }
                    }
                }
                // hand the accepted hits to their owners. Within one cycle the visiting order is the order of seq;
                // a hit from an earlier cycle (best_seq == 0) was visited before all of them.
                unsigned pm = __ballot_sync(kFull, pass);
                while (pm) {
                    const int src = __ffs(pm) - 1;
                    pm &= pm - 1;
                    const uint32_t o_ = __shfl_sync(kFull, owner, src);
                    const float r_ = __shfl_sync(kFull, r, src);
                    const float s_ = __shfl_sync(kFull, s, src);
                    const float t_ = __shfl_sync(kFull, t, src);
                    const uint32_t id_ = __shfl_sync(kFull, id, src);
                    const uint32_t q_ = __shfl_sync(kFull, seq, src);
                    if (lane == o_) {
                        if (ANY) {
                            occluded = true;
                        } else if (r_ < best_r || (r_ == best_r && q_ < best_seq)) {
                            best_id = id_;
                            best_r = r_;
                            best_s = s_;
                            best_t = t_;
                            best_seq = q_;
                        }
                    }
                }
                if (lane == 0) sm.nsurv = sbase;
                __syncwarp();
            } else if (base < nch) {
                // TEST: one chunk per lane. Pre-filter: with a ~ n.d and b ~ n.(v0 - o) (FMA arithmetic, |a - denom| <= F,
                // |b - nom| <= E for the reference's denom, nom), A = |a|, B = b * sign(a):
                //   0 <= lo <= nom/denom <= hi   ==>   A <= F  or  (B + E >= lo (A - F)  and  B - E <= hi (A + F)).
                // Triangles that fail cannot have their exact plane distance inside [lo, hi].
                const uint32_t g = base + lane;
                if (g < nch) {
                    const uint4 d = sm.chunk[g];
                    const uint32_t first = d.x, cnt = d.y & 0xffu, owner = d.y >> 8;
                    const float lo = __uint_as_float(d.z), hi = __uint_as_float(d.w);
                    const float4 ro = sm.ray[2 * owner], rd = sm.ray[2 * owner + 1];
                    const float E = ro.w, F = rd.w;
                    const float c1 = fmaf(-lo, F, -E), c2 = fmaf(hi, F, E);
                    for (uint32_t i = 0; i < cnt; i += 2) {
                        const bool two = i + 1 < cnt;
                        const uint32_t ida = __ldg(&sc.prefs[first + i]);
                        const uint32_t idb = two ? __ldg(&sc.prefs[first + i + 1]) : ida;
                        const float4 pa = __ldg(&planes[ida]), pb = __ldg(&planes[idb]);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const float4 p = h ? pb : pa;
                            const float a = fmaf(p.x, rd.x, fmaf(p.y, rd.y, p.z * rd.z));
                            const float b = fmaf(-p.x, ro.x, fmaf(-p.y, ro.y, fmaf(-p.z, ro.z, p.w)));
                            const float A = fabsf(a);
                            const float B = __uint_as_float(__float_as_uint(b) ^ (__float_as_uint(a) & 0x80000000u));
                            const bool keep = (h == 0 || two) && (A <= F || (B >= fmaf(lo, A, c1) && B <= fmaf(hi, A, c2)));
                            if (keep) {
                                const uint32_t sl = atomicAdd(&sm.nsurv, 1u);
                                sm.surv[sl] = make_uint2(h ? idb : ida, owner | ((g * 8u + i + h + 1u) << 5));
                            }
                        }
                    }
                }
                base += 32u;
                __syncwarp();
            } else {
                break;
            }
        }
        if (lane == 0) {
            sm.nchunk = 0;
            sm.nvalid = kPqChunks;
        }

        // ------------------------------------------------------------------ finished rays
        if (busy) {
            best_seq = 0;
            bool finished;
            if (ANY) finished = occluded || !walking;
            else finished = !walking || (best_id != kMiss && (best_r <= last_texit || best_r < tenter));
            if (finished) {
                busy = false;
                if (ANY) {
                    if (!occluded) accumulate(acc, __float_as_uint(__ldcs(&rb[idx]).w), __ldcs(&rc[idx]));
                } else {
                    __stcs(&hits[idx], make_uint4(best_id, __float_as_uint(best_r), __float_as_uint(best_s), __float_as_uint(best_t)));
                }
            }
        }
        __syncwarp();
    }
}

} // namespace trn
