// Pooled kd traversal -- persistent warps whose lanes WALK their own rays but whose triangle tests are pooled over the
// warp. Production kernel for closest-hit and shadow waves of every tree with >= 1024 leaves (A/B numbers, the SIMT
// model and the CPU warp emulation behind the design: profiles/README.md, tools/simt_sim.cpp, tools/pooled_emul.cpp).
//
// Why: the while-while kernel (traverse_persistent.cuh) issues 930 warp-instructions per secondary ray with 8.5 of 32
// lanes active: the lanes of a warp need different numbers of inner steps to reach their next leaf (mean 3, max ~10),
// their leaves hold different numbers of triangles (1..40), and the few plane hits that need the full ray/triangle
// test (6 % of the tests) are evaluated by 2-3 lanes while the rest wait. Here every warp runs a three-phase cycle:
//
//   WALK   up to walk_iters warp-wide iterations; in each, a lane at an inner node takes ONE step of its own ray. The
//          lanes that have reached a leaf queue a 16-byte leaf descriptor in shared memory (slots by ballot/popc; the queue
//          fill is a warp-uniform register, no atomics) and pop their stack -- in a block that is only issued once
//          leaf_gate lanes wait (or nobody can step). A leaf that does not fit the queue (or one descriptor) keeps its
//          remaining references in the node register and is continued in the next cycle. The walk is speculative: it does
//          not wait for the leaf's outcome (98.6 % of the leaf visits of the benchmark's secondary rays do not end the ray).
//   TEST   the queued leaves of ALL rays, 32 at a time (one descriptor per lane), are cut into chunks of 4 triangle
//          references; the chunks are dealt out 32 per round, one per lane, by a shuffle binary search over the running
//          chunk totals. The lane reads the owner ray from a table in shared memory, the chunk's ids with one 16-byte
//          load (leaf runs are 16-byte aligned), four 16-byte plane records, and runs a division-free, conservative plane
//          pre-filter (FMA arithmetic with explicit error bounds). Survivors are appended to the warp's survivor queue
//          with four ballots -- no atomics.
//   EXACT  whenever 32 survivors are waiting (and at the end of the cycle) they get, 32 at a time, the reference's exact
//          operation sequence (lib/intersection.h:40-89: plane distance with IEEE division, barycentric part); accepted
//          hits are handed to the owner lane, ordered by a visiting-order sequence number.
//
// Bit-exact contract (tests/test_gpu_parity.py): a triangle is accepted, and its (r, s, t) computed, ONLY by the exact
// sequence in EXACT; the pre-filter can only discard triangles whose exact plane distance lies outside the current
// cell's parameter range (a hit there is found again in the cell that contains it) -- it never decides a hit.
// Visiting order, first-visited-wins on exact ties and the early exit are those of traverse_pairs<> in kernels.cuh;
// rays with an exact-zero direction component take traverse_pairs<> itself (reference schedule verbatim).
#pragma once
#include "kernels.cuh"

namespace trn {

#ifndef TRN_PQ_LEAVES
#define TRN_PQ_LEAVES 64 // leaf descriptors per warp queue
#endif
#ifndef TRN_PQ_MINBLOCKS
#define TRN_PQ_MINBLOCKS 8
#endif
constexpr int kPqLeaves = TRN_PQ_LEAVES;
constexpr int kPqChunkTris = 4;                  // triangle references a lane tests per TEST round
constexpr int kPqLeafMaxRefs = 4096;             // references one descriptor carries (a multiple of kPqChunkTris; the rest
                                                 // of a longer leaf is queued in the next cycle)
constexpr int kPqSurv = 32 + 32 * kPqChunkTris; // survivors: < 32 left over + one TEST round
static_assert(kPqLeafMaxRefs % kPqChunkTris == 0 && kPqLeafMaxRefs < (1 << 20) && kPqLeaves <= 128,
              "seq = (leaf slot << 20) + in-leaf offset + 1 must fit the 27 bits next to the owner lane");

struct PooledWarpSmem {
    float4 ray_o[32];        // o.xyz, E   (E, F: error bounds of the pre-filter)
    float4 ray_d[32];        // d.xyz, F
    uint4 leaf[kPqLeaves];   // first ref, count | owner << 24, lo bits, hi bits (parameter range of the cell)
    uint2 surv[kPqSurv];     // triangle id, owner | seq << 5
    uint4 best[32];          // per lane: best hit so far (id, r, s, t) -- cold state kept out of the registers
    uint32_t ray_idx[32];    // per lane: index of its ray in the wave
    float walk_o[3][32];     // per lane: origin and reciprocal direction by axis -- the WALK step reads the split axis'
    float walk_i[3][32];     // component with one conflict-free LDS instead of holding six registers + selects
};

#ifndef TRN_PQ_STEPS
#define TRN_PQ_STEPS 2 // inner-node steps per walk iteration (1: 1602, 2: 1669, 3: 1648 Mrays/s on the 1M mesh, profiles/README.md)
#endif
#ifndef TRN_PQ_NODE_HINT
#define TRN_PQ_NODE_HINT 0 // 1: node-pair loads carry L1::evict_last (A/B in profiles/README.md)
#endif
#ifndef TRN_PQ_TRI_HINT
#define TRN_PQ_TRI_HINT 0 // 1: id vectors and plane records are loaded L1::evict_first
#endif
#ifndef TRN_PQ_REFPLANES
#define TRN_PQ_REFPLANES 1 // plane records per leaf REFERENCE: a chunk's four planes are one 64-byte block behind the leaf's first
#endif                     // reference (two 32-byte loads, no id -> plane gather); 0 = per triangle, through the ids (A/B, profiles/README.md)
#ifndef TRN_PQ_TREELET
#define TRN_PQ_TREELET 0 // node pairs of the top treelet staged in shared memory per CTA (0 = off; A/B in profiles/README.md)
#endif

// visit counts of the instrumented instantiation (COUNT): what the production schedule really requests from memory
struct PooledCounts {
    unsigned steps, chunks, tris, exact, cold, push, pop, leaves, cut_steps;
};
constexpr int kPooledCounters = 10; // slots per kernel mode in the visit array (9 used)

__device__ __forceinline__ uint4 ld_node_pair(const uint4* p) {
#if TRN_PQ_NODE_HINT == 1
    uint4 v;
    asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
template <typename T> __device__ __forceinline__ T ld_tri(const T* p) {
#if TRN_PQ_TRI_HINT == 1
    return __ldcs(p);
#else
    return __ldg(p);
#endif
}

// exact-zero direction component: the reference's schedule verbatim (see traverse_pairs<>)
template <bool ANY>
__device__ __noinline__ bool trace_axis_parallel(const DevScene& sc, float ox, float oy, float oz, float dx, float dy, float dz,
                                                 float tmax_any, HitRec& h) {
    return traverse_pairs<ANY>(sc, ox, oy, oz, dx, dy, dz, tmax_any, h);
}

// MODE 0: closest hit, rays from a RayWave (a,b); result -> hits[idx]
// MODE 1: any-hit shadow rays from a ShadowWave (a,b,c); unoccluded -> acc[pixel] += c
// MODE 2: closest hit, rays from plain (o,d) float arrays; result -> hits[idx]
// COUNT: also tally what the schedule requests (walk steps, chunks, triangle pre-tests, exact tests, cold-record reads,
//        stack pushes / pops, leaves, steps at empty-space cuts) into visits[0..9) -- the roofline leg of bench.py and the counter parity test only
template <int MODE, bool COUNT = false>
__global__ void __launch_bounds__(128, TRN_PQ_MINBLOCKS) trace_pooled_kernel(
    DevScene sc, const float4* __restrict__ planes, const float4* __restrict__ ra, const float4* __restrict__ rb,
    const float4* __restrict__ rc, const float* __restrict__ po, const float* __restrict__ pd, uint32_t count_arg,
    const uint32_t* __restrict__ count_ptr, uint32_t* __restrict__ cursor, uint4* __restrict__ hits, float4* __restrict__ acc,
    int refill_below, int walk_iters, uint32_t pool_chunk, int leaf_gate, unsigned long long* __restrict__ visits) {
    constexpr bool ANY = MODE == 1;
    constexpr unsigned kFull = 0xffffffffu;
    __shared__ PooledWarpSmem smem[4];
    PooledWarpSmem& sm = smem[threadIdx.x >> 5];
#if TRN_PQ_TREELET > 0
    // top treelet staged in shared memory: the first TRN_PQ_TREELET pairs of the breadth-first-on-top layout
    __shared__ uint4 s_pairs[TRN_PQ_TREELET];
    for (uint32_t k = threadIdx.x; k < TRN_PQ_TREELET; k += blockDim.x)
        s_pairs[k] = k < sc.treelet_pairs ? __ldg(reinterpret_cast<const uint4*>(sc.pnodes) + k) : make_uint4(0u, 3u, 0u, 3u);
    __syncthreads();
#endif
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t count = count_ptr ? *count_ptr : count_arg;
    float scale = 0.f; // largest |coordinate| of the scene box
#pragma unroll
    for (int c = 0; c < 3; ++c) scale = fmaxf(scale, fmaxf(fabsf(sc.lo[c]), fabsf(sc.hi[c])));
    PooledCounts pc{0, 0, 0, 0, 0, 0, 0, 0, 0};

    uint4 stack[kStackDepth];
    // Walk state of the lane's ray without extra flags (they cost register moves in the hot loop): sp >= 0 = walking with
    // sp stack entries, sp == -1 = the walk is over (or the lane is idle); bit 31 of a LEAF's n.y = "blocked": the leaf did
    // not fit the queue in this cycle and continues from n in the next one (a leaf's count has 24 bits, an inner node's
    // pair index 29: the bit is free).
    constexpr uint32_t kBlocked = 0x80000000u;
    int sp = -1;
    float tenter = 0, texit = 0, tmax_any = 0, last_texit = -kFltMax;
    uint2 n = make_uint2(0u, 3u);
    uint32_t best_seq = 0;
    float best_r = kFltMax; // mirrors sm.best[lane].y (kFltMax while there is no hit)
    bool busy = false, occluded = false, exhausted = false;
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
                    const uint32_t idx = pool_next + rank;
                    float ox, oy, oz, dx, dy, dz;
                    if (MODE == 2) {
                        ox = po[3 * idx]; oy = po[3 * idx + 1]; oz = po[3 * idx + 2];
                        dx = pd[3 * idx]; dy = pd[3 * idx + 1]; dz = pd[3 * idx + 2];
                    } else {
                        const float4 a = __ldcs(&ra[idx]);
                        const float4 b = __ldcs(&rb[idx]);
                        ox = a.x; oy = a.y; oz = a.z; dx = a.w; dy = b.x; dz = b.y;
                        if (ANY) tmax_any = b.z;
                    }
                    const bool axis_parallel = dx == 0.f || dy == 0.f || dz == 0.f || sc.verbatim != 0u;
                    bool done = false, hit = false;
                    HitRec h;
                    h.id = kMiss; h.r = kFltMax; h.s = 0.f; h.t = 0.f;
                    if (axis_parallel) {
                        hit = trace_axis_parallel<ANY>(sc, ox, oy, oz, dx, dy, dz, tmax_any, h);
                        done = true;
                    } else {
                        // fix_direction (lib/kdtree.cpp:503-511) changes nothing: no component is zero
                        const float ix = 1 / dx, iy = 1 / dy, iz = 1 / dz;
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
                            n = __ldg(&sc.pnodes[0]);
                            best_r = kFltMax;
                            best_seq = 0;
                            sm.best[lane] = make_uint4(kMiss, __float_as_uint(kFltMax), 0u, 0u);
                            sm.ray_idx[lane] = idx;
                            last_texit = -kFltMax;
                            occluded = false;
                            busy = true;
                            sp = (ANY && cell_lo(tenter) > tmax_any) ? -1 : 0;
                            // Error bounds of the pre-filter (u = 2^-24): its plane numerator b = dp - n.o (dp = n.v0 rounded
                            // once, three FMAs) and the reference's n.(v0 - o) (lib/intersection.h:47) both lie within
                            // 11 u (3 S + |o|_1) of each other, S = largest |coordinate| of the scene; the denominators
                            // n.d within 6 u |d|_1. E and F carry a safety factor of ~3.
                            const float E = 1.9073486e-6f * (3.f * scale + (fabsf(ox) + fabsf(oy) + fabsf(oz)));
                            const float F = 9.5367432e-7f * (fabsf(dx) + fabsf(dy) + fabsf(dz));
                            sm.ray_o[lane] = make_float4(ox, oy, oz, E);
                            sm.ray_d[lane] = make_float4(dx, dy, dz, F);
                            sm.walk_o[0][lane] = ox; sm.walk_o[1][lane] = oy; sm.walk_o[2][lane] = oz;
                            sm.walk_i[0][lane] = ix; sm.walk_i[1][lane] = iy; sm.walk_i[2][lane] = iz;
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
        // Every iteration the lanes at inner nodes take one step. The lanes that have reached a leaf queue its chunks and
        // pop -- but that block is only issued when at least leaf_gate lanes wait at a leaf (or no lane can step): it is
        // as long as the step itself and would otherwise run for 4 of 32 lanes in nearly every iteration.
        uint32_t nleaf = 0; // leaf descriptors queued in this cycle (warp-uniform register: every lane runs the leaf block)
        int it = 0;
#pragma unroll 1
        for (;;) {
            const bool at_leaf = sp >= 0 && (n.y & (kBlocked | 3u)) == 3u;
            const bool at_inner = sp >= 0 && (n.y & 3u) != 3u;
            const unsigned lm = __ballot_sync(kFull, at_leaf);
            const unsigned im = __ballot_sync(kFull, at_inner);
            const bool last = it >= walk_iters || im == 0u;
            if (lm != 0u && (last || __popc(lm) >= leaf_gate)) {
                // leaf: queue it (one descriptor: reference range, owner, parameter range of the cell plus slack) and pop
                // at once. A count-0 leaf (cut-off void) can only be the root of an empty tree.
                const uint32_t cnt = at_leaf ? n.y >> 2 : 0u;
                const unsigned want = __ballot_sync(kFull, cnt != 0u);
                const uint32_t slot = nleaf + __popc(want & lt_mask);
                nleaf = min(nleaf + static_cast<uint32_t>(__popc(want)), static_cast<uint32_t>(kPqLeaves));
                if (at_leaf) {
                    // a descriptor carries at most kPqLeafMaxRefs references (its count and the in-leaf offset of the
                    // visiting-order number are packed fields); a longer leaf -- or any leaf when the queue is full --
                    // keeps its remaining references in the node register (leaf runs are contiguous) and is continued
                    // in the next cycle
                    uint32_t take = 0;
                    if (cnt != 0u && slot < static_cast<uint32_t>(kPqLeaves)) {
                        take = min(cnt, static_cast<uint32_t>(kPqLeafMaxRefs));
                        float lo = cell_lo(tenter);
                        float hi = texit + kCellSlack * (fabsf(texit) + 1.f);
                        lo = fmaxf(lo, 0.f);
                        hi = ANY ? fminf(hi, tmax_any) : fminf(hi, best_r);
                        sm.leaf[slot] = make_uint4(n.x, take | (lane << 24), __float_as_uint(lo), __float_as_uint(hi));
                        if (COUNT) pc.leaves += 1;
                    }
                    if (take < cnt) {
                        n.x += take;
                        n.y = (n.y - (take << 2)) | kBlocked;
                    } else {
                        last_texit = texit;
                        if (sp == 0) {
                            sp = -1;
                        } else {
                            const uint4 e = stack[--sp];
                            if (COUNT) pc.pop += 1;
                            n = make_uint2(e.x, e.y);
                            tenter = __uint_as_float(e.z);
                            texit = __uint_as_float(e.w);
                            // front to back: nothing at or behind a cell that starts beyond the light / the best hit matters
                            if (ANY ? cell_lo(tenter) > tmax_any : tenter > best_r) sp = -1;
                        }
                    }
                }
            }
            if (last) break;
            ++it;
            // one inner-node step, lib/kdtree.cpp:540-563 on the sibling-pair layout (see traverse_pairs<>); TRN_PQ_STEPS > 1
            // repeats it for the lanes that are still at an inner node, halving the loop control per step
            bool stepping = at_inner;
#pragma unroll
            for (int rep = 0; rep < TRN_PQ_STEPS; ++rep) {
                if (stepping) {
                    const uint32_t ax = n.y & 3u;
                    const float split = __uint_as_float(n.x);
#if TRN_PQ_TREELET > 0
                    const uint32_t pi = n.y >> 3; // pair index
                    const uint4 pair = pi < TRN_PQ_TREELET ? s_pairs[pi] : ld_node_pair(reinterpret_cast<const uint4*>(sc.pnodes + (n.y >> 2)));
#else
                    const uint4 pair = ld_node_pair(reinterpret_cast<const uint4*>(sc.pnodes + (n.y >> 2)));
#endif
                    if (COUNT) {
                        pc.steps += 1;
                        if ((pair.y == 3u) != (pair.w == 3u)) pc.cut_steps += 1; // the node is an empty-space cut
                    }
                    const float o_ax = sm.walk_o[ax][lane], i_ax = sm.walk_i[ax][lane];
                    const float t = (split - o_ax) * i_ax;
                    const bool flip = (__float_as_uint(i_ax) >> 31) != 0u;
                    const uint2 near = flip ? make_uint2(pair.z, pair.w) : make_uint2(pair.x, pair.y);
                    const uint2 far = flip ? make_uint2(pair.x, pair.y) : make_uint2(pair.z, pair.w);
                    const bool near_only = texit < t;
                    const bool far_only = !near_only && (t < tenter);
                    const bool both = !near_only && !far_only;
                    const bool go_far = far_only || (both && near.y == 3u);
                    if (both && near.y != 3u && far.y != 3u) {
                        stack[sp++] = make_uint4(far.x, far.y, __float_as_uint(t), __float_as_uint(texit));
                        if (COUNT) pc.push += 1;
                    }
                    n = go_far ? far : near;
                    tenter = (both && go_far) ? t : tenter;
                    texit = (both && !go_far) ? t : texit;
                    if (TRN_PQ_STEPS > 1) stepping = (n.y & 3u) != 3u;
                }
            }
        }
        __syncwarp();

        // ------------------------------------------------------------------ TEST (pre-filter) and EXACT rounds
        // The queued leaves are taken 32 at a time, one per lane; their triangle references are cut into chunks of 4
        // and the chunks of the whole batch are dealt out 32 per round (a lane finds its chunk's leaf by a shuffle
        // binary search over the running chunk totals), so that every lane tests 4 triangles per round whatever the
        // leaf sizes are. Survivors go to the warp's survivor queue; whenever 32 are waiting (and at the end) they
        // get the exact test, 32 at a time.
        uint32_t ns = 0;        // survivors waiting (warp-uniform; only this loop appends)
        uint32_t lb = 0;        // first leaf of the current batch
        uint32_t base = 0, total = 0, P = 0; // chunk cursor / chunk count of the batch / inclusive chunk totals per lane
        uint4 ld = make_uint4(0u, 0u, 0u, 0u);
        bool have_batch = false;
        for (;;) {
            const bool more_tests = have_batch ? (base < total || lb + 32u < nleaf) : (lb < nleaf);
            if (ns >= 32u || (!more_tests && ns > 0u)) {
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
                    const float4 ro = sm.ray_o[owner], rd = sm.ray_d[owner];
                    const float4* rec = sc.isect_hot + 2 * static_cast<size_t>(id);
                    const float4* rec2 = sc.isect_cold + 2 * static_cast<size_t>(id);
                    const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1);
                    if (COUNT) pc.exact += 1;
                    const float nx = q0.w, ny = q1.x, nz = q1.y;
                    const float denom = nx * rd.x + ny * rd.y + nz * rd.z; // intersect_ray_plane, lib/intersection.h:40-49
                    const float nom = nx * (q0.x - ro.x) + ny * (q0.y - ro.y) + nz * (q0.z - ro.z);
                    r = nom / denom;
                    // r < 0 rejects (intersection.h:66); only a hit not farther than the owner's best (ties are settled
                    // below), resp. within the light distance, matters
                    if (denom != 0.f && r >= 0.f && r <= lim) {
                        const float4 q2 = __ldg(rec2), q3 = __ldg(rec2 + 1);
                        if (COUNT) pc.cold += 1;
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
                            best_r = r_;
                            best_seq = q_;
                            sm.best[lane] = make_uint4(id_, __float_as_uint(r_), __float_as_uint(s_), __float_as_uint(t_));
                        }
                    }
                }
                ns = sbase;
                __syncwarp();
            } else if (more_tests) {
                if (!have_batch || base >= total) {
                    // next batch of up to 32 leaves: one descriptor per lane, inclusive scan of their chunk counts
                    if (have_batch) lb += 32u;
                    have_batch = true;
                    ld = lb + lane < nleaf ? sm.leaf[lb + lane] : make_uint4(0u, 0u, 0u, 0u);
                    P = ((ld.y & 0xffffffu) + kPqChunkTris - 1) / kPqChunkTris;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const uint32_t v = __shfl_up_sync(kFull, P, off);
                        if (static_cast<int>(lane) >= off) P += v;
                    }
                    total = __shfl_sync(kFull, P, 31);
                    base = 0;
                    continue;
                }
                // TEST: one chunk per lane. Pre-filter: with a ~ n.d and b ~ n.(v0 - o) (FMA arithmetic, |a - denom| <= F,
                // |b - nom| <= E for the reference's denom, nom), A = |a|, B = b * sign(a):
                //   0 <= lo <= nom/denom <= hi   ==>   A <= F  or  (B + E >= lo (A - F)  and  B - E <= hi (A + F)).
                // Triangles that fail cannot have their exact plane distance inside [lo, hi].
                const uint32_t g = base + lane;
                // leaf of chunk g = number of lanes whose inclusive total is <= g
                uint32_t j = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const uint32_t pj = __shfl_sync(kFull, P, (j + step - 1) & 31u);
                    if (pj <= g) j += step;
                }
                j &= 31u; // g >= total (idle lane): any leaf, masked below
                const uint32_t first = __shfl_sync(kFull, ld.x, j), cw = __shfl_sync(kFull, ld.y, j);
                const float lo = __uint_as_float(__shfl_sync(kFull, ld.z, j)), hi = __uint_as_float(__shfl_sync(kFull, ld.w, j));
                const uint32_t pend = __shfl_sync(kFull, P, j);
                const uint32_t lcnt = cw & 0xffffffu, owner = cw >> 24;
                const uint32_t sub = g - (pend - (lcnt + kPqChunkTris - 1) / kPqChunkTris); // chunk index inside the leaf
                const uint32_t off0 = sub * kPqChunkTris;
                const uint32_t cnt = g < total ? min(static_cast<uint32_t>(kPqChunkTris), lcnt - off0) : 0u;
                uint4 ids = make_uint4(0u, 0u, 0u, 0u); // the chunk's triangle ids
                uint32_t km = 0;                         // bit k: triangle k survives the pre-filter
                if (cnt > 0u) {
                    const float4 ro = sm.ray_o[owner], rd = sm.ray_d[owner];
                    const float E = ro.w, F = rd.w;
                    const float c1 = fmaf(-lo, F, -E), c2 = fmaf(hi, F, E);
                    // leaf runs start at multiples of 4 references and the array is padded (kdtree_build.cpp): one 16-byte
                    // load brings the chunk's ids; ids beyond cnt are valid triangles whose result is masked
#if TRN_PQ_REFPLANES
                    // one plane record per reference, in the order of the reference array (leaf runs start at multiples of 4 and
                    // the array is padded): the chunk's planes are one aligned 64-byte block, fetched as two LDG.E.256
                    struct alignas(32) F8 { float4 a, b; };
                    const float4* pr = planes + first + off0;
                    const F8 v0 = *reinterpret_cast<const F8*>(pr), v1 = *reinterpret_cast<const F8*>(pr + 2);
                    const float4 p0 = v0.a, p1 = v0.b, p2 = v1.a, p3 = v1.b;
#else
                    ids = ld_tri(reinterpret_cast<const uint4*>(sc.prefs + first + off0));
                    const float4 p0 = ld_tri(&planes[ids.x]), p1 = ld_tri(&planes[ids.y]), p2 = ld_tri(&planes[ids.z]), p3 = ld_tri(&planes[ids.w]);
#endif
                    if (COUNT) {
                        pc.chunks += 1;
                        pc.tris += cnt;
                    }
#pragma unroll
                    for (int k = 0; k < kPqChunkTris; ++k) {
                        const float4 p = k == 0 ? p0 : (k == 1 ? p1 : (k == 2 ? p2 : p3));
                        const float a = fmaf(p.x, rd.x, fmaf(p.y, rd.y, p.z * rd.z));
                        const float b = fmaf(-p.x, ro.x, fmaf(-p.y, ro.y, fmaf(-p.z, ro.z, p.w)));
                        const float A = fabsf(a);
                        const float B = __uint_as_float(__float_as_uint(b) ^ (__float_as_uint(a) & 0x80000000u));
                        // bitwise, not short-circuit: no branches in the round
                        const bool keep = (static_cast<uint32_t>(k) < cnt) & ((A <= F) | ((B >= fmaf(lo, A, c1)) & (B <= fmaf(hi, A, c2))));
                        km |= keep ? (1u << k) : 0u;
                    }
#if TRN_PQ_REFPLANES
                    if (km != 0u) ids = ld_tri(reinterpret_cast<const uint4*>(sc.prefs + first + off0)); // ids only for survivors
#endif
                }
                // append the survivors of the round, k-major (the order inside the queue is irrelevant: seq carries the
                // visiting order): four ballots, no scan, no atomics
                const uint32_t seq0 = ((lb + j) << 20) + off0 + 1u;
                uint32_t round_total = 0;
#pragma unroll
                for (int k = 0; k < kPqChunkTris; ++k) {
                    const unsigned bk = __ballot_sync(kFull, (km >> k) & 1u);
                    if ((km >> k) & 1u) {
                        const uint32_t id = k == 0 ? ids.x : (k == 1 ? ids.y : (k == 2 ? ids.z : ids.w));
                        sm.surv[ns + round_total + __popc(bk & lt_mask)] = make_uint2(id, owner | ((seq0 + k) << 5));
                    }
                    round_total += __popc(bk);
                }
                ns += round_total;
                base += 32u;
                __syncwarp();
            } else {
                break;
            }
        }

        // ------------------------------------------------------------------ finished rays
        if (busy) {
            best_seq = 0;
            bool finished;
            const bool blocked = sp >= 0 && (n.y & (kBlocked | 3u)) == (kBlocked | 3u);
            if (blocked) n.y &= ~kBlocked;
            if (ANY) finished = occluded || sp < 0;
            // (a lane whose leaf is only partly queued waits for the rest of the leaf)
            else finished = sp < 0 || (!blocked && best_r < kFltMax && (best_r <= last_texit || best_r < tenter));
            if (finished) {
                busy = false;
                sp = -1;
                const uint32_t idx = sm.ray_idx[lane];
                if (ANY) {
                    if (!occluded) accumulate(acc, __float_as_uint(__ldcs(&rb[idx]).w), __ldcs(&rc[idx]));
                } else {
                    __stcs(&hits[idx], sm.best[lane]);
                }
            }
        }
        __syncwarp();
    }
    if (COUNT) {
        unsigned v[9] = {pc.steps, pc.chunks, pc.tris, pc.exact, pc.cold, pc.push, pc.pop, pc.leaves, pc.cut_steps};
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            unsigned x = v[k];
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(kFull, x, off);
            if (lane == 0 && x) atomicAdd(visits + k, static_cast<unsigned long long>(x));
        }
    }
}

} // namespace trn
