// Brute-force queries for scenes whose kd-tree is ONE leaf (cornell_box: 36 triangles, "Kd-Tree Height 0", README.md:30-31;
// BASELINE configs 1, 2 and 4). The reference's traversal of such a tree is: slab test against the scene box with the
// "fixed" direction (lib/kdtree.cpp:503-522), then every triangle of the leaf in order (lib/kdtree.cpp:580-607). There is
// nothing to walk, so the kernel is organised around the triangle loop instead of around the tree:
//
//   SCAN   a warp takes 32 rays, one per lane, and runs ONE warp-uniform loop over the leaf's triangles. The triangle's
//          plane (16 bytes) and grown bounding box (2 x 16 bytes) come from shared memory as broadcasts; each lane runs the
//          division-free conservative plane pre-filter of the pooled kernel (FMA arithmetic, explicit error bounds E, F;
//          range 0 <= r <= best hit so far / light distance) and a bounding-box test of the APPROXIMATE hit point
//          (reciprocal instead of division; the box is grown by 1e-4 x scene scale, 100x the error of that point).
//          Survivors -- about two per ray -- go to a per-warp queue with one ballot per triangle.
//   EXACT  whenever 32 survivors wait (and at the end): the reference's exact operation sequence
//          (lib/intersection.h:40-89), 32 at a time with all lanes busy, hits handed to the owner lane; first-visited wins
//          an exact tie (the leaf's order = the scan's order).
//
// Bit-exact contract as in traverse_pooled.cuh: only EXACT accepts a triangle and computes (r, s, t); SCAN can only discard
// triangles whose exact plane distance is negative / beyond the limit or whose hit point lies outside the triangle's box.
#pragma once
#include "kernels.cuh"

namespace trn {

#ifndef TRN_FLAT_SCAN_BOX
#define TRN_FLAT_SCAN_BOX 0 // 1: SCAN also tests the approximate hit point against the triangle's box (fewer survivors, 2.5x the scan cost)
#endif
constexpr int kFlatMaxTris = 256;  // triangles of the single leaf this kernel accepts (48 + 4 bytes of shared memory each)
constexpr int kFlatSurv = 64;      // survivor queue: < 32 left over + one triangle's 32 lanes

struct FlatWarpSmem {
    float4 ray_o[32]; // o.xyz, E
    float4 ray_d[32]; // d.xyz, F
    uint2 surv[kFlatSurv]; // triangle slot (= visiting order), owner lane
};

// MODE 0: closest hit, rays from a RayWave (a,b) -> hits[idx];  MODE 1: any-hit shadow rays -> acc[pixel] += c when
// unoccluded;  MODE 2: closest hit, plain (o,d) arrays -> hits[idx]
template <int MODE>
__global__ void __launch_bounds__(128) trace_flat_kernel(DevScene sc, const float4* __restrict__ planes, uint32_t first_ref, uint32_t ntris,
                                                         const float4* __restrict__ ra, const float4* __restrict__ rb,
                                                         const float4* __restrict__ rc, const float* __restrict__ po,
                                                         const float* __restrict__ pd, uint32_t count_arg,
                                                         const uint32_t* __restrict__ count_ptr, uint32_t* __restrict__ cursor,
                                                         uint4* __restrict__ hits, float4* __restrict__ acc) {
    constexpr bool ANY = MODE == 1;
    constexpr unsigned kFull = 0xffffffffu;
    __shared__ float4 s_plane[kFlatMaxTris];
    __shared__ float4 s_blo[kFlatMaxTris], s_bhi[kFlatMaxTris];
    __shared__ uint32_t s_id[kFlatMaxTris];
    __shared__ FlatWarpSmem smem[4];
    for (uint32_t k = threadIdx.x; k < ntris; k += blockDim.x) {
        const uint32_t id = __ldg(&sc.prefs[first_ref + k]);
        s_id[k] = id;
        s_plane[k] = __ldg(&planes[id]);
        s_blo[k] = __ldg(sc.tri_box + 2 * static_cast<size_t>(id));
        s_bhi[k] = __ldg(sc.tri_box + 2 * static_cast<size_t>(id) + 1);
    }
    __syncthreads();
    FlatWarpSmem& sm = smem[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t count = count_ptr ? *count_ptr : count_arg;
    float scale = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) scale = fmaxf(scale, fmaxf(fabsf(sc.lo[c]), fabsf(sc.hi[c])));

    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(cursor, 32u);
        base = __shfl_sync(kFull, base, 0);
        if (base >= count) break;
        const uint32_t idx = base + lane;
        bool valid = idx < count;
        float ox = 0, oy = 0, oz = 0, dx = 1, dy = 1, dz = 1, tmax_any = 0;
        if (valid) {
            if (MODE == 2) {
                ox = po[3 * idx]; oy = po[3 * idx + 1]; oz = po[3 * idx + 2];
                dx = pd[3 * idx]; dy = pd[3 * idx + 1]; dz = pd[3 * idx + 2];
            } else {
                const float4 a = __ldcs(&ra[idx]);
                const float4 b = __ldcs(&rb[idx]);
                ox = a.x; oy = a.y; oz = a.z; dx = a.w; dy = b.x; dz = b.y;
                if (ANY) tmax_any = b.z;
            }
        }
        const bool in_wave = valid;
        {
            // intersect_ray_box with the fixed direction, lib/kdtree.cpp:503-522, lib/intersection.h:105-128: a ray that
            // misses the scene box tests nothing
            const float fdx = dx == 0.f ? kEpsDir : dx, fdy = dy == 0.f ? kEpsDir : dy, fdz = dz == 0.f ? kEpsDir : dz;
            const float ix = 1 / fdx, iy = 1 / fdy, iz = 1 / fdz;
            float tx1 = (sc.lo[0] - ox) * ix, tx2 = (sc.hi[0] - ox) * ix;
            float t0 = fminf(tx1, tx2), t1 = fmaxf(tx1, tx2);
            float ty1 = (sc.lo[1] - oy) * iy, ty2 = (sc.hi[1] - oy) * iy;
            t0 = fmaxf(t0, fminf(ty1, ty2));
            t1 = fminf(t1, fmaxf(ty1, ty2));
            float tz1 = (sc.lo[2] - oz) * iz, tz2 = (sc.hi[2] - oz) * iz;
            t0 = fmaxf(t0, fminf(tz1, tz2));
            t1 = fminf(t1, fmaxf(tz1, tz2));
            if (t1 < t0) valid = false;
        }
        // error bounds of the pre-filter, as in traverse_pooled.cuh
        const float E = 1.9073486e-6f * (3.f * scale + (fabsf(ox) + fabsf(oy) + fabsf(oz)));
        const float F = 9.5367432e-7f * (fabsf(dx) + fabsf(dy) + fabsf(dz));
        // the approximate hit point o + (B / A) d is off by at most (E + r F) / (A - F) * max|d_i| per coordinate; the box test
        // only counts where that stays below half of the boxes' growth (grazing rays keep the triangle instead)
        const float dmax = fmaxf(fabsf(dx), fmaxf(fabsf(dy), fabsf(dz)));
        const float box_tol = 0.5e-4f * scale;
        sm.ray_o[lane] = make_float4(ox, oy, oz, E);
        sm.ray_d[lane] = make_float4(dx, dy, dz, F);
        __syncwarp();
        uint32_t best_id = kMiss, best_seq = 0;
        float best_r = kFltMax, best_s = 0.f, best_t = 0.f;
        bool occluded = false;
        uint32_t ns = 0; // survivors waiting (warp-uniform)

        auto exact_round = [&]() {
            const uint32_t take = min(32u, ns), sbase = ns - take;
            bool pass = false;
            uint32_t slot = 0, owner = 0, id = 0;
            float r = 0.f, s = 0.f, t = 0.f;
            if (lane < take) {
                const uint2 e = sm.surv[sbase + lane];
                slot = e.x;
                owner = e.y;
                id = s_id[slot];
            }
            const float lim = __shfl_sync(kFull, ANY ? tmax_any : best_r, owner);
            if (lane < take) {
                const float4 ro = sm.ray_o[owner], rd = sm.ray_d[owner];
                const float4* rec = sc.isect_hot + 2 * static_cast<size_t>(id);
                const float4* rec2 = sc.isect_cold + 2 * static_cast<size_t>(id);
                const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1);
                const float nx = q0.w, ny = q1.x, nz = q1.y;
                const float denom = nx * rd.x + ny * rd.y + nz * rd.z; // intersect_ray_plane, lib/intersection.h:40-49
                const float nom = nx * (q0.x - ro.x) + ny * (q0.y - ro.y) + nz * (q0.z - ro.z);
                r = nom / denom;
                bool cand = denom != 0.f && r >= 0.f && r <= lim;
                if (cand) { // a hit point outside the triangle's (grown) box cannot pass the barycentric part (as traverse_pairs<>)
                    const float4 blo = s_blo[slot], bhi = s_bhi[slot];
                    const float hx = ro.x + r * rd.x, hy = ro.y + r * rd.y, hz = ro.z + r * rd.z;
                    cand = !(hx < blo.x || hy < blo.y || hz < blo.z || hx > bhi.x || hy > bhi.y || hz > bhi.z);
                }
                if (cand) {
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
            unsigned pm = __ballot_sync(kFull, pass);
            while (pm) {
                const int src = __ffs(pm) - 1;
                pm &= pm - 1;
                const uint32_t o_ = __shfl_sync(kFull, owner, src);
                const float r_ = __shfl_sync(kFull, r, src);
                const float s_ = __shfl_sync(kFull, s, src);
                const float t_ = __shfl_sync(kFull, t, src);
                const uint32_t id_ = __shfl_sync(kFull, id, src);
                const uint32_t q_ = __shfl_sync(kFull, slot, src);
                if (lane == o_) {
                    if (ANY) {
                        occluded = true;
                    } else if (r_ < best_r || (r_ == best_r && q_ < best_seq)) { // strict '<' keeps the first-visited triangle on ties
                        best_r = r_;
                        best_s = s_;
                        best_t = t_;
                        best_id = id_;
                        best_seq = q_;
                    }
                }
            }
            ns = sbase;
            __syncwarp();
        };

        if (__ballot_sync(kFull, valid) != 0u) {
#pragma unroll 1
            for (uint32_t k = 0; k < ntris; ++k) {
                const float4 p = s_plane[k];
                const float a = fmaf(p.x, dx, fmaf(p.y, dy, p.z * dz));
                const float b = fmaf(-p.x, ox, fmaf(-p.y, oy, fmaf(-p.z, oz, p.w)));
                const float A = fabsf(a);
                const float B = __uint_as_float(__float_as_uint(b) ^ (__float_as_uint(a) & 0x80000000u));
                // 0 <= nom/denom <= lim  ==>  A <= F  or  (B + E >= 0  and  B - E <= lim (A + F))   (traverse_pooled.cuh)
                const float lim = ANY ? tmax_any : best_r;
                const bool in_range = (B >= -E) & (B <= fmaf(lim, A + F, E));
#if TRN_FLAT_SCAN_BOX
                // approximate hit point against the triangle's grown box (kernels.cuh: tri_box is grown by 1e-4 x scene scale)
                const float4 blo = s_blo[k], bhi = s_bhi[k];
                const float rcp = __fdividef(1.f, fmaxf(A - F, 1e-30f));
                const float rr = B * rcp;
                const float hx = fmaf(rr, dx, ox), hy = fmaf(rr, dy, oy), hz = fmaf(rr, dz, oz);
                const bool trust = fmaf(fabsf(rr), F, E) * rcp * dmax <= box_tol;
                const bool inside = !trust | ((hx >= blo.x) & (hy >= blo.y) & (hz >= blo.z) & (hx <= bhi.x) & (hy <= bhi.y) & (hz <= bhi.z));
#else
                const bool inside = true;
                (void)dmax;
                (void)box_tol;
#endif
                const bool keep = valid & !(ANY && occluded) & ((A <= F) | (in_range & inside));
                const unsigned bk = __ballot_sync(kFull, keep);
                if (keep) sm.surv[ns + __popc(bk & lt_mask)] = make_uint2(k, lane);
                ns += __popc(bk);
                if (ns >= 32u) {
                    __syncwarp();
                    exact_round();
                }
            }
            __syncwarp();
            while (ns > 0u) exact_round();
        }
        if (in_wave) {
            if (ANY) {
                if (!occluded) accumulate(acc, __float_as_uint(__ldcs(&rb[idx]).w), __ldcs(&rc[idx]));
            } else {
                __stcs(&hits[idx], make_uint4(best_id, __float_as_uint(best_r), __float_as_uint(best_s), __float_as_uint(best_t)));
            }
        }
        __syncwarp();
    }
}

} // namespace trn
