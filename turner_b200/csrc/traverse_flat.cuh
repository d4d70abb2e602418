// Brute-force queries for scenes whose kd-tree is ONE leaf (cornell_box: 36 triangles, "Kd-Tree Height 0", README.md:30-31;
// BASELINE configs 1, 2 and 4). The reference's traversal of such a tree is: slab test against the scene box with the
// "fixed" direction (lib/kdtree.cpp:503-522), then every triangle of the leaf in order (lib/kdtree.cpp:580-607). There is
// nothing to walk, so the kernel is organised around the triangle loop instead of around the tree:
//
//   SCAN   one ray per lane; ONE warp-uniform loop over the leaf's triangles. The triangle's plane and grown bounding box
//          (40 bytes) are KERNEL PARAMETERS: they sit in the constant bank and reach the FMAs as uniform operands, so the
//          loop issues no load at all. Each lane runs the division-free conservative plane pre-filter of the pooled kernel
//          (FMA arithmetic, explicit error bounds E, F; range 0 <= r <= light distance) and tests the APPROXIMATE hit
//          point (one MUFU reciprocal instead of an IEEE division) against the triangle's box -- the box only counts where
//          the point's error bound stays below half of the box's growth. A survivor sets one bit of the lane's own 64-bit
//          candidate mask (two registers).
//   EXACT  each lane walks its own candidate bits in ascending (= visiting) order and runs the reference's exact operation
//          sequence (lib/intersection.h:40-89) on records staged in shared memory; strict '<' keeps the first-visited
//          triangle on an exact tie. About two candidates per ray are left, so the divergent part is short; nothing is
//          handed from lane to lane.
//
// Round-2 history (profiles/README.md): the first brute-force kernel pooled the survivors of a warp in a shared-memory queue
// and handed accepted hits to their owners by shuffles -- 123 warp-instructions per ray against 104 of the while-while
// kernel, so it lost. This one needs ~45.
//
// Bit-exact contract as in traverse_pooled.cuh: only EXACT accepts a triangle and computes (r, s, t); SCAN can only discard
// triangles whose exact plane distance is negative / beyond the limit or whose hit point lies outside the triangle's box.
#pragma once
#include "kernels.cuh"

namespace trn {

constexpr int kFlatMaxTris = 64; // triangles of the single leaf this kernel accepts: one candidate bit each in two registers
// growth of the scan's triangle boxes, relative to the scene scale. The approximate hit point of a ray that meets the plane at
// |cos| = A carries an error of ~1e-5 x scale / A, and the box only counts where that stays below half the growth: with the
// exact path's 1e-4 (DevScene::tri_box) every pair with A < 0.18 stayed a candidate (10 per ray); 5e-3 leaves A < 0.004.
constexpr float kFlatBoxGrow = 5e-3f;

// Pre-filter record of one SCAN GROUP: a triangle, or two coplanar triangles with (nearly) the same box -- the halves of a
// quad, which is every face of cornell_box. A pair is scanned once, with the leader's plane and the union box; the host only
// pairs triangles whose unit normals / plane offsets agree to 8 u / 8 u x scale, and E, F below carry that difference.
struct FlatTri {
    float nx, ny, nz, dp;    // plane: n, n.v0 (the pooled kernel's record)
    float blo[3], bhi[3];    // bounding box grown by kFlatBoxGrow x scene scale
};
struct FlatParams {
    FlatTri t[kFlatMaxTris];        // one per scan group; 2560 bytes of the 4 KB parameter space
    uint16_t members[kFlatMaxTris]; // the group's triangles as leaf slots + 1: k0 + 1 | (k1 + 1) << 8 (0 = no second triangle)
};

// MODE 0: closest hit, rays from a RayWave (a,b) -> hits[idx];  MODE 1: any-hit shadow rays -> acc[pixel] += c when
// unoccluded;  MODE 2: closest hit, plain (o,d) arrays -> hits[idx]
// NT: the number of scan groups rounded up to a multiple of 4 -- the SCAN loop is fully unrolled so that every constant is an
// immediate constant-bank operand (a runtime loop costs ~12 more instructions per record: indexed LDC/LDCU, moves, the
// shifted mask bit); the padding records' bits are masked off
template <int MODE, int NT>
__global__ void __launch_bounds__(128) trace_flat_kernel(DevScene sc, const __grid_constant__ FlatParams P, uint32_t first_ref, uint32_t ntris,
                                                         uint32_t ngroups,
                                                         const float4* __restrict__ ra, const float4* __restrict__ rb,
                                                         const float4* __restrict__ rc, const float* __restrict__ po,
                                                         const float* __restrict__ pd, uint32_t count_arg,
                                                         const uint32_t* __restrict__ count_ptr, uint4* __restrict__ hits,
                                                         float4* __restrict__ acc) {
    constexpr bool ANY = MODE == 1;
    // exact-test records, one array per 16-byte word so that lanes with different triangles spread over the banks
    __shared__ float4 s_rec[4][kFlatMaxTris];
    __shared__ uint32_t s_id[kFlatMaxTris];
    __shared__ uint32_t s_members[kFlatMaxTris];
    if (threadIdx.x < static_cast<uint32_t>(kFlatMaxTris)) s_members[threadIdx.x] = threadIdx.x < ngroups ? P.members[threadIdx.x] : 0u;
    for (uint32_t k = threadIdx.x; k < ntris; k += blockDim.x) {
        const uint32_t id = __ldg(&sc.prefs[first_ref + k]);
        s_id[k] = id;
        s_rec[0][k] = __ldg(sc.isect_hot + 2 * static_cast<size_t>(id));
        s_rec[1][k] = __ldg(sc.isect_hot + 2 * static_cast<size_t>(id) + 1);
        s_rec[2][k] = __ldg(sc.isect_cold + 2 * static_cast<size_t>(id));
        s_rec[3][k] = __ldg(sc.isect_cold + 2 * static_cast<size_t>(id) + 1);
    }
    __syncthreads();
    const uint32_t count = count_ptr ? *count_ptr : count_arg;
    float scale = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) scale = fmaxf(scale, fmaxf(fabsf(sc.lo[c]), fabsf(sc.hi[c])));
    // candidate bits that belong to real scan groups
    const uint32_t real0 = ngroups >= 32u ? 0xffffffffu : (1u << ngroups) - 1u;
    const uint32_t real1 = ngroups >= 64u ? 0xffffffffu : (ngroups > 32u ? (1u << (ngroups - 32u)) - 1u : 0u);

    // every ray costs the same (all triangles are scanned): a static grid-stride split needs no work cursor
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x) {
        float ox, oy, oz, dx, dy, dz, tmax_any = 0.f;
        if (MODE == 2) {
            ox = po[3 * idx]; oy = po[3 * idx + 1]; oz = po[3 * idx + 2];
            dx = pd[3 * idx]; dy = pd[3 * idx + 1]; dz = pd[3 * idx + 2];
        } else {
            const float4 a = __ldcs(&ra[idx]);
            const float4 b = __ldcs(&rb[idx]);
            ox = a.x; oy = a.y; oz = a.z; dx = a.w; dy = b.x; dz = b.y;
            if (ANY) tmax_any = b.z;
        }
        bool valid;
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
            valid = !(t1 < t0);
        }
        // Error bounds of the pre-filter (u = 2^-24): |b - nom| <= E, |a - denom| <= F for the reference's nom = n.(v0 - o),
        // denom = n.d of EVERY triangle of the group. traverse_pooled.cuh derives 11 u (3 S + |o|_1) and 6 u |d|_1 for a
        // triangle's own plane; a pair's partner adds 8 u (S + |o|_1) and 8 u |d|_1: E = 48 u (...), F = 24 u |d|_1.
        const float E = 2.8610229e-6f * (3.f * scale + (fabsf(ox) + fabsf(oy) + fabsf(oz)));
        const float F = 1.4305115e-6f * (fabsf(dx) + fabsf(dy) + fabsf(dz));
        // Approximate hit point h = o + (B / A) d. Against the reference's r = nom / denom (both bounds above hold, signs aligned):
        //   |B / A - r| = |(nom - B) denom - nom (denom - A)| / (|denom| A) <= (E + |r| F) / A
        // and a hit the reference accepts lies in the triangle, so |r| max|d_i| <= max|o_i| + scale: every coordinate of h is
        // within  (E max|d_i| + (max|o_i| + scale) F) / A  of the accepted hit point, plus the rounding of the reciprocal, the
        // product and the three FMAs (a few ulp of |o| + scale). The group's box is grown by kFlatBoxGrow x scale; the box test
        // is TRUSTED only where that bound stays below `tol` = half of the growth minus the rounding allowance, i.e. for
        // A >= Amin -- a per-ray constant. A grazing ray (and a ray so far away that tol <= 0) keeps the group instead.
        const float dmax = fmaxf(fabsf(dx), fmaxf(fabsf(dy), fabsf(dz)));
        const float omax = fmaxf(fabsf(ox), fmaxf(fabsf(oy), fabsf(oz)));
        const float tol = 0.5f * kFlatBoxGrow * scale - 1e-6f * (omax + 2.f * scale);
        const float Amin = tol > 0.f ? (E * dmax + (omax + scale) * F) / tol : kFltMax; // > F
        const float limE = ANY ? fmaf(tmax_any, F, E) : 0.f; // B <= lim (A + F) + E  <=>  B <= fma(lim, A, limE)

        uint32_t m0 = 0, m1 = 0; // candidate bits of triangles 0..31 / 32..63
        const float negE = -E;
        // one group of the scan: sets `bit` in `m` iff  ((A <= F) | in_range) & (A < Amin | inside).  The predicate logic is
        // written as PTX so that every compare folds its AND / OR into the FSETP (the C++ form compiles to a SEL per term).
        auto scan = [&](int k, uint32_t& m, uint32_t bit) {
            const FlatTri& T = P.t[k];
            // (one constant-bank operand per instruction: b = dp - n.o rather than a chain that starts from dp)
            const float a = fmaf(T.nx, dx, fmaf(T.ny, dy, T.nz * dz));
            const float b = T.dp - fmaf(T.nx, ox, fmaf(T.ny, oy, T.nz * oz));
            const float A = fabsf(a);
            const float B = __uint_as_float(__float_as_uint(b) ^ (__float_as_uint(a) & 0x80000000u));
            float rcp;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(A));
            const float rr = B * rcp; // A == 0 -> inf or NaN: untrusted below
            const float hx = fmaf(rr, dx, ox), hy = fmaf(rr, dy, oy), hz = fmaf(rr, dz, oz);
            // 0 <= nom/denom <= lim  ==>  A <= F  or  (B + E >= 0  and  B - E <= lim (A + F))   (traverse_pooled.cuh)
#define TRN_FLAT_BOX_PRED                                                                                                  \
    "setp.ge.f32 p, %1, %4;\n\t"        /* inside the grown box */                                                          \
    "setp.ge.and.f32 p, %2, %5, p;\n\t"                                                                                     \
    "setp.ge.and.f32 p, %3, %6, p;\n\t"                                                                                     \
    "setp.le.and.f32 p, %1, %7, p;\n\t"                                                                                     \
    "setp.le.and.f32 p, %2, %8, p;\n\t"                                                                                     \
    "setp.le.and.f32 p, %3, %9, p;\n\t"                                                                                     \
    "setp.ltu.or.f32 p, %10, %11, p;\n\t" /* ... or the box test is not trusted: A < Amin (or NaN) */                                          \
    "setp.ge.f32 q, %12, %13;\n\t"        /* B >= -E */
            if constexpr (ANY) {
                const float blim = fmaf(tmax_any, A, limE); // B <= lim A + (lim F + E)
                asm("{\n\t.reg .pred p, q;\n\t" TRN_FLAT_BOX_PRED
                    "setp.le.and.f32 q, %12, %17, q;\n\t"
                    "setp.le.or.f32 q, %14, %15, q;\n\t" // ... or A <= F
                    "and.pred p, p, q;\n\t"
                    "@p or.b32 %0, %0, %16;\n\t}"
                    : "+r"(m)
                    : "f"(hx), "f"(hy), "f"(hz), "f"(T.blo[0]), "f"(T.blo[1]), "f"(T.blo[2]), "f"(T.bhi[0]), "f"(T.bhi[1]), "f"(T.bhi[2]),
                      "f"(A), "f"(Amin), "f"(B), "f"(negE), "f"(A), "f"(F), "r"(bit), "f"(blim));
            } else {
                asm("{\n\t.reg .pred p, q;\n\t" TRN_FLAT_BOX_PRED
                    "setp.le.or.f32 q, %14, %15, q;\n\t" // ... or A <= F
                    "and.pred p, p, q;\n\t"
                    "@p or.b32 %0, %0, %16;\n\t}"
                    : "+r"(m)
                    : "f"(hx), "f"(hy), "f"(hz), "f"(T.blo[0]), "f"(T.blo[1]), "f"(T.blo[2]), "f"(T.bhi[0]), "f"(T.bhi[1]), "f"(T.bhi[2]),
                      "f"(A), "f"(Amin), "f"(B), "f"(negE), "f"(A), "f"(F), "r"(bit));
            }
        };
        if (valid) {
#pragma unroll
            for (int k = 0; k < (NT < 32 ? NT : 32); ++k) scan(k, m0, 1u << k);
#pragma unroll
            for (int k = 32; k < NT; ++k) scan(k, m1, 1u << (k - 32));
            m0 &= real0;
            m1 &= real1;
        }

        // EXACT: the lane's own candidate groups; an exact tie in r goes to the triangle visited first (the lower leaf slot)
        uint32_t best_id = kMiss, best_k = 0;
        float best_r = kFltMax, best_s = 0.f, best_t = 0.f;
        bool occluded = false;
        while ((m0 | m1) != 0u) {
            uint32_t g;
            if (m0 != 0u) {
                g = __ffs(m0) - 1;
                m0 &= m0 - 1u;
            } else {
                g = 32u + (__ffs(m1) - 1);
                m1 &= m1 - 1u;
            }
            uint32_t kk = s_members[g];
            do {
                const uint32_t k = (kk & 0xffu) - 1u;
                kk >>= 8;
                const float4 q0 = s_rec[0][k], q1 = s_rec[1][k];
                const float nx = q0.w, ny = q1.x, nz = q1.y;
                const float denom = nx * dx + ny * dy + nz * dz; // intersect_ray_plane, lib/intersection.h:40-49
                const float nom = nx * (q0.x - ox) + ny * (q0.y - oy) + nz * (q0.z - oz);
                const float r = nom / denom;
                // r < 0 rejects (intersection.h:66); a hit only matters if it can beat the running minimum (lib/kdtree.cpp:591),
                // resp. lies within the light distance
                if (denom != 0.f && r >= 0.f && r <= (ANY ? tmax_any : best_r)) {
                    const float4 q2 = s_rec[2][k], q3 = s_rec[3][k];
                    const float wx = (ox + r * dx) - q0.x, wy = (oy + r * dy) - q0.y, wz = (oz + r * dz) - q0.z; // :70-71
                    const float ux = q1.z, uy = q1.w, uz = q2.x, vx = q2.y, vy = q2.z, vz = q2.w;
                    const float wv = wx * vx + wy * vy + wz * vz;
                    const float wu = wx * ux + wy * uy + wz * uz;
                    const float s = (q3.x * wv - q3.y * wu) / q3.w; // :78-86
                    if (!(s < 0.f)) {
                        const float t = (q3.x * wu - q3.z * wv) / q3.w;
                        if (!(t < 0.f || 1.f < s + t)) {
                            if (ANY) {
                                occluded = true;
                                m0 = 0u;
                                m1 = 0u;
                                kk = 0u;
                            } else if (r < best_r || k < best_k) { // (r == best_r here unless r < best_r)
                                best_r = r;
                                best_s = s;
                                best_t = t;
                                best_k = k;
                                best_id = s_id[k];
                            }
                        }
                    }
                }
            } while (kk != 0u);
        }
        if (ANY) {
            if (!occluded) accumulate(acc, __float_as_uint(__ldcs(&rb[idx]).w), __ldcs(&rc[idx]));
        } else {
            __stcs(&hits[idx], make_uint4(best_id, __float_as_uint(best_r), __float_as_uint(best_s), __float_as_uint(best_t)));
        }
    }
}

} // namespace trn

