// CPU emulation of one warp of trace_pooled_kernel<0> (traverse_pooled.cuh), lane by lane in lockstep, with the same
// control flow, queues and arithmetic (fmaf for the pre-filter, unfused fp32 for the exact part). Development tool:
// finds logic errors (hangs, lost hits) and checks the pre-filter's conservativeness against the plain per-ray
// traversal on the benchmark mesh without spending GPU time.
//
//   g++ -O2 -std=c++17 -ffp-contract=off -fopenmp tools/pooled_emul.cpp turner_b200/csrc/kdtree_build.o -o /tmp/sim/pooled_emul
//   /tmp/sim/pooled_emul /tmp/sim/mesh1m.bin [rows] [device-built tree | -] [1 = also compare with the exhaustive search]
// The optional third argument is a pair-layout tree written by the device builder under TRN_KD_DUMP=<file>
// (kdtree_build_gpu.cu): the emulated warp then walks the production tree of the benchmark instead of the host-built one.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../turner_b200/csrc/kdtree_build.h"

using namespace trn;

static const float kFltMax = 3.402823466e+38f, kCellSlack = 1e-4f;
static const uint32_t kMiss = 0x40000000u;
static const int kPqSteps = 2, kPqLeaves = 64, kPqLeafMaxRefs = 4096, kPqChunkTris = 4, kPqSurv = 32 + 32 * 4;

struct Ray {
    float o[3], d[3];
    float tmax = 0.f; // any-hit (shadow) queries: inclusive distance to the light
};
struct Hit {
    uint32_t id = kMiss;
    float r = kFltMax, s = 0, t = 0;
};
struct Scene {
    HostTriangles tris;
    KdTree tree;
    std::vector<float> planes;
};
static inline uint32_t fbits(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
static inline float bfloat(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

// plain per-ray traversal (= traverse_pairs<false> for rays without zero direction components)
static void trace_plain(const Scene& sc, const Ray& ray, Hit& out) {
    const float ox = ray.o[0], oy = ray.o[1], oz = ray.o[2], dx = ray.d[0], dy = ray.d[1], dz = ray.d[2];
    const float ix = 1 / dx, iy = 1 / dy, iz = 1 / dz;
    const float* box = sc.tree.box;
    float tx1 = (box[0] - ox) * ix, tx2 = (box[3] - ox) * ix;
    float tenter = std::fmin(tx1, tx2), texit = std::fmax(tx1, tx2);
    float ty1 = (box[1] - oy) * iy, ty2 = (box[4] - oy) * iy;
    tenter = std::fmax(tenter, std::fmin(ty1, ty2));
    texit = std::fmin(texit, std::fmax(ty1, ty2));
    float tz1 = (box[2] - oz) * iz, tz2 = (box[5] - oz) * iz;
    tenter = std::fmax(tenter, std::fmin(tz1, tz2));
    texit = std::fmin(texit, std::fmax(tz1, tz2));
    out = Hit();
    if (texit < tenter) return;
    if (tenter < 0.f) tenter = 0.f;
    struct E {
        uint32_t x, y;
        float a, b;
    } stack[64];
    int sp = 0;
    const uint64_t* pn = sc.tree.pair_nodes.data();
    uint32_t nx_ = uint32_t(pn[0]), ny_ = uint32_t(pn[0] >> 32);
    for (;;) {
        while ((ny_ & 3u) != 3u) {
            const int ax = int(ny_ & 3u);
            const float split = bfloat(nx_);
            const uint32_t ci = ny_ >> 2;
            const uint32_t px = uint32_t(pn[ci]), py = uint32_t(pn[ci] >> 32), pz = uint32_t(pn[ci + 1]), pw = uint32_t(pn[ci + 1] >> 32);
            const float o_ax = ax == 0 ? ox : (ax == 1 ? oy : oz), i_ax = ax == 0 ? ix : (ax == 1 ? iy : iz);
            const float t = (split - o_ax) * i_ax;
            const bool flip = std::signbit(i_ax);
            const uint32_t nearx = flip ? pz : px, neary = flip ? pw : py, farx = flip ? px : pz, fary = flip ? py : pw;
            const bool near_only = texit < t, far_only = !near_only && (t < tenter), both = !near_only && !far_only;
            const bool go_far = far_only || (both && neary == 3u);
            if (both && neary != 3u && fary != 3u) stack[sp++] = E{farx, fary, t, texit};
            nx_ = go_far ? farx : nearx;
            ny_ = go_far ? fary : neary;
            const float te = (both && go_far) ? t : tenter, tx = (both && !go_far) ? t : texit;
            tenter = te;
            texit = tx;
        }
        const uint32_t first = nx_, cnt = ny_ >> 2;
        const float r_lo = tenter - kCellSlack * (std::fabs(tenter) + 1.f), r_hi = texit + kCellSlack * (std::fabs(texit) + 1.f);
        for (uint32_t i = 0; i < cnt; ++i) {
            const uint32_t id = sc.tree.pair_leaf_refs[first + i];
            const float* q = &sc.tris.isect[size_t(id) * 16];
            const float nx = q[3], ny = q[4], nz = q[5];
            const float denom = nx * dx + ny * dy + nz * dz;
            if (denom == 0.f) continue;
            const float nom = nx * (q[0] - ox) + ny * (q[1] - oy) + nz * (q[2] - oz);
            const float r = nom / denom;
            if (!(r >= 0.f) || !(r < out.r) || !(r >= r_lo && r <= r_hi)) continue;
            const float wx = (ox + r * dx) - q[0], wy = (oy + r * dy) - q[1], wz = (oz + r * dz) - q[2];
            const float wv = wx * q[9] + wy * q[10] + wz * q[11];
            const float wu = wx * q[6] + wy * q[7] + wz * q[8];
            const float s = (q[12] * wv - q[13] * wu) / q[15];
            if (s < 0.f) continue;
            const float t = (q[12] * wu - q[14] * wv) / q[15];
            if (t < 0.f || 1.f < s + t) continue;
            out.id = id;
            out.r = r;
            out.s = s;
            out.t = t;
        }
        if (out.id != kMiss && out.r <= texit) break;
        if (sp == 0) break;
        const E e = stack[--sp];
        nx_ = e.x;
        ny_ = e.y;
        tenter = e.a;
        texit = e.b;
    }
}

// plain per-ray any-hit traversal (= traverse_pairs<true>): occluded <=> an accepted triangle with 0 <= r <= tmax
static bool occluded_plain(const Scene& sc, const Ray& ray) {
    const float ox = ray.o[0], oy = ray.o[1], oz = ray.o[2], dx = ray.d[0], dy = ray.d[1], dz = ray.d[2], tmax = ray.tmax;
    const float ix = 1 / dx, iy = 1 / dy, iz = 1 / dz;
    const float* box = sc.tree.box;
    float tx1 = (box[0] - ox) * ix, tx2 = (box[3] - ox) * ix;
    float tenter = std::fmin(tx1, tx2), texit = std::fmax(tx1, tx2);
    float ty1 = (box[1] - oy) * iy, ty2 = (box[4] - oy) * iy;
    tenter = std::fmax(tenter, std::fmin(ty1, ty2));
    texit = std::fmin(texit, std::fmax(ty1, ty2));
    float tz1 = (box[2] - oz) * iz, tz2 = (box[5] - oz) * iz;
    tenter = std::fmax(tenter, std::fmin(tz1, tz2));
    texit = std::fmin(texit, std::fmax(tz1, tz2));
    if (texit < tenter) return false;
    if (tenter < 0.f) tenter = 0.f;
    struct E {
        uint32_t x, y;
        float a, b;
    } stack[64];
    int sp = 0;
    const uint64_t* pn = sc.tree.pair_nodes.data();
    uint32_t nx_ = uint32_t(pn[0]), ny_ = uint32_t(pn[0] >> 32);
    for (;;) {
        while ((ny_ & 3u) != 3u) {
            const int ax = int(ny_ & 3u);
            const float split = bfloat(nx_);
            const uint32_t ci = ny_ >> 2;
            const uint32_t px = uint32_t(pn[ci]), py = uint32_t(pn[ci] >> 32), pz = uint32_t(pn[ci + 1]), pw = uint32_t(pn[ci + 1] >> 32);
            const float o_ax = ax == 0 ? ox : (ax == 1 ? oy : oz), i_ax = ax == 0 ? ix : (ax == 1 ? iy : iz);
            const float t = (split - o_ax) * i_ax;
            const bool flip = std::signbit(i_ax);
            const uint32_t nearx = flip ? pz : px, neary = flip ? pw : py, farx = flip ? px : pz, fary = flip ? py : pw;
            const bool near_only = texit < t, far_only = !near_only && (t < tenter), both = !near_only && !far_only;
            const bool go_far = far_only || (both && neary == 3u);
            if (both && neary != 3u && fary != 3u) stack[sp++] = E{farx, fary, t, texit};
            nx_ = go_far ? farx : nearx;
            ny_ = go_far ? fary : neary;
            const float te = (both && go_far) ? t : tenter, tx = (both && !go_far) ? t : texit;
            tenter = te;
            texit = tx;
        }
        const uint32_t first = nx_, cnt = ny_ >> 2;
        const float r_lo = tenter - kCellSlack * (std::fabs(tenter) + 1.f), r_hi = texit + kCellSlack * (std::fabs(texit) + 1.f);
        for (uint32_t i = 0; i < cnt; ++i) {
            const uint32_t id = sc.tree.pair_leaf_refs[first + i];
            const float* q = &sc.tris.isect[size_t(id) * 16];
            const float nx = q[3], ny = q[4], nz = q[5];
            const float denom = nx * dx + ny * dy + nz * dz;
            if (denom == 0.f) continue;
            const float nom = nx * (q[0] - ox) + ny * (q[1] - oy) + nz * (q[2] - oz);
            const float r = nom / denom;
            if (!(r >= 0.f) || !(r <= tmax) || !(r >= r_lo && r <= r_hi)) continue;
            const float wx = (ox + r * dx) - q[0], wy = (oy + r * dy) - q[1], wz = (oz + r * dz) - q[2];
            const float wv = wx * q[9] + wy * q[10] + wz * q[11];
            const float wu = wx * q[6] + wy * q[7] + wz * q[8];
            const float s = (q[12] * wv - q[13] * wu) / q[15];
            if (s < 0.f) continue;
            const float t = (q[12] * wu - q[14] * wv) / q[15];
            if (t < 0.f || 1.f < s + t) continue;
            return true;
        }
        if (sp == 0) return false;
        const E e = stack[--sp];
        nx_ = e.x;
        ny_ = e.y;
        tenter = e.a;
        texit = e.b;
        if (tenter - kCellSlack * (std::fabs(tenter) + 1.f) > tmax) return false;
    }
}

struct U4 {
    uint32_t x, y, z, w;
};
struct F4 {
    float x, y, z, w;
};

// exhaustive closest hit over ALL triangles with the exact sequence (no tree): what any correct tree must return, up to which
// of several triangles reports an exact tie in r
static void trace_brute(const Scene& sc, const Ray& ray, Hit& out) {
    const float ox = ray.o[0], oy = ray.o[1], oz = ray.o[2], dx = ray.d[0], dy = ray.d[1], dz = ray.d[2];
    out = Hit();
    for (uint32_t id = 0; id < sc.tris.count; ++id) {
        const float* q = &sc.tris.isect[size_t(id) * 16];
        const float nx = q[3], ny = q[4], nz = q[5];
        const float denom = nx * dx + ny * dy + nz * dz;
        if (denom == 0.f) continue;
        const float nom = nx * (q[0] - ox) + ny * (q[1] - oy) + nz * (q[2] - oz);
        const float r = nom / denom;
        if (!(r >= 0.f) || !(r < out.r)) continue;
        const float wx = (ox + r * dx) - q[0], wy = (oy + r * dy) - q[1], wz = (oz + r * dz) - q[2];
        const float wv = wx * q[9] + wy * q[10] + wz * q[11];
        const float wu = wx * q[6] + wy * q[7] + wz * q[8];
        const float s = (q[12] * wv - q[13] * wu) / q[15];
        if (s < 0.f) continue;
        const float t = (q[12] * wu - q[14] * wv) / q[15];
        if (t < 0.f || 1.f < s + t) continue;
        out.id = id;
        out.r = r;
        out.s = s;
        out.t = t;
    }
}

// one emulated warp over rays[begin, end); returns false on a detected hang
// any: MODE 1 of the kernel (shadow rays: occluded <=> an accepted triangle with 0 <= r <= tmax); hits[i].id is then 0 for
// occluded, kMiss for unoccluded
static bool warp_run(const Scene& sc, const std::vector<Ray>& rays, size_t begin, size_t end, std::vector<Hit>& hits, int refill_below,
                     int walk_iters, int leaf_gate, uint64_t* stat, bool any = false) {
    F4 s_ray[64];
    U4 s_leaf[kPqLeaves];
    uint32_t s_surv[kPqSurv][2];
    uint32_t s_nleaf = 0, s_nsurv = 0;
    float scale = 0.f;
    for (int c = 0; c < 6; ++c) scale = std::fmax(scale, std::fabs(sc.tree.box[c]));
    const uint64_t* pn = sc.tree.pair_nodes.data();
    struct LaneS {
        U4 stack[64];
        int sp = 0;
        float ox, oy, oz, ix, iy, iz, tenter, texit, last_texit;
        uint32_t nx, ny;
        uint32_t best_id, best_seq, idx;
        float best_r, best_s, best_t;
        bool busy = false, walking = false, occluded = false;
        float tmax = 0.f;
    };
    std::vector<LaneS> L(32);
    size_t next = begin;
    uint64_t cycles = 0;
    for (;;) {
        if (++cycles > 50000000ull) return false;
        int nbusy = 0;
        for (auto& l : L) nbusy += l.busy;
        if (nbusy < refill_below && next < end) {
            for (int lane = 0; lane < 32 && next < end; ++lane) {
                LaneS& l = L[lane];
                if (l.busy) continue;
                l.idx = uint32_t(next++);
                const Ray& ry = rays[l.idx];
                const float dx = ry.d[0], dy = ry.d[1], dz = ry.d[2];
                l.ox = ry.o[0];
                l.oy = ry.o[1];
                l.oz = ry.o[2];
                if (dx == 0.f || dy == 0.f || dz == 0.f) {
                    std::fprintf(stderr, "axis-parallel ray skipped\n");
                    continue;
                }
                l.ix = 1 / dx;
                l.iy = 1 / dy;
                l.iz = 1 / dz;
                const float* box = sc.tree.box;
                float tx1 = (box[0] - l.ox) * l.ix, tx2 = (box[3] - l.ox) * l.ix;
                float t0 = std::fmin(tx1, tx2), t1 = std::fmax(tx1, tx2);
                float ty1 = (box[1] - l.oy) * l.iy, ty2 = (box[4] - l.oy) * l.iy;
                t0 = std::fmax(t0, std::fmin(ty1, ty2));
                t1 = std::fmin(t1, std::fmax(ty1, ty2));
                float tz1 = (box[2] - l.oz) * l.iz, tz2 = (box[5] - l.oz) * l.iz;
                t0 = std::fmax(t0, std::fmin(tz1, tz2));
                t1 = std::fmin(t1, std::fmax(tz1, tz2));
                if (t1 < t0) {
                    hits[l.idx] = Hit();
                    continue;
                }
                l.tenter = t0 < 0.f ? 0.f : t0;
                l.texit = t1;
                l.sp = 0;
                l.nx = uint32_t(pn[0]);
                l.ny = uint32_t(pn[0] >> 32);
                l.best_id = kMiss;
                l.best_r = kFltMax;
                l.best_s = l.best_t = 0.f;
                l.best_seq = 0;
                l.last_texit = -kFltMax;
                                l.busy = true;
                l.occluded = false;
                l.tmax = ry.tmax;
                l.walking = !(any && l.tenter - kCellSlack * (std::fabs(l.tenter) + 1.f) > l.tmax);
                const float E = 1.9073486e-6f * (3.f * scale + (std::fabs(l.ox) + std::fabs(l.oy) + std::fabs(l.oz)));
                const float F = 9.5367432e-7f * (std::fabs(dx) + std::fabs(dy) + std::fabs(dz));
                s_ray[2 * lane] = F4{l.ox, l.oy, l.oz, E};
                s_ray[2 * lane + 1] = F4{dx, dy, dz, F};
            }
            nbusy = 0;
            for (auto& l : L) nbusy += l.busy;
        }
        if (nbusy == 0) {
            if (next >= end) break;
            continue;
        }
        // WALK (gated leaf block, as in the kernel)
        bool blocked[32] = {false};
        for (int it = 0;; ) {
            bool can[32], at_leaf[32];
            int n_leaf = 0, n_inner = 0;
            for (int lane = 0; lane < 32; ++lane) {
                can[lane] = L[lane].busy && L[lane].walking && !blocked[lane];
                at_leaf[lane] = can[lane] && (L[lane].ny & 3u) == 3u;
                n_leaf += at_leaf[lane];
                n_inner += can[lane] && !at_leaf[lane];
            }
            const bool last = it >= walk_iters || n_inner == 0;
            if (n_leaf != 0 && (last || n_leaf >= leaf_gate)) {
                for (int lane = 0; lane < 32; ++lane) {
                    if (!at_leaf[lane]) continue;
                    LaneS& l = L[lane];
                    const uint32_t cnt = l.ny >> 2;
                    uint32_t take = 0;
                    if (cnt > 0u) {
                        const uint32_t slot = s_nleaf++;
                        if (slot < uint32_t(kPqLeaves)) {
                            take = std::min<uint32_t>(cnt, kPqLeafMaxRefs);
                            float lo = l.tenter - kCellSlack * (std::fabs(l.tenter) + 1.f);
                            float hi = l.texit + kCellSlack * (std::fabs(l.texit) + 1.f);
                            lo = std::fmax(lo, 0.f);
                            hi = any ? std::fmin(hi, l.tmax) : std::fmin(hi, l.best_r);
                            s_leaf[slot] = U4{l.nx, take | (uint32_t(lane) << 24), fbits(lo), fbits(hi)};
                        }
                    }
                    if (take < cnt) { // queue full, or a leaf longer than one descriptor: continue from the node register
                        l.nx += take;
                        l.ny -= take << 2;
                        blocked[lane] = true;
                    } else {
                        l.last_texit = l.texit;
                        if (l.sp == 0) l.walking = false;
                        else {
                            const U4 e = l.stack[--l.sp];
                            l.nx = e.x;
                            l.ny = e.y;
                            l.tenter = bfloat(e.z);
                            l.texit = bfloat(e.w);
                            if (any ? l.tenter - kCellSlack * (std::fabs(l.tenter) + 1.f) > l.tmax : l.tenter > l.best_r) l.walking = false;
                        }
                    }
                }
            }
            if (last) break;
            ++it;
            stat[0]++;
            for (int lane = 0; lane < 32; ++lane) { // what the 32 lanes do in this iteration (after the leaf block)
                const LaneS& l = L[lane];
                const bool leaf = (l.ny & 3u) == 3u;
                const int k = !l.busy ? 0 : (!l.walking ? 1 : (blocked[lane] ? 2 : (leaf ? ((l.ny >> 2) == 0u ? 3 : 4) : 5)));
                stat[8 + k]++;
            }
            for (int lane = 0; lane < 32; ++lane) {
                if (!(can[lane] && !at_leaf[lane])) continue;
                LaneS& l = L[lane];
                for (int rep = 0; rep < kPqSteps && (l.ny & 3u) != 3u; ++rep) { // TRN_PQ_STEPS inner steps per iteration
                stat[1]++;
                const uint32_t ax = l.ny & 3u;
                const float split = bfloat(l.nx);
                const uint32_t ci = l.ny >> 2;
                const uint32_t px = uint32_t(pn[ci]), py = uint32_t(pn[ci] >> 32), pz = uint32_t(pn[ci + 1]), pw = uint32_t(pn[ci + 1] >> 32);
                float o_ax = l.oz, i_ax = l.iz;
                if (ax == 0u) { o_ax = l.ox; i_ax = l.ix; }
                if (ax == 1u) { o_ax = l.oy; i_ax = l.iy; }
                const float t = (split - o_ax) * i_ax;
                const bool flip = (fbits(i_ax) >> 31) != 0u;
                const uint32_t nearx = flip ? pz : px, neary = flip ? pw : py, farx = flip ? px : pz, fary = flip ? py : pw;
                const bool near_only = l.texit < t, far_only = !near_only && (t < l.tenter), both = !near_only && !far_only;
                const bool go_far = far_only || (both && neary == 3u);
                if (both && neary != 3u && fary != 3u) l.stack[l.sp++] = U4{farx, fary, fbits(t), fbits(l.texit)};
                l.nx = go_far ? farx : nearx;
                l.ny = go_far ? fary : neary;
                const float te = (both && go_far) ? t : l.tenter, tx = (both && !go_far) ? t : l.texit;
                l.tenter = te;
                l.texit = tx;
                }
            }
        }
        // TEST / EXACT: leaves 32 at a time, their chunks of 4 references dealt out 32 per round
        const uint32_t nleaf = std::min<uint32_t>(s_nleaf, kPqLeaves);
        auto exact_round = [&]() {
            const uint32_t ns = s_nsurv;
            const uint32_t take = std::min(32u, ns), sbase = ns - take;
            stat[4]++;
            stat[5] += take;
            struct P {
                bool pass;
                uint32_t id, owner, seq;
                float r, s, t;
            } res[32];
            for (uint32_t lane = 0; lane < 32; ++lane) {
                P& p = res[lane];
                p.pass = false;
                if (lane >= take) continue;
                p.id = s_surv[sbase + lane][0];
                p.owner = s_surv[sbase + lane][1] & 31u;
                p.seq = s_surv[sbase + lane][1] >> 5;
                const float lim = any ? L[p.owner].tmax : L[p.owner].best_r;
                const F4 ro = s_ray[2 * p.owner], rd = s_ray[2 * p.owner + 1];
                const float* q = &sc.tris.isect[size_t(p.id) * 16];
                const float nx = q[3], ny = q[4], nz = q[5];
                const float denom = nx * rd.x + ny * rd.y + nz * rd.z;
                const float nom = nx * (q[0] - ro.x) + ny * (q[1] - ro.y) + nz * (q[2] - ro.z);
                p.r = nom / denom;
                if (denom != 0.f && p.r >= 0.f && p.r <= lim) {
                    const float wx = (ro.x + p.r * rd.x) - q[0], wy = (ro.y + p.r * rd.y) - q[1], wz = (ro.z + p.r * rd.z) - q[2];
                    const float wv = wx * q[9] + wy * q[10] + wz * q[11];
                    const float wu = wx * q[6] + wy * q[7] + wz * q[8];
                    p.s = (q[12] * wv - q[13] * wu) / q[15];
                    if (!(p.s < 0.f)) {
                        p.t = (q[12] * wu - q[14] * wv) / q[15];
                        p.pass = !(p.t < 0.f || 1.f < p.s + p.t);
                    }
                }
            }
            for (uint32_t lane = 0; lane < 32; ++lane) {
                const P& p = res[lane];
                if (!p.pass) continue;
                LaneS& o = L[p.owner];
                if (any) {
                    o.occluded = true;
                } else if (p.r < o.best_r || (p.r == o.best_r && p.seq < o.best_seq)) {
                    o.best_id = p.id;
                    o.best_r = p.r;
                    o.best_s = p.s;
                    o.best_t = p.t;
                    o.best_seq = p.seq;
                }
            }
            s_nsurv = sbase;
        };
        for (uint32_t lb = 0; lb < nleaf; lb += 32) {
            uint32_t P[32], total = 0; // inclusive chunk totals of the batch's leaves
            for (uint32_t j = 0; j < 32; ++j) {
                const uint32_t c = lb + j < nleaf ? (s_leaf[lb + j].y & 0xffffffu) : 0u;
                total += (c + kPqChunkTris - 1) / kPqChunkTris;
                P[j] = total;
            }
            for (uint32_t base = 0; base < total; base += 32) {
                while (s_nsurv >= 32u) exact_round();
                stat[2]++;
                uint32_t km[32], owner_[32], seq0[32], ids[32][4];
                for (uint32_t lane = 0; lane < 32; ++lane) {
                    km[lane] = 0;
                    const uint32_t g = base + lane;
                    if (g >= total) continue;
                    stat[3]++;
                    uint32_t j = 0;
                    while (P[j] <= g) ++j;
                    const U4 d = s_leaf[lb + j];
                    const uint32_t lcnt = d.y & 0xffffffu, owner = d.y >> 24;
                    const uint32_t sub = g - (P[j] - (lcnt + kPqChunkTris - 1) / kPqChunkTris), off0 = sub * kPqChunkTris;
                    const uint32_t cnt = std::min<uint32_t>(kPqChunkTris, lcnt - off0);
                    const float lo = bfloat(d.z), hi = bfloat(d.w);
                    const F4 ro = s_ray[2 * owner], rd = s_ray[2 * owner + 1];
                    const float E = ro.w, F = rd.w;
                    const float c1 = std::fmaf(-lo, F, -E), c2 = std::fmaf(hi, F, E);
                    owner_[lane] = owner;
                    seq0[lane] = ((lb + j) << 20) + off0 + 1u;
                    for (uint32_t k = 0; k < cnt; ++k) {
                        const uint32_t id = sc.tree.pair_leaf_refs[d.x + off0 + k];
                        ids[lane][k] = id;
                        const float* p = &sc.planes[size_t(id) * 4];
                        const float a = std::fmaf(p[0], rd.x, std::fmaf(p[1], rd.y, p[2] * rd.z));
                        const float b = std::fmaf(-p[0], ro.x, std::fmaf(-p[1], ro.y, std::fmaf(-p[2], ro.z, p[3])));
                        const float A = std::fabs(a);
                        const float B = bfloat(fbits(b) ^ (fbits(a) & 0x80000000u));
                        stat[6]++;
                        if (A <= F || (B >= std::fmaf(lo, A, c1) && B <= std::fmaf(hi, A, c2))) km[lane] |= 1u << k;
                    }
                }
                for (uint32_t k = 0; k < uint32_t(kPqChunkTris); ++k) // k-major append, as the kernel's four ballots
                    for (uint32_t lane = 0; lane < 32; ++lane)
                        if (km[lane] & (1u << k)) {
                            if (s_nsurv >= uint32_t(kPqSurv)) {
                                std::fprintf(stderr, "survivor queue overflow\n");
                                return false;
                            }
                            s_surv[s_nsurv][0] = ids[lane][k];
                            s_surv[s_nsurv][1] = owner_[lane] | ((seq0[lane] + k) << 5);
                            ++s_nsurv;
                        }
            }
        }
        while (s_nsurv > 0u) exact_round();
        s_nleaf = 0;
        for (int lane = 0; lane < 32; ++lane) {
            LaneS& l = L[lane];
            if (!l.busy) continue;
            l.best_seq = 0;
            const bool finished = any ? (l.occluded || !l.walking)
                                      : (!l.walking || (!blocked[lane] && l.best_id != kMiss && (l.best_r <= l.last_texit || l.best_r < l.tenter)));
            if (finished) {
                l.busy = false;
                Hit h;
                h.id = any ? (l.occluded ? 0u : kMiss) : l.best_id;
                h.r = l.best_r;
                h.s = l.best_s;
                h.t = l.best_t;
                hits[l.idx] = h;
            }
        }
    }
    return true;
}

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "/tmp/sim/mesh1m.bin";
    const int rows = argc > 2 ? std::atoi(argv[2]) : 16;
    const bool brute = argc > 4 && std::atoi(argv[4]) != 0; // also check the per-ray traversal against the exhaustive search
    FILE* f = std::fopen(path, "rb");
    if (!f) return 1;
    uint32_t n = 0;
    if (std::fread(&n, 4, 1, f) != 1) return 1;
    std::vector<float> V(size_t(n) * 9), N(size_t(n) * 9), D(size_t(n) * 4, 0.7f);
    if (std::fread(V.data(), 4, V.size(), f) != V.size() || std::fread(N.data(), 4, N.size(), f) != N.size()) return 1;
    std::fclose(f);
    Scene sc;
    precompute_triangles(V.data(), N.data(), D.data(), n, sc.tris);
    if (argc > 3 && std::strcmp(argv[3], "-") != 0) {
        // the production tree of a device build (TRN_KD_DUMP): u64 pair-node count, u64 reference count, nodes, references
        FILE* g = std::fopen(argv[3], "rb");
        uint64_t hdr[2] = {0, 0};
        if (!g || std::fread(hdr, 8, 2, g) != 2 || hdr[0] < 2 || hdr[0] > (1ull << 31) || hdr[1] > (1ull << 32)) return 2;
        sc.tree.pair_nodes.resize(hdr[0]);
        sc.tree.pair_leaf_refs.resize(hdr[1]);
        if (std::fread(sc.tree.pair_nodes.data(), 8, hdr[0], g) != hdr[0] || std::fread(sc.tree.pair_leaf_refs.data(), 4, hdr[1], g) != hdr[1]) return 2;
        std::fclose(g);
        for (int c = 0; c < 3; ++c) { // scene box as KDTree::KDTree makes it: min / max over the vertices
            float lo = V[c], hi = V[c];
            for (size_t i = 0; i < size_t(n) * 3; ++i) {
                lo = std::fmin(lo, V[i * 3 + c]);
                hi = std::fmax(hi, V[i * 3 + c]);
            }
            sc.tree.box[c] = lo;
            sc.tree.box[3 + c] = hi;
        }
    } else {
        build_kdtree(sc.tris, sc.tree, 0);
    }
    sc.planes.resize(size_t(n) * 4);
    for (size_t i = 0; i < n; ++i) {
        const float* q = &sc.tris.isect[i * 16];
        sc.planes[i * 4] = q[3];
        sc.planes[i * 4 + 1] = q[4];
        sc.planes[i * 4 + 2] = q[5];
        sc.planes[i * 4 + 3] = float(double(q[3]) * q[0] + double(q[4]) * q[1] + double(q[5]) * q[2]);
    }
    std::printf("triangles %u, pair nodes %zu\n", n, sc.tree.pair_nodes.size());
    const float R[9] = {0.9438583850860596f, -0.07586748898029327f, 0.3215206563472748f, 0.0f, 0.9732714891433716f, 0.2296576052904129f,
                        -0.33035042881965637f, -0.21676425635814667f, 0.9186304211616516f};
    const float P[3] = {1.4881685972213745f, 1.0629775524139404f, 4.251910209655762f};
    const float delta = std::tan(0.4287780225276947f);
    const int W = 1920, H = 1920;
    std::vector<Ray> wave;
    const int y0 = H / 2 - rows / 2 - 200;
    for (int y = y0; y < y0 + rows; ++y)
        for (int x = 0; x < W; ++x) {
            const float px = x + 0.5f, py = y + 0.5f;
            const float vx = -delta * (1 - 2 * px / W), vy = delta * (1 - 2 * py / H), vz = -1.f;
            Ray r;
            for (int c = 0; c < 3; ++c) {
                r.o[c] = P[c];
                r.d[c] = R[3 * c] * vx + R[3 * c + 1] * vy + R[3 * c + 2] * vz;
            }
            wave.push_back(r);
        }
    std::mt19937_64 rng(1);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    const int m = 4;
    for (int depth = 0; depth < 3; ++depth) {
        std::vector<Hit> ref(wave.size()), got(wave.size());
#pragma omp parallel for schedule(dynamic, 256)
        for (long i = 0; i < long(wave.size()); ++i) trace_plain(sc, wave[i], ref[i]);
        const size_t per = 4096;
        const long nw = long((wave.size() + per - 1) / per);
        uint64_t stat[16] = {0};
        bool ok = true;
#pragma omp parallel for schedule(dynamic, 1)
        for (long w = 0; w < nw; ++w) {
            uint64_t st[16] = {0};
            const bool r = warp_run(sc, wave, size_t(w) * per, std::min(wave.size(), size_t(w + 1) * per), got, 28, 12, 10, st);
#pragma omp critical
            {
                ok &= r;
                for (int k = 0; k < 16; ++k) stat[k] += st[k];
            }
        }
        size_t id_diff = 0, bit_diff = 0, hits = 0;
        for (size_t i = 0; i < wave.size(); ++i) {
            hits += ref[i].id != kMiss;
            if (ref[i].id != got[i].id) ++id_diff;
            else if (ref[i].id != kMiss && (fbits(ref[i].r) != fbits(got[i].r) || fbits(ref[i].s) != fbits(got[i].s) || fbits(ref[i].t) != fbits(got[i].t))) ++bit_diff;
        }
        if (brute) {
            // the tree itself (a device-built one comes from outside): the per-ray traversal against the exhaustive search
            size_t bd = 0, ties = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : bd, ties)
            for (long i = 0; i < long(wave.size()); ++i) {
                Hit b;
                trace_brute(sc, wave[i], b);
                if (b.id == ref[i].id) continue;
                if (b.id != kMiss && ref[i].id != kMiss && fbits(b.r) == fbits(ref[i].r)) ++ties;
                else ++bd;
            }
            std::printf("brute %d: %zu rays, differences %zu, exact ties resolved differently %zu\n", depth, wave.size(), bd, ties);
        }
        const double nr = double(wave.size());
        std::printf("depth %d: %zu rays, %zu hits, completed %d, id differences %zu, (r,s,t) bit differences %zu\n", depth, wave.size(), hits,
                    int(ok), id_diff, bit_diff);
        std::printf("   per ray: walk iterations %.2f (lanes %.1f)  test rounds %.3f (chunks/round %.1f)  exact rounds %.3f (survivors/round %.1f)  "
                    "tests %.1f survivors %.2f\n",
                    stat[0] / nr, double(stat[1]) / std::max<uint64_t>(1, stat[0]), stat[2] / nr, double(stat[3]) / std::max<uint64_t>(1, stat[2]),
                    stat[4] / nr, double(stat[5]) / std::max<uint64_t>(1, stat[4]), stat[6] / nr, stat[5] / nr);
        {
            const double its = double(std::max<uint64_t>(1, stat[0]));
            std::printf("   lanes per walk iteration: idle %.1f, walk over %.1f, leaf continues next cycle %.1f, waiting at a void %.1f, "
                        "waiting at a leaf %.1f, stepping %.1f\n",
                        stat[8] / its, stat[9] / its, stat[10] / its, stat[11] / its, stat[12] / its, stat[13] / its);
        }
        {
            // shadow wave of this depth (pathtracer.cpp:44-58): from the offset hit point towards the light of
            // scenes.cubesphere(), inclusive tmax = distance to the light; any-hit mode of the kernel vs the plain traversal
            const float Lp[3] = {3.f, 4.f, 5.f};
            std::vector<Ray> sh;
            for (size_t i = 0; i < wave.size(); ++i) {
                if (ref[i].id == kMiss) continue;
                const float* q = &sc.tris.isect[size_t(ref[i].id) * 16];
                Ray r;
                float l2 = 0.f, q2 = 0.f, ld[3];
                for (int c = 0; c < 3; ++c) {
                    const float p = wave[i].o[c] + ref[i].r * wave[i].d[c];
                    r.o[c] = p + 0.0001f * q[3 + c];
                    ld[c] = Lp[c] - p;
                    l2 += ld[c] * ld[c];
                    q2 += (Lp[c] - r.o[c]) * (Lp[c] - r.o[c]);
                }
                const float inv = 1 / std::sqrt(l2);
                for (int c = 0; c < 3; ++c) r.d[c] = inv * ld[c];
                r.tmax = std::sqrt(q2);
                if (r.d[0] * q[3] + r.d[1] * q[4] + r.d[2] * q[5] > 0.f) sh.push_back(r); // lit side only, as shade_bounce_kernel
            }
            std::vector<Hit> sgot(sh.size());
            std::vector<uint8_t> sref(sh.size());
#pragma omp parallel for schedule(dynamic, 256)
            for (long i = 0; i < long(sh.size()); ++i) sref[i] = occluded_plain(sc, sh[i]);
            const long snw = long((sh.size() + per - 1) / per);
            bool sok = true;
#pragma omp parallel for schedule(dynamic, 1)
            for (long w = 0; w < snw; ++w) {
                uint64_t st[16] = {0};
                const bool r = warp_run(sc, sh, size_t(w) * per, std::min(sh.size(), size_t(w + 1) * per), sgot, 28, 12, 10, st, true);
#pragma omp critical
                sok &= r;
            }
            size_t sdiff = 0, socc = 0;
            for (size_t i = 0; i < sh.size(); ++i) {
                socc += sref[i];
                sdiff += (sgot[i].id != kMiss) != (sref[i] != 0);
            }
            std::printf("shadow %d: %zu rays, %zu occluded, completed %d, differences %zu\n", depth, sh.size(), socc, int(sok), sdiff);
        }
        std::vector<Ray> next;
        for (size_t base = 0; base < wave.size(); base += 32) {
            std::vector<size_t> hl;
            for (size_t i = base; i < std::min(wave.size(), base + 32); ++i)
                if (ref[i].id != kMiss) hl.push_back(i);
            const size_t nh = hl.size(), cbase = next.size();
            next.resize(cbase + nh * m);
            for (size_t rank = 0; rank < nh; ++rank) {
                const size_t i = hl[rank];
                const float* q = &sc.tris.isect[size_t(ref[i].id) * 16];
                float nrm[3] = {q[3], q[4], q[5]}, p2[3];
                for (int c = 0; c < 3; ++c) p2[c] = (wave[i].o[c] + ref[i].r * wave[i].d[c]) + 0.0001f * nrm[c];
                float a[3] = {std::fabs(nrm[0]) < 0.9f ? 1.f : 0.f, std::fabs(nrm[0]) < 0.9f ? 0.f : 1.f, 0.f};
                float t1[3] = {nrm[1] * a[2] - nrm[2] * a[1], nrm[2] * a[0] - nrm[0] * a[2], nrm[0] * a[1] - nrm[1] * a[0]};
                const float l1 = std::sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
                for (int c = 0; c < 3; ++c) t1[c] /= l1;
                float t2[3] = {nrm[1] * t1[2] - nrm[2] * t1[1], nrm[2] * t1[0] - nrm[0] * t1[2], nrm[0] * t1[1] - nrm[1] * t1[0]};
                for (int k = 0; k < m; ++k) {
                    const float u1 = U(rng), u2 = U(rng);
                    const float z = u1, rr = std::sqrt(std::max(0.f, 1.f - z * z)), phi = 6.2831853f * u2;
                    const float lx = rr * std::cos(phi), ly = rr * std::sin(phi);
                    Ray ch;
                    for (int c = 0; c < 3; ++c) {
                        ch.o[c] = p2[c];
                        ch.d[c] = t1[c] * lx + t2[c] * ly + nrm[c] * z;
                    }
                    next[cbase + k * nh + rank] = ch;
                }
            }
        }
        wave.swap(next);
        if (wave.empty()) break;
    }
    return 0;
}
