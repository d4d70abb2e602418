// Offline SIMT model of the closest-hit traversal schedules (development tool, CPU only, not part of the product).
//
// Builds the production kd layout for a mesh (same builder as the library), generates a primary wave for a band of
// image rows and its depth-1 / depth-2 child waves the way shade_bounce_kernel lays them out, records per ray the
// sequence of (inner steps, leaf size, plane-test survivors) of traverse_pairs<>, and then replays those sequences
// through warp-level models of different lane schedules, counting issued warp-instructions and active lanes.
// The instruction costs per step are read off the SASS of the shipped kernel (profiles/README.md).
//
//   g++ -O2 -std=c++17 -ffp-contract=off -pthread tools/simt_sim.cpp turner_b200/csrc/kdtree_build.o -o /tmp/sim/simt_sim
//   /tmp/sim/simt_sim /tmp/sim/mesh1m.bin [rows]
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <random>
#include <vector>

#include "../turner_b200/csrc/kdtree_build.h"

using namespace trn;

static const float kEpsDir = 0.00001f, kFltMax = 3.402823466e+38f, kCellSlack = 1e-4f;
static const uint32_t kMiss = 0x40000000u;

struct Ev {
    uint16_t inner, cnt, ncand;
};
struct RayRec {
    std::vector<Ev> ev;
    uint32_t id = kMiss;
    float r = 0, s = 0, t = 0;
};
struct Ray {
    float o[3], d[3];
};

struct Scene {
    HostTriangles tris;
    KdTree tree;
};

static inline float sel3(int ax, float x, float y, float z) { return ax == 0 ? x : (ax == 1 ? y : z); }

static void trace(const Scene& sc, const Ray& ray, RayRec& out) {
    const float ox = ray.o[0], oy = ray.o[1], oz = ray.o[2], dx = ray.d[0], dy = ray.d[1], dz = ray.d[2];
    const float fdx = dx == 0.f ? kEpsDir : dx, fdy = dy == 0.f ? kEpsDir : dy, fdz = dz == 0.f ? kEpsDir : dz;
    const float ix = 1 / fdx, iy = 1 / fdy, iz = 1 / fdz;
    const float* box = sc.tree.box;
    float tx1 = (box[0] - ox) * ix, tx2 = (box[3] - ox) * ix;
    float tenter = std::fmin(tx1, tx2), texit = std::fmax(tx1, tx2);
    float ty1 = (box[1] - oy) * iy, ty2 = (box[4] - oy) * iy;
    tenter = std::fmax(tenter, std::fmin(ty1, ty2));
    texit = std::fmin(texit, std::fmax(ty1, ty2));
    float tz1 = (box[2] - oz) * iz, tz2 = (box[5] - oz) * iz;
    tenter = std::fmax(tenter, std::fmin(tz1, tz2));
    texit = std::fmin(texit, std::fmax(tz1, tz2));
    out.ev.clear();
    out.id = kMiss;
    float best = kFltMax;
    if (texit < tenter) return;
    if (tenter < 0.f) tenter = 0.f;
    struct E {
        uint32_t x, y;
        float a, b;
    } stack[64];
    int sp = 0;
    const uint64_t* pn = sc.tree.pair_nodes.data();
    auto X = [&](uint64_t v) { return uint32_t(v & 0xffffffffu); };
    auto Y = [&](uint64_t v) { return uint32_t(v >> 32); };
    uint32_t nx_ = X(pn[0]), ny_ = Y(pn[0]);
    for (;;) {
        uint16_t inner = 0;
        while ((ny_ & 3u) != 3u) {
            ++inner;
            const int ax = int(ny_ & 3u);
            float split;
            std::memcpy(&split, &nx_, 4);
            const uint32_t ci = ny_ >> 2;
            const uint32_t px = X(pn[ci]), py = Y(pn[ci]), pz = X(pn[ci + 1]), pw = Y(pn[ci + 1]);
            const float o_ax = sel3(ax, ox, oy, oz), i_ax = sel3(ax, ix, iy, iz);
            const float t = (split - o_ax) * i_ax;
            const bool flip = std::signbit(i_ax);
            const uint32_t nearx = flip ? pz : px, neary = flip ? pw : py, farx = flip ? px : pz, fary = flip ? py : pw;
            const bool near_only = texit < t;
            const bool far_only = !near_only && (t < tenter);
            const bool both = !near_only && !far_only;
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
        uint16_t ncand = 0;
        for (uint32_t i = 0; i < cnt; ++i) {
            const uint32_t id = sc.tree.pair_leaf_refs[first + i];
            const float* q = &sc.tris.isect[size_t(id) * 16]; // v0 n u v uv vv uu denom
            const float nx = q[3], ny = q[4], nz = q[5];
            const float denom = nx * dx + ny * dy + nz * dz;
            if (denom == 0.f) continue;
            const float nom = nx * (q[0] - ox) + ny * (q[1] - oy) + nz * (q[2] - oz);
            const float r = nom / denom;
            if (!(r >= 0.f)) continue;
            if (!(r < best)) continue;
            if (!(r >= r_lo && r <= r_hi)) continue;
            ++ncand;
            const float wx = (ox + r * dx) - q[0], wy = (oy + r * dy) - q[1], wz = (oz + r * dz) - q[2];
            const float wv = wx * q[9] + wy * q[10] + wz * q[11];
            const float wu = wx * q[6] + wy * q[7] + wz * q[8];
            const float s = (q[12] * wv - q[13] * wu) / q[15];
            if (s < 0.f) continue;
            const float t = (q[12] * wu - q[14] * wv) / q[15];
            if (t < 0.f || 1.f < s + t) continue;
            out.id = id;
            best = r;
            out.r = r;
            out.s = s;
            out.t = t;
        }
        out.ev.push_back(Ev{inner, uint16_t(cnt), ncand});
        if (out.id != kMiss && best <= texit) break;
        if (sp == 0) break;
        const E e = stack[--sp];
        nx_ = e.x;
        ny_ = e.y;
        tenter = e.a;
        texit = e.b;
    }
}

// ---------------------------------------------------------------- cost model (warp-instructions, from the SASS)
struct Costs {
    int refill_base = 40, refill_ray = 110; // vote/pool bookkeeping; ray load + slab test (executed once per refill event)
    int inner = 50;                         // one inner-node step
    int leaf_setup = 22;
    int plane = 44;                         // one plane test (loop body)
    int plane_pair = 75;                    // two plane tests per trip
    int cand = 72;                          // barycentric part for one survivor
    int record = 10;                        // recording a survivor
    int post = 30;                          // finish test, pop, hit store
    int quantum_overhead = 6;
};

struct Tally {
    double warp_inst = 0, thread_inst = 0;
    void add(double inst, double lanes) {
        warp_inst += inst;
        thread_inst += inst * lanes;
    }
};

struct Lane {
    int ray = -1;
    size_t pos = 0; // next event
    bool busy = false;
};

// Model of trace_persistent_ww_kernel: persistent warps, per-lane quantum = walk to next leaf, test it, pop.
// cand_cap: survivors that can be recorded before an in-loop drain (2 = shipped kernel); pair: two tests per trip.
static Tally sim_ww(const std::vector<RayRec>& rays, const Costs& c, int nwarps, int refill_below, int quanta, int pool_chunk,
                    int cand_cap, bool pair, Tally* parts = nullptr) {
    struct Warp {
        Lane l[32];
        uint32_t pool_next = 0, pool_end = 0;
        bool exhausted = false;
        double clock = 0;
    };
    std::vector<Warp> warps(nwarps);
    uint32_t cursor = 0;
    const uint32_t count = uint32_t(rays.size());
    Tally tot;
    Tally p_inner, p_plane, p_cand, p_refill, p_other;
    using QE = std::pair<double, int>;
    std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
    for (int w = 0; w < nwarps; ++w) pq.push({0.0, w});
    while (!pq.empty()) {
        const int wi = pq.top().second;
        pq.pop();
        Warp& w = warps[wi];
        double cost = 0;
        auto add = [&](Tally& part, double inst, double lanes) {
            tot.add(inst, lanes);
            part.add(inst, lanes);
            cost += inst;
        };
        int nbusy = 0;
        for (auto& l : w.l) nbusy += l.busy;
        if (nbusy < refill_below && !w.exhausted) {
            if (w.pool_next == w.pool_end) {
                const uint32_t b = cursor;
                cursor += pool_chunk;
                if (b >= count) w.exhausted = true;
                else {
                    w.pool_next = b;
                    w.pool_end = std::min(b + uint32_t(pool_chunk), count);
                }
            }
            if (!w.exhausted) {
                int took = 0;
                for (auto& l : w.l) {
                    if (!l.busy && w.pool_next < w.pool_end) {
                        l.ray = int(w.pool_next++);
                        l.pos = 0;
                        l.busy = !rays[l.ray].ev.empty();
                        ++took;
                    }
                }
                add(p_refill, c.refill_base, 32);
                if (took) add(p_refill, c.refill_ray, took);
            }
            nbusy = 0;
            for (auto& l : w.l) nbusy += l.busy;
        }
        if (nbusy == 0) {
            if (w.exhausted) continue; // warp retires
            w.clock += std::max(cost, 1.0);
            pq.push({w.clock, wi});
            continue;
        }
        for (int q = 0; q < quanta; ++q) {
            int nb = 0;
            for (auto& l : w.l) nb += l.busy;
            if (!nb) break;
            add(p_other, c.quantum_overhead, 32);
            // inner loop: iterations = max over lanes, lanes active in iteration k = #lanes with inner > k
            int maxi = 0;
            for (auto& l : w.l)
                if (l.busy) maxi = std::max<int>(maxi, rays[l.ray].ev[l.pos].inner);
            for (int k = 0; k < maxi; ++k) {
                int act = 0;
                for (auto& l : w.l)
                    if (l.busy && rays[l.ray].ev[l.pos].inner > k) ++act;
                add(p_inner, c.inner, act);
            }
            add(p_other, c.leaf_setup, nb);
            // plane-test loop
            int maxc = 0;
            for (auto& l : w.l)
                if (l.busy) maxc = std::max<int>(maxc, rays[l.ray].ev[l.pos].cnt);
            if (pair) {
                for (int k = 0; k < maxc; k += 2) {
                    int act = 0;
                    for (auto& l : w.l)
                        if (l.busy && rays[l.ray].ev[l.pos].cnt > k) ++act;
                    add(p_plane, c.plane_pair, act);
                }
            } else {
                for (int k = 0; k < maxc; ++k) {
                    int act = 0;
                    for (auto& l : w.l)
                        if (l.busy && rays[l.ray].ev[l.pos].cnt > k) ++act;
                    add(p_plane, c.plane, act);
                }
            }
            // survivors: recording (in loop, few lanes), in-loop drains when more than cand_cap, then the post-loop evaluations
            int nrec = 0, maxpost = 0;
            for (auto& l : w.l) {
                if (!l.busy) continue;
                const int nc = rays[l.ray].ev[l.pos].ncand;
                nrec += nc;
                int left = nc;
                while (left > cand_cap) { // a drain of cand_cap survivors, executed by this lane alone inside the loop
                    add(p_cand, c.cand * cand_cap, 1);
                    left -= cand_cap;
                }
                maxpost = std::max(maxpost, left);
            }
            if (nrec) add(p_cand, c.record * std::min(nrec, maxc), std::max(1.0, double(nrec) / std::max(1, std::min(nrec, maxc))));
            for (int k = 0; k < maxpost; ++k) {
                int act = 0;
                for (auto& l : w.l) {
                    if (!l.busy) continue;
                    int nc = rays[l.ray].ev[l.pos].ncand;
                    while (nc > cand_cap) nc -= cand_cap;
                    if (nc > k) ++act;
                }
                add(p_cand, c.cand, act);
            }
            add(p_other, c.post, nb);
            for (auto& l : w.l) {
                if (!l.busy) continue;
                if (++l.pos >= rays[l.ray].ev.size()) l.busy = false;
            }
        }
        w.clock += cost;
        pq.push({w.clock, wi});
    }
    if (parts) {
        parts[0] = p_inner;
        parts[1] = p_plane;
        parts[2] = p_cand;
        parts[3] = p_refill;
        parts[4] = p_other;
    }
    return tot;
}


// Model of the v3 schedule: plane pre-filter (cheap plane loop), survivors deferred into a per-lane list and evaluated
// one per lane per quantum for all lanes that have any (plus a private drain when a list is full or the ray ends),
// optional batched descent: every lane owns a parked second ray that was already walked to its first leaf; new rays are
// fetched and descended only when at least descend_batch lanes have an empty parking slot.
static Tally sim_v3(const std::vector<RayRec>& rays, const Costs& c, int nwarps, int refill_below, int quanta, int pool_chunk,
                    int list_cap, bool defer, int descend_batch, Tally* parts, int min_round = 1) {
    struct L {
        int ray = -1;
        size_t pos = 0;
        bool busy = false;
        int pending = 0;
        int parked = -1; // ray index parked after its descent (descend_batch > 0)
    };
    struct Warp {
        L l[32];
        uint32_t pool_next = 0, pool_end = 0;
        bool exhausted = false;
        double clock = 0;
    };
    std::vector<Warp> warps(nwarps);
    uint32_t cursor = 0;
    const uint32_t count = uint32_t(rays.size());
    Tally tot, p_inner, p_plane, p_cand, p_refill, p_other;
    using QE = std::pair<double, int>;
    std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
    for (int w = 0; w < nwarps; ++w) pq.push({0.0, w});
    while (!pq.empty()) {
        const int wi = pq.top().second;
        pq.pop();
        Warp& w = warps[wi];
        double cost = 0;
        auto add = [&](Tally& part, double inst, double lanes) {
            if (inst <= 0) return;
            tot.add(inst, lanes);
            part.add(inst, lanes);
            cost += inst;
        };
        auto take = [&](int& out) -> bool {
            if (w.pool_next == w.pool_end) {
                if (w.exhausted) return false;
                const uint32_t b = cursor;
                cursor += pool_chunk;
                if (b >= count) {
                    w.exhausted = true;
                    return false;
                }
                w.pool_next = b;
                w.pool_end = std::min(b + uint32_t(pool_chunk), count);
            }
            out = int(w.pool_next++);
            return true;
        };
        int nbusy = 0;
        for (auto& l : w.l) nbusy += l.busy;
        if (descend_batch > 0) {
            // swap parked rays into idle lanes (cheap)
            int swapped = 0;
            for (auto& l : w.l)
                if (!l.busy && l.parked >= 0) {
                    l.ray = l.parked;
                    l.parked = -1;
                    l.pos = 0;
                    l.busy = !rays[l.ray].ev.empty();
                    ++swapped;
                }
            if (swapped) add(p_refill, 30, swapped);
            int empty = 0;
            for (auto& l : w.l) empty += l.parked < 0;
            nbusy = 0;
            for (auto& l : w.l) nbusy += l.busy;
            if (!w.exhausted && (empty >= descend_batch || nbusy < 8)) {
                // batch: fetch + slab test + descend to the first leaf + park
                int took = 0, maxi = 0;
                std::vector<int> d;
                for (auto& l : w.l)
                    if (l.parked < 0) {
                        int r;
                        if (!take(r)) break;
                        l.parked = r;
                        ++took;
                        const int in = rays[r].ev.empty() ? 0 : rays[r].ev[0].inner;
                        d.push_back(in);
                        maxi = std::max(maxi, in);
                    }
                add(p_refill, c.refill_base, 32);
                if (took) {
                    add(p_refill, c.refill_ray + 30, took);
                    for (int k = 0; k < maxi; ++k) {
                        int act = 0;
                        for (int in : d) act += in > k;
                        add(p_inner, c.inner, act);
                    }
                }
            }
            if (nbusy == 0) {
                bool any = false;
                for (auto& l : w.l) any |= l.parked >= 0;
                if (!any && w.exhausted) continue;
                w.clock += std::max(cost, 1.0);
                pq.push({w.clock, wi});
                continue;
            }
        } else {
            if (nbusy < refill_below && !w.exhausted) {
                int took = 0;
                for (auto& l : w.l)
                    if (!l.busy) {
                        int r;
                        if (!take(r)) break;
                        l.ray = r;
                        l.pos = 0;
                        l.pending = 0;
                        l.busy = !rays[r].ev.empty();
                        ++took;
                    }
                add(p_refill, c.refill_base, 32);
                if (took) add(p_refill, c.refill_ray, took);
                nbusy = 0;
                for (auto& l : w.l) nbusy += l.busy;
            }
            if (nbusy == 0) {
                if (w.exhausted) continue;
                w.clock += std::max(cost, 1.0);
                pq.push({w.clock, wi});
                continue;
            }
        }
        for (int q = 0; q < quanta; ++q) {
            int nb = 0;
            for (auto& l : w.l) nb += l.busy;
            if (!nb) break;
            add(p_other, c.quantum_overhead, 32);
            int maxi = 0;
            for (auto& l : w.l)
                if (l.busy) {
                    const int in = (descend_batch > 0 && l.pos == 0) ? 0 : rays[l.ray].ev[l.pos].inner;
                    maxi = std::max(maxi, in);
                }
            for (int k = 0; k < maxi; ++k) {
                int act = 0;
                for (auto& l : w.l)
                    if (l.busy && !(descend_batch > 0 && l.pos == 0) && rays[l.ray].ev[l.pos].inner > k) ++act;
                add(p_inner, c.inner, act);
            }
            add(p_other, c.leaf_setup, nb);
            int maxc = 0;
            for (auto& l : w.l)
                if (l.busy) maxc = std::max<int>(maxc, rays[l.ray].ev[l.pos].cnt);
            for (int k = 0; k < maxc; k += 2) {
                int act = 0;
                for (auto& l : w.l)
                    if (l.busy && rays[l.ray].ev[l.pos].cnt > k) ++act;
                add(p_plane, c.plane_pair, act);
            }
            int nrec = 0;
            for (auto& l : w.l)
                if (l.busy) {
                    nrec += rays[l.ray].ev[l.pos].ncand;
                    l.pending += rays[l.ray].ev[l.pos].ncand;
                }
            if (nrec) add(p_cand, c.record * std::min(nrec, maxc), std::max(1.0, double(nrec) / std::max(1, std::min(nrec, maxc))));
            // survivor evaluation
            if (defer) {
                // lanes whose ray ends here, or whose list is (nearly) full, drain privately first
                for (auto& l : w.l) {
                    if (!l.busy) continue;
                    const bool last = l.pos + 1 >= rays[l.ray].ev.size();
                    if (!last && l.pending > list_cap - 3) {
                        add(p_cand, double(c.cand) * (l.pending - 1), 1);
                        l.pending = 1;
                    }
                }
                int act = 0;
                for (auto& l : w.l) act += l.busy && l.pending > 0;
                if (act >= min_round) {
                    add(p_cand, c.cand, act);
                    for (auto& l : w.l)
                        if (l.busy && l.pending > 0) --l.pending;
                }
                // rays that end: drain the rest together
                int maxd = 0;
                for (auto& l : w.l)
                    if (l.busy && l.pos + 1 >= rays[l.ray].ev.size()) maxd = std::max(maxd, l.pending);
                for (int k = 0; k < maxd; ++k) {
                    int a2 = 0;
                    for (auto& l : w.l)
                        if (l.busy && l.pos + 1 >= rays[l.ray].ev.size() && l.pending > k) ++a2;
                    add(p_cand, c.cand, a2);
                }
                for (auto& l : w.l)
                    if (l.busy && l.pos + 1 >= rays[l.ray].ev.size()) l.pending = 0;
            } else {
                int maxp = 0;
                for (auto& l : w.l)
                    if (l.busy) maxp = std::max(maxp, l.pending);
                for (int k = 0; k < maxp; ++k) {
                    int act = 0;
                    for (auto& l : w.l) act += l.busy && l.pending > k;
                    add(p_cand, c.cand, act);
                }
                for (auto& l : w.l) l.pending = 0;
            }
            add(p_other, c.post, nb);
            for (auto& l : w.l) {
                if (!l.busy) continue;
                if (++l.pos >= rays[l.ray].ev.size()) l.busy = false;
            }
        }
        w.clock += cost;
        pq.push({w.clock, wi});
    }
    if (parts) {
        parts[0] = p_inner;
        parts[1] = p_plane;
        parts[2] = p_cand;
        parts[3] = p_refill;
        parts[4] = p_other;
    }
    return tot;
}


// State-machine schedule: every loop iteration each busy lane advances by ONE unit of whatever it needs next (an inner
// step, a pair of plane tests, one survivor evaluation + leaf epilogue); the warp issues each of the three blocks if at
// least min_lanes[block] lanes want it (a lane whose block is skipped waits), plus a fixed dispatch overhead.
static Tally sim_sm(const std::vector<RayRec>& rays, const Costs& c, int nwarps, int refill_below, int pool_chunk, int dispatch,
                    int period, Tally* parts, int min2 = 1) {
    struct L {
        int ray = -1;
        size_t pos = 0;
        bool busy = false;
        int st = 0;  // 0 inner, 1 plane, 2 cand/epilogue
        int rem = 0; // units left in the current state
    };
    struct Warp {
        L l[32];
        uint32_t pool_next = 0, pool_end = 0;
        bool exhausted = false;
        double clock = 0;
        int it = 0;
    };
    std::vector<Warp> warps(nwarps);
    uint32_t cursor = 0;
    const uint32_t count = uint32_t(rays.size());
    Tally tot, p_inner, p_plane, p_cand, p_refill, p_other;
    using QE = std::pair<double, int>;
    std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
    for (int w = 0; w < nwarps; ++w) pq.push({0.0, w});
    auto enter = [&](L& l) { // set up state for event l.pos
        const Ev& e = rays[l.ray].ev[l.pos];
        if (e.inner > 0) {
            l.st = 0;
            l.rem = e.inner;
        } else if (e.cnt > 0) {
            l.st = 1;
            l.rem = (e.cnt + 1) / 2;
        } else {
            l.st = 2;
            l.rem = 1;
        }
    };
    while (!pq.empty()) {
        const int wi = pq.top().second;
        pq.pop();
        Warp& w = warps[wi];
        double cost = 0;
        auto add = [&](Tally& part, double inst, double lanes) {
            tot.add(inst, lanes);
            part.add(inst, lanes);
            cost += inst;
        };
        int nbusy = 0;
        for (auto& l : w.l) nbusy += l.busy;
        if ((w.it++ % period) == 0 && nbusy < refill_below && !w.exhausted) {
            int took = 0;
            for (auto& l : w.l) {
                if (l.busy) continue;
                if (w.pool_next == w.pool_end) {
                    const uint32_t b = cursor;
                    cursor += pool_chunk;
                    if (b >= count) {
                        w.exhausted = true;
                        break;
                    }
                    w.pool_next = b;
                    w.pool_end = std::min(b + uint32_t(pool_chunk), count);
                }
                l.ray = int(w.pool_next++);
                l.pos = 0;
                l.busy = !rays[l.ray].ev.empty();
                if (l.busy) enter(l);
                ++took;
            }
            add(p_refill, c.refill_base, 32);
            if (took) add(p_refill, c.refill_ray, took);
            nbusy = 0;
            for (auto& l : w.l) nbusy += l.busy;
        }
        if (nbusy == 0) {
            if (w.exhausted) continue;
            w.clock += std::max(cost, 1.0);
            pq.push({w.clock, wi});
            continue;
        }
        int n0 = 0, n1 = 0, n2 = 0;
        for (auto& l : w.l)
            if (l.busy) (l.st == 0 ? n0 : l.st == 1 ? n1 : n2)++;
        add(p_other, dispatch, 32);
        if (n0) add(p_inner, c.inner, n0);
        if (n1) add(p_plane, c.plane_pair, n1);
        const bool run2 = n2 >= min2 || (n2 > 0 && n0 + n1 == 0);
        if (run2) {
            // survivors of the leaf (serial per lane) + epilogue
            int maxs = 0;
            for (auto& l : w.l)
                if (l.busy && l.st == 2) maxs = std::max<int>(maxs, rays[l.ray].ev[l.pos].ncand);
            for (int k = 0; k < maxs; ++k) {
                int a = 0;
                for (auto& l : w.l)
                    if (l.busy && l.st == 2 && rays[l.ray].ev[l.pos].ncand > k) ++a;
                add(p_cand, c.cand, a);
            }
            add(p_other, c.post, n2);
        }
        for (auto& l : w.l) {
            if (!l.busy) continue;
            if (l.st == 2 && !run2) continue;
            if (--l.rem > 0) continue;
            const Ev& e = rays[l.ray].ev[l.pos];
            if (l.st == 0) {
                l.st = e.cnt > 0 ? 1 : 2;
                l.rem = e.cnt > 0 ? (e.cnt + 1) / 2 : 1;
            } else if (l.st == 1) {
                l.st = 2;
                l.rem = 1;
            } else {
                if (++l.pos >= rays[l.ray].ev.size()) l.busy = false;
                else enter(l);
            }
        }
        w.clock += cost;
        pq.push({w.clock, wi});
    }
    if (parts) {
        parts[0] = p_inner;
        parts[1] = p_plane;
        parts[2] = p_cand;
        parts[3] = p_refill;
        parts[4] = p_other;
    }
    return tot;
}


// Speculative walk/test schedule: WALK = a fixed number of warp-wide iterations in which every busy lane does one inner
// step or, at a leaf, appends the leaf to its pending list and pops; TEST = every lane runs the plane pre-filter over
// all triangles of its pending leaves (two per trip); survivors of the whole warp are pooled and evaluated 32 at a time.
static Tally sim_spec(const std::vector<RayRec>& rays, const Costs& c, int nwarps, int refill_below, int pool_chunk, int walk_iters,
                      int pend_cap, int walk_cost, Tally* parts, int test_cap = 1000, int flatten_chunk = 0, int flatten_overhead = 37, int qcap = 1 << 30) {
    struct L {
        int ray = -1;
        size_t wpos = 0; // event being walked
        int wrem = 0;    // inner steps left before its leaf
        bool walking = false;
        int npend = 0;
        size_t tpos = 0; // first pending event
        bool busy = false;
    };
    struct Warp {
        L l[32];
        uint32_t pool_next = 0, pool_end = 0;
        bool exhausted = false;
        double clock = 0;
    };
    std::vector<Warp> warps(nwarps);
    uint32_t cursor = 0;
    const uint32_t count = uint32_t(rays.size());
    Tally tot, p_inner, p_plane, p_cand, p_refill, p_other;
    using QE = std::pair<double, int>;
    std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
    for (int w = 0; w < nwarps; ++w) pq.push({0.0, w});
    while (!pq.empty()) {
        const int wi = pq.top().second;
        pq.pop();
        Warp& w = warps[wi];
        double cost = 0;
        auto add = [&](Tally& part, double inst, double lanes) {
            if (inst <= 0) return;
            tot.add(inst, lanes);
            part.add(inst, lanes);
            cost += inst;
        };
        int nbusy = 0;
        for (auto& l : w.l) nbusy += l.busy;
        if (nbusy < refill_below && !w.exhausted) {
            int took = 0;
            for (auto& l : w.l) {
                if (l.busy) continue;
                if (w.pool_next == w.pool_end) {
                    const uint32_t b = cursor;
                    cursor += pool_chunk;
                    if (b >= count) {
                        w.exhausted = true;
                        break;
                    }
                    w.pool_next = b;
                    w.pool_end = std::min(b + uint32_t(pool_chunk), count);
                }
                l.ray = int(w.pool_next++);
                l.wpos = l.tpos = 0;
                l.npend = 0;
                l.busy = !rays[l.ray].ev.empty();
                l.walking = l.busy;
                if (l.busy) l.wrem = rays[l.ray].ev[0].inner;
                ++took;
            }
            add(p_refill, c.refill_base, 32);
            if (took) add(p_refill, c.refill_ray, took);
            nbusy = 0;
            for (auto& l : w.l) nbusy += l.busy;
        }
        if (nbusy == 0) {
            if (w.exhausted) continue;
            w.clock += std::max(cost, 1.0);
            pq.push({w.clock, wi});
            continue;
        }
        // WALK
        int qfill = 0;
        for (int it = 0; it < walk_iters; ++it) {
            int act = 0;
            for (auto& l : w.l) {
                if (!l.busy || !l.walking || l.npend >= pend_cap) continue;
                {
                    int pt = 0;
                    for (int k = 0; k < l.npend; ++k) pt += rays[l.ray].ev[l.tpos + k].cnt;
                    if (pt >= test_cap) continue;
                }
                ++act;
                if (l.wrem > 0) --l.wrem;
                else { // at the leaf: append + pop
                    if (flatten_chunk > 0) {
                        const int ch = (rays[l.ray].ev[l.wpos].cnt + flatten_chunk - 1) / flatten_chunk;
                        if (qfill + ch > qcap) continue; // blocked until the next cycle
                        qfill += ch;
                    }
                    ++l.npend;
                    if (++l.wpos >= rays[l.ray].ev.size()) l.walking = false;
                    else l.wrem = rays[l.ray].ev[l.wpos].inner;
                }
            }
            if (!act) break;
            add(p_inner, walk_cost, act);
        }
        // TEST: per lane all triangles of its pending leaves
        int maxt = 0, nsurv = 0, nt = 0;
        for (auto& l : w.l) {
            if (!l.busy) continue;
            int t = 0;
            for (int k = 0; k < l.npend; ++k) {
                const Ev& e = rays[l.ray].ev[l.tpos + k];
                t += (e.cnt + 1) / 2;
                nsurv += e.ncand;
            }
            maxt = std::max(maxt, t);
            nt += l.npend > 0;
        }
        add(p_other, c.leaf_setup, std::max(nt, 1));
        if (flatten_chunk > 0) {
            // chunks of <= flatten_chunk tests pooled over the warp, one chunk per lane per round
            std::vector<int> chunks;
            for (auto& l : w.l) {
                if (!l.busy) continue;
                for (int j = 0; j < l.npend; ++j) {
                    int cnt = rays[l.ray].ev[l.tpos + j].cnt;
                    while (cnt > 0) {
                        chunks.push_back(std::min(cnt, flatten_chunk));
                        cnt -= flatten_chunk;
                    }
                }
            }
            for (size_t b = 0; b < chunks.size(); b += 32) {
                const size_t e = std::min(chunks.size(), b + 32);
                add(p_plane, flatten_overhead, double(e - b));
                int mx = 0;
                for (size_t i = b; i < e; ++i) mx = std::max(mx, (chunks[i] + 1) / 2);
                for (int k = 0; k < mx; ++k) {
                    int act = 0;
                    for (size_t i = b; i < e; ++i) act += (chunks[i] + 1) / 2 > k;
                    add(p_plane, c.plane_pair, act);
                }
            }
        } else
        for (int k = 0; k < maxt; ++k) {
            int act = 0;
            for (auto& l : w.l) {
                if (!l.busy) continue;
                int t = 0;
                for (int j = 0; j < l.npend; ++j) t += (rays[l.ray].ev[l.tpos + j].cnt + 1) / 2;
                if (t > k) ++act;
            }
            add(p_plane, c.plane_pair + 4, act);
        }
        if (nsurv) add(p_cand, c.record, std::min(nsurv, 8));
        for (int b = 0; b < nsurv; b += 32) add(p_cand, c.cand + 20, std::min(32, nsurv - b));
        add(p_other, c.post, nbusy);
        for (auto& l : w.l) {
            if (!l.busy) continue;
            l.tpos += l.npend;
            l.npend = 0;
            if (l.tpos >= rays[l.ray].ev.size()) l.busy = false;
        }
        w.clock += cost;
        pq.push({w.clock, wi});
    }
    if (parts) {
        parts[0] = p_inner;
        parts[1] = p_plane;
        parts[2] = p_cand;
        parts[3] = p_refill;
        parts[4] = p_other;
    }
    return tot;
}

static void report(const char* name, const Tally& t, size_t nrays, const Tally* parts) {
    std::printf("%-46s warp-inst/ray %7.1f  lanes %5.2f", name, t.warp_inst / nrays, t.thread_inst / t.warp_inst);
    if (parts) {
        const char* nm[5] = {"inner", "plane", "cand", "refill", "other"};
        for (int i = 0; i < 5; ++i) std::printf("  %s %4.1f%%@%4.1f", nm[i], 100 * parts[i].warp_inst / t.warp_inst, parts[i].thread_inst / std::max(1.0, parts[i].warp_inst));
    }
    std::printf("\n");
}

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "/tmp/sim/mesh1m.bin";
    const int rows = argc > 2 ? std::atoi(argv[2]) : 24;
    FILE* f = std::fopen(path, "rb");
    if (!f) return 1;
    uint32_t n = 0;
    if (std::fread(&n, 4, 1, f) != 1) return 1;
    std::vector<float> V(size_t(n) * 9), N(size_t(n) * 9), D(size_t(n) * 4, 0.7f);
    if (std::fread(V.data(), 4, V.size(), f) != V.size() || std::fread(N.data(), 4, N.size(), f) != N.size()) return 1;
    std::fclose(f);
    Scene sc;
    precompute_triangles(V.data(), N.data(), D.data(), n, sc.tris);
    build_kdtree(sc.tris, sc.tree, 0);
    std::printf("triangles %u  pair nodes %zu  refs %zu  height %llu  build %.0f ms\n", n, sc.tree.pair_nodes.size(),
                sc.tree.pair_leaf_refs.size(), (unsigned long long)sc.tree.height, sc.tree.build_ms);
    {
        size_t leaves = 0, voids = 0, refs = 0, hist[12] = {0}; uint32_t maxc = 0;
        for (uint64_t v : sc.tree.pair_nodes) {
            const uint32_t y = uint32_t(v >> 32);
            if ((y & 3u) == 3u) {
                const uint32_t c = y >> 2;
                if (c == 0) ++voids;
                else {
                    ++leaves;
                    refs += c;
                    hist[std::min<uint32_t>(c, 11)]++; maxc = std::max(maxc, c);
                }
            }
        }
        std::printf("leaves %zu voids %zu refs/leaf %.2f max %u  hist:", leaves, voids, double(refs) / leaves, maxc);
        for (int i = 1; i < 12; ++i) std::printf(" %d:%.1f%%", i, 100.0 * hist[i] / leaves);
        std::printf("\n");
    }

    // camera of scenes.cubesphere(): trafo (row-major 3x3 + position), hfov
    const float R[9] = {0.9438583850860596f, -0.07586748898029327f, 0.3215206563472748f, 0.0f, 0.9732714891433716f, 0.2296576052904129f,
                        -0.33035042881965637f, -0.21676425635814667f, 0.9186304211616516f};
    const float P[3] = {1.4881685972213745f, 1.0629775524139404f, 4.251910209655762f};
    const float delta = std::tan(0.4287780225276947f);
    const int W = 1920, H = 1920;
    std::vector<Ray> wave;
    const int y0 = H / 2 - rows / 2 - 200; // a band above the centre: silhouette + interior
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
        std::vector<RayRec> recs(wave.size());
#pragma omp parallel for schedule(dynamic, 256)
        for (long i = 0; i < long(wave.size()); ++i) trace(sc, wave[i], recs[i]);
        size_t hits = 0;
        double inner = 0, leaves = 0, tests = 0, cands = 0, first_inner = 0, nonempty = 0;
        for (auto& r : recs) {
            hits += r.id != kMiss;
            if (!r.ev.empty()) {
                first_inner += r.ev[0].inner;
                nonempty += 1;
            }
            for (auto& e : r.ev) {
                inner += e.inner;
                leaves += 1;
                tests += e.cnt;
                cands += e.ncand;
            }
        }
        const double nr = double(recs.size());
        std::printf("\n== depth %d wave: %zu rays, hit %.1f%%; per ray: inner %.1f (initial descent %.1f) leaves %.1f tests %.1f survivors %.2f\n", depth,
                    recs.size(), 100.0 * hits / nr, inner / nr, first_inner / std::max(1.0, nonempty), leaves / nr, tests / nr, cands / nr);
        {
            Costs c;
            Tally parts[5];
            const int nwarps = 64;
            Tally t = sim_ww(recs, c, nwarps, 28, 2, 32, 2, true, parts);
            report("shipped: ww, refill<28, q2, cap2, pairs", t, recs.size(), parts);
            t = sim_ww(recs, c, nwarps, 28, 2, 32, 8, true, parts);
            report("  + survivor list of 8 (no in-loop drain)", t, recs.size(), parts);
            t = sim_ww(recs, c, nwarps, 16, 2, 32, 8, true, parts);
            report("  + refill<16", t, recs.size(), parts);
            Costs c2 = c;
            c2.inner = 36;
            c2.plane_pair = 46;
            c2.plane = 26;
            t = sim_ww(recs, c2, nwarps, 28, 2, 32, 8, true, parts);
            report("  cap8 + cheaper inner(36)/plane pair(46)", t, recs.size(), parts);
            Costs c3 = c;
            c3.inner = 36;
            c3.plane_pair = 40;
            c3.cand = 100; // exact plane + barycentric part
            c3.record = 6;
            t = sim_v3(recs, c3, nwarps, 28, 2, 32, 8, false, 0, parts);
            report("v3 costs, survivors after each leaf", t, recs.size(), parts);
            t = sim_v3(recs, c3, nwarps, 28, 2, 32, 8, true, 0, parts);
            report("v3 + deferred survivors (1/lane/quantum)", t, recs.size(), parts);
            for (int mr : {4, 8, 12}) {
                t = sim_v3(recs, c3, nwarps, 28, 2, 32, 8, true, 0, parts, mr);
                char nm[64];
                std::snprintf(nm, sizeof nm, "v3 + deferred, round only if >= %d lanes", mr);
                report(nm, t, recs.size(), parts);
            }
            for (int m2 : {1, 4, 8, 12, 16}) {
                t = sim_sm(recs, c3, nwarps, 28, 32, 12, 4, parts, m2);
                char nm[64];
                std::snprintf(nm, sizeof nm, "state machine, dispatch 12, cand block if >= %d", m2);
                report(nm, t, recs.size(), parts);
            }
            for (int wi_ : {4, 8, 12, 16}) {
                for (int pc : {2, 4}) {
                    t = sim_spec(recs, c3, nwarps, 24, 32, wi_, pc, 50, parts);
                    char nm[64];
                    std::snprintf(nm, sizeof nm, "speculative walk %d iters, pend cap %d", wi_, pc);
                    report(nm, t, recs.size(), parts);
                }
            }
            for (int tc : {6, 10, 16}) {
                t = sim_spec(recs, c3, nwarps, 24, 32, 12, 4, 50, parts, tc);
                char nm[64];
                std::snprintf(nm, sizeof nm, "spec walk 12, cap 4 leaves / %d tests", tc);
                report(nm, t, recs.size(), parts);
            }
            for (int fc : {2, 4, 8}) {
                t = sim_spec(recs, c3, nwarps, 24, 32, 12, 4, 50, parts, 1000, fc);
                char nm[64];
                std::snprintf(nm, sizeof nm, "spec walk 12, cap 4, pooled chunks of %d", fc);
                report(nm, t, recs.size(), parts);
            }
            for (int qc : {48, 64, 96, 128}) {
                t = sim_spec(recs, c3, nwarps, 24, 32, 12, 8, 50, parts, 1000, 4, 20, qc);
                char nm[64];
                std::snprintf(nm, sizeof nm, "spec walk<=12, pooled chunks 4, queue %d, ovh 20", qc);
                report(nm, t, recs.size(), parts);
            }
            for (int db : {16}) {
                t = sim_v3(recs, c3, nwarps, 28, 2, 32, 8, true, db, parts);
                char nm[64];
                std::snprintf(nm, sizeof nm, "v3 + deferred + batched descent >= %d", db);
                report(nm, t, recs.size(), parts);
            }
        }
        // next wave, laid out as shade_bounce_kernel does: CTA of 256 threads = 8 warps, per warp block k-major
        std::vector<Ray> next;
        for (size_t base = 0; base < wave.size(); base += 32) {
            std::vector<size_t> hl;
            for (size_t i = base; i < std::min(wave.size(), base + 32); ++i)
                if (recs[i].id != kMiss) hl.push_back(i);
            const size_t nh = hl.size();
            const size_t cbase = next.size();
            next.resize(cbase + nh * m);
            for (size_t rank = 0; rank < nh; ++rank) {
                const size_t i = hl[rank];
                const RayRec& h = recs[i];
                const Ray& ry = wave[i];
                const float* q = &sc.tris.isect[size_t(h.id) * 16];
                float nrm[3] = {q[3], q[4], q[5]};
                float p[3];
                for (int c = 0; c < 3; ++c) p[c] = ry.o[c] + h.r * ry.d[c];
                // face the normal against the incoming ray like a shading normal of the outward-oriented mesh would
                // (the mesh is oriented outward; keep as is)
                float p2[3];
                for (int c = 0; c < 3; ++c) p2[c] = p[c] + 0.0001f * nrm[c];
                // orthonormal frame
                float a[3] = {std::fabs(nrm[0]) < 0.9f ? 1.f : 0.f, std::fabs(nrm[0]) < 0.9f ? 0.f : 1.f, 0.f};
                float t1[3] = {nrm[1] * a[2] - nrm[2] * a[1], nrm[2] * a[0] - nrm[0] * a[2], nrm[0] * a[1] - nrm[1] * a[0]};
                float l1 = std::sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
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
