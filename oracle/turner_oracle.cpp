// TEST INFRASTRUCTURE ONLY -- CPU oracle for the path-tracing hot path.
//
// A from-scratch restatement (plain C++14, no dependencies) of what the
// reference computes on this path, used as the checker in tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline leg. It is never linked
// into, imported by, or executed from the product path (turner_b200/).
//
// Parity status: PINNED. tests/test_oracle_vs_reference.py checks every
// function here bit-for-bit against oracle/_ref (the reference's own kdtree.cpp /
// pathtracer.cpp / raycaster.cpp compiled from /root/reference, see
// oracle/build_ref.sh) and against the reference's known-answer tests.
//
// Build: g++ -std=c++14 -O2 -ffp-contract=off (no -ffast-math, no -march): the
// reference's CMakeLists.txt:3-7 sets no -march, so its fp32 has no FMA.
//
// Follows (paths relative to /root/reference):
//   src/geometry.h:280-304,786-805   dot / cross(in double) / normalize conventions
//   lib/triangle.h:22-71             precomputed triangle fields, normal interpolation, bbox
//   lib/intersection.h:40-128        ray/plane, ray/triangle, ray/box
//   lib/clipping.h:120-235           polygon/thick-plane clipping, clipped triangle box
//   lib/kdtree.cpp:128-174,178-408   SAH build (Wald-Havran Alg. 4), classification
//   lib/kdtree.cpp:420-467, kdtree.h:62-154,197-218   flatten, node encoding, height
//   lib/kdtree.cpp:503-607           traversal without early exit + leaf runs
//   lib/xorshift.h:27-62, lib/sampling.h:12-32   RNG + uniform hemisphere
//   pathtracer.cpp:14-102, raycaster.cpp:7-24    integrators
//   lib/types.h:92-123               camera
//   main.cpp:181-236                 render loop
//   lib/effects.h:15-48, lib/raster.h:79-100     tone map + P3 writer
// and assimp@a5a5343 (not vendored) for FromToMatrix / matrix*vector / colour ops
// (restated from its published .inl files; call sites pathtracer.cpp:55-101).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace orc {

constexpr float kEps = 0.00001f;                                // lib/types.h:13
constexpr float kFltMax = std::numeric_limits<float>::max();    // lib/types.h:14
constexpr uint32_t kMissId = 1u << 30;                          // kdtree.h:66,158
constexpr uint32_t kInvalidTri = 0xFFFFFFFFu >> 2;              // kdtree.h:67

struct V3 {
    float x, y, z;
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 smul(float s, V3 v) { return {s * v.x, s * v.y, s * v.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } // geometry.h:280-282
inline float length(V3 v) { return std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }
inline V3 normalize(V3 v) { return smul(1 / length(v), v); } // geometry.h:243-246,302-304
inline V3 cross_d(V3 a, V3 b) {                              // geometry.h:288-300 (double, rounded once)
    double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
    return {static_cast<float>((ay * bz) - (az * by)), static_cast<float>((az * bx) - (ax * bz)),
            static_cast<float>((ax * by) - (ay * bx))};
}

struct C4 {
    float r, g, b, a;
};
inline C4 cadd(C4 x, C4 y) { return {x.r + y.r, x.g + y.g, x.b + y.b, x.a + y.a}; }
inline C4 cmul(C4 x, C4 y) { return {x.r * y.r, x.g * y.g, x.b * y.b, x.a * y.a}; }
inline C4 cscale(float f, C4 v) { return {f * v.r, f * v.g, f * v.b, f * v.a}; }

struct Box {
    V3 lo, hi;
};
inline float fmin2(float a, float b) { return std::fmin(a, b); }
inline float fmax2(float a, float b) { return std::fmax(a, b); }
// std::min / std::max semantics of geometry.h's min()/max() on points
inline float smin(float a, float b) { return b < a ? b : a; }
inline float smax(float a, float b) { return a < b ? b : a; }
inline Box box_from(V3 p1, V3 p2) { // Bbox3(p1,p2), geometry.h:998-999
    return {{smin(p1.x, p2.x), smin(p1.y, p2.y), smin(p1.z, p2.z)}, {smax(p1.x, p2.x), smax(p1.y, p2.y), smax(p1.z, p2.z)}};
}
inline Box box_default() { // Bbox3(), geometry.h:991-996: "everything"
    float lo = std::numeric_limits<float>::lowest();
    return {{lo, lo, lo}, {kFltMax, kFltMax, kFltMax}};
}
inline float surface_area(const Box& b) { // geometry.h:1044-1047
    V3 d = sub(b.hi, b.lo);
    return 2 * (d.x * d.y + d.x * d.z + d.y * d.z);
}
inline void box_split(const Box& b, int ax, float pos, Box& l, Box& r) { // geometry.h:1064-1076
    V3 lmax = b.hi;
    lmax[ax] = pos;
    V3 rmin = b.lo;
    rmin[ax] = pos;
    l = box_from(b.lo, lmax);
    r = box_from(rmin, b.hi);
}
inline Box box_union(const Box& a, const Box& b) {
    return box_from({smin(a.lo.x, b.lo.x), smin(a.lo.y, b.lo.y), smin(a.lo.z, b.lo.z)},
                    {smax(a.hi.x, b.hi.x), smax(a.hi.y, b.hi.y), smax(a.hi.z, b.hi.z)});
}

// 192-byte AoS record like the reference's Triangle (lib/triangle.h:95-114)
struct Tri {
    V3 p[3];
    V3 n[3];
    C4 ambient, diffuse, emissive, reflective;
    float reflectivity;
    V3 u, v, fn;
    float uv, vv, uu, denom;
};
static_assert(sizeof(Tri) == 192, "reference Triangle is 192 bytes");

inline Tri make_tri(const float* vp, const float* np, const float* dc, const float* mirror_rgba = nullptr, float reflectivity = 0.f) {
    Tri t;
    std::memset(&t, 0, sizeof(t));
    for (int k = 0; k < 3; ++k) {
        t.p[k] = {vp[3 * k], vp[3 * k + 1], vp[3 * k + 2]};
        t.n[k] = {np[3 * k], np[3 * k + 1], np[3 * k + 2]};
    }
    t.diffuse = {dc[0], dc[1], dc[2], dc[3]};
    t.emissive = t.diffuse; // main.cpp:43 reads DIFFUSE into emissive
    if (mirror_rgba) t.reflective = {mirror_rgba[0], mirror_rgba[1], mirror_rgba[2], mirror_rgba[3]}; // main.cpp:44
    t.reflectivity = reflectivity;                                                                    // main.cpp:46-47
    t.u = sub(t.p[1], t.p[0]);
    t.v = sub(t.p[2], t.p[0]);
    t.fn = normalize(cross_d(t.u, t.v));
    t.uv = dot(t.u, t.v);
    t.vv = dot(t.v, t.v);
    t.uu = dot(t.u, t.u);
    t.denom = t.uv * t.uv - t.uu * t.vv;
    return t;
}
inline float fmin3(float x, float y, float z) { return std::fmin(x, std::min(y, z)); } // types.h:66-68
inline float fmax3(float x, float y, float z) { return std::fmax(x, std::max(y, z)); }
inline Box tri_bbox(const Tri& t) { // triangle.h:61-71
    V3 mn = {fmin3(t.p[0].x, t.p[1].x, t.p[2].x), fmin3(t.p[0].y, t.p[1].y, t.p[2].y), fmin3(t.p[0].z, t.p[1].z, t.p[2].z)};
    V3 mx = {fmax3(t.p[0].x, t.p[1].x, t.p[2].x), fmax3(t.p[0].y, t.p[1].y, t.p[2].y), fmax3(t.p[0].z, t.p[1].z, t.p[2].z)};
    return box_from(mn, mx);
}

struct Ray {
    V3 o, d;
};

// ---------------------------------------------------------------- intersection
// lib/intersection.h:40-49 + 63-89. Returns accept/reject; r,s,t as the reference leaves them.
inline bool ray_triangle(const Ray& ray, const Tri& tri, float& r, float& s, float& t) {
    float denom = dot(tri.fn, ray.d);
    if (denom == 0.f) {
        r = std::numeric_limits<float>::lowest();
    } else {
        float nom = dot(tri.fn, sub(tri.p[0], ray.o));
        r = nom / denom;
    }
    if (r < 0) return false;
    V3 P = add(ray.o, smul(r, ray.d));
    V3 w = sub(P, tri.p[0]);
    float wv = dot(w, tri.v);
    float wu = dot(w, tri.u);
    s = (tri.uv * wv - tri.vv * wu) / tri.denom;
    if (s < 0) return false;
    t = (tri.uv * wu - tri.uu * wv) / tri.denom;
    if (t < 0 || 1 < s + t) return false;
    return true;
}

// lib/intersection.h:105-128
inline bool ray_box(const Ray& ray, const Box& box, float& tmin, float& tmax) {
    V3 di = {1 / ray.d.x, 1 / ray.d.y, 1 / ray.d.z};
    float tx1 = (box.lo.x - ray.o.x) * di.x;
    float tx2 = (box.hi.x - ray.o.x) * di.x;
    tmin = fmin2(tx1, tx2);
    tmax = fmax2(tx1, tx2);
    float ty1 = (box.lo.y - ray.o.y) * di.y;
    float ty2 = (box.hi.y - ray.o.y) * di.y;
    tmin = fmax2(tmin, fmin2(ty1, ty2));
    tmax = fmin2(tmax, fmax2(ty1, ty2));
    float tz1 = (box.lo.z - ray.o.z) * di.z;
    float tz2 = (box.hi.z - ray.o.z) * di.z;
    tmin = fmax2(tmin, fmin2(tz1, tz2));
    tmax = fmin2(tmax, fmax2(tz1, tz2));
    return !(tmax < tmin);
}

// -------------------------------------------------------------------- clipping
// lib/clipping.h:120-187: Sutherland-Hodgman against the thick plane n.x = d (EPS slab).
inline int side_of(V3 p, V3 n, float d) {
    float dist = dot(n, p) - d;
    if (dist > kEps) return 1;   // in front
    if (dist < -kEps) return -1; // behind
    return 0;                    // on plane
}
inline void clip_poly(const std::vector<V3>& poly, V3 n, float d, std::vector<V3>& out) {
    out.clear();
    V3 a = poly.back();
    int sa = side_of(a, n, d);
    for (const V3& b : poly) {
        int sb = side_of(b, n, d);
        if (sb == 1) {
            if (sa == -1) {
                V3 ab = sub(b, a);
                float t = (d - dot(n, a)) / dot(n, ab); // intersection.h:18-23
                out.push_back(add(a, smul(t, sub(b, a))));
            }
            out.push_back(b);
        } else if (sb == -1) {
            if (sa == 1) {
                V3 ab = sub(b, a);
                float t = (d - dot(n, a)) / dot(n, ab);
                out.push_back(add(a, smul(t, sub(b, a))));
            }
        } else {
            out.push_back(b);
        }
        a = b;
        sa = sb;
    }
}
// lib/clipping.h:199-235. NB: "fewer than 2 points left" yields the *default* box
// (everything), which is not empty() -- the reference's quirk, kept.
inline Box clipped_tri_box(const Tri& tri, const Box& box) {
    std::vector<V3> pts(tri.p, tri.p + 3), tmp;
    for (int ax = 0; ax < 3; ++ax) {
        for (int side = 0; side < 2; ++side) {
            V3 n = {0, 0, 0};
            n[ax] = side == 0 ? 1.f : -1.f;
            float dist = side == 0 ? box.lo[ax] : -box.hi[ax];
            clip_poly(pts, n, dist, tmp);
            pts.swap(tmp);
            if (pts.size() < 2) return box_default();
        }
    }
    float lowest = std::numeric_limits<float>::lowest();
    V3 mn = {kFltMax, kFltMax, kFltMax}, mx = {lowest, lowest, lowest};
    for (int ax = 0; ax < 3; ++ax)
        for (const V3& p : pts) {
            if (p[ax] < mn[ax]) mn[ax] = p[ax];
            if (mx[ax] < p[ax]) mx[ax] = p[ax];
        }
    return box_from(mn, mx);
}
inline bool box_empty(const Box& b) { return b.hi.x <= b.lo.x && b.hi.y <= b.lo.y && b.hi.z <= b.lo.z; }
inline bool box_planar(const Box& b, int ax) { return std::abs(b.hi[ax] - b.lo[ax]) < kEps; }

// -------------------------------------------------------------------- kd build
struct BNode {
    int axis = -1; // -1: leaf
    float split = 0;
    std::unique_ptr<BNode> l, r;
    std::vector<uint32_t> ids;
};

struct Builder {
    const std::vector<Tri>& tris;
    explicit Builder(const std::vector<Tri>& t) : tris(t) {}

    static float lambda(size_t nl, size_t nr) { return (nl == 0 || nr == 0) ? 0.8f : 1.f; } // kdtree.cpp:182-187
    static float cost(float lr, float rr, size_t nl, size_t nr) {                             // kdtree.cpp:197-203
        return lambda(nl, nr) * (15 + 20 * (lr * nl + rr * nr));
    }
    // kdtree.cpp:218-241; returns cost, sets left=true when planar triangles go left
    static float sah(int ax, float pos, const Box& box, size_t nl, size_t nr, size_t np, bool& left) {
        Box lb, rb;
        box_split(box, ax, pos, lb, rb);
        float area = surface_area(box);
        float lr = surface_area(lb) / area;
        float rr = surface_area(rb) / area;
        float cl = cost(lr, rr, nl + np, nr);
        float cr = cost(lr, rr, nl, np + nr);
        if (cl < cr) {
            left = true;
            return cl;
        }
        left = false;
        return cr;
    }

    struct Event {
        uint32_t id;
        float point;
        float aux;
        int type; // 0 ending, 1 planar, 2 starting (kdtree.cpp:254-256)
    };

    // kdtree.cpp:245-408
    float find_plane(const std::vector<uint32_t>& ids, const Box& box, int& best_ax, float& best_pos,
                     std::vector<uint32_t>& lt, std::vector<uint32_t>& rt) const {
        std::vector<Event> ev[3];
        size_t num = 0;
        for (uint32_t id : ids) {
            Box cb = clipped_tri_box(tris[id], box);
            if (box_empty(cb)) continue;
            num += 1;
            for (int ax = 0; ax < 3; ++ax) {
                if (box_planar(cb, ax)) {
                    ev[ax].push_back({id, cb.lo[ax], cb.lo[ax], 1});
                } else {
                    ev[ax].push_back({id, cb.lo[ax], cb.hi[ax], 2});
                    ev[ax].push_back({id, cb.hi[ax], cb.lo[ax], 0});
                }
            }
        }
        best_ax = 0;
        best_pos = 0;
        lt.clear();
        rt.clear();
        if (num == 0) return kFltMax;

        float min_cost = kFltMax;
        bool best_left = true;
        for (int ax = 0; ax < 3; ++ax) {
            auto& e = ev[ax];
            std::sort(e.begin(), e.end(), [](const Event& a, const Event& b) {
                return a.point < b.point || (a.point == b.point && a.type < b.type);
            });
            size_t nl = 0, np = 0, nr = num;
            for (size_t i = 0; i < e.size();) {
                float p = e[i].point;
                int s = 0, en = 0, pl = 0;
                while (i < e.size() && e[i].point == p && e[i].type == 0) { en += 1; i += 1; }
                while (i < e.size() && e[i].point == p && e[i].type == 1) { pl += 1; i += 1; }
                while (i < e.size() && e[i].point == p && e[i].type == 2) { s += 1; i += 1; }
                np = pl;
                nr -= pl + en;
                bool left;
                float c = sah(ax, p, box, nl, nr, np, left);
                if (c < min_cost) {
                    min_cost = c;
                    best_ax = ax;
                    best_pos = p;
                    best_left = left;
                }
                nl += s + pl;
                np = 0;
                // NaN points never compare equal: the three while-loops above would not
                // advance. The reference would spin forever; inputs must be finite.
                if (s + en + pl == 0) i += 1;
            }
        }
        for (const Event& e : ev[best_ax]) {
            if (e.point < best_pos) {
                if (e.type == 0 || e.type == 1) {
                    lt.push_back(e.id);
                } else if (best_pos < e.aux) {
                    lt.push_back(e.id);
                    rt.push_back(e.id);
                }
            } else if (e.point == best_pos) {
                if (e.type == 0) {
                    lt.push_back(e.id);
                } else if (e.type == 1) {
                    (best_left ? lt : rt).push_back(e.id);
                } else {
                    rt.push_back(e.id);
                }
            } else if (e.type == 2 || e.type == 1) {
                rt.push_back(e.id);
            }
        }
        return min_cost;
    }

    // kdtree.cpp:128-174
    std::unique_ptr<BNode> build(std::vector<uint32_t> ids, const Box& box) const {
        if (ids.empty()) return nullptr;
        auto leaf = [&]() {
            std::unique_ptr<BNode> n(new BNode);
            n->ids = std::move(ids);
            return n;
        };
        if (ids.size() <= 3) return leaf();
        if (surface_area(box) == 0) return leaf();
        int ax;
        float pos;
        std::vector<uint32_t> lt, rt;
        float min_cost = find_plane(ids, box, ax, pos, lt, rt);
        if (20 * ids.size() * lambda(lt.size(), rt.size()) < min_cost) return leaf();
        Box lb, rb;
        box_split(box, ax, pos, lb, rb);
        ids.clear();
        ids.shrink_to_fit();
        auto l = build(std::move(lt), lb);
        auto r = build(std::move(rt), rb);
        if (!l) return r;
        if (!r) return l;
        std::unique_ptr<BNode> n(new BNode);
        n->axis = ax;
        n->split = pos;
        n->l = std::move(l);
        n->r = std::move(r);
        return n;
    }
};

inline uint32_t fbits(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline float bitsf(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
// node encoding, kdtree.h:62-154
inline uint64_t enc_inner(int axis, float split, uint32_t right) {
    return (static_cast<uint64_t>(fbits(split)) << 32) | static_cast<uint32_t>((right << 2) | static_cast<uint32_t>(axis));
}
inline uint64_t enc_leaf(uint32_t a, uint32_t b) {
    return (static_cast<uint64_t>(a) << 32) | static_cast<uint64_t>(static_cast<uint32_t>((b << 2) | 3u));
}
inline bool n_is_leaf(uint64_t n) { return (n & 3) == 3; }
inline int n_axis(uint64_t n) { return static_cast<int>(n & 3); }
inline float n_split(uint64_t n) { return bitsf(static_cast<uint32_t>(n >> 32)); }
inline uint32_t n_right(uint64_t n) { return static_cast<uint32_t>(n) >> 2; }
inline uint32_t n_first(uint64_t n) { return static_cast<uint32_t>(n >> 32); }
inline bool n_has_second(uint64_t n) { return static_cast<uint32_t>(n & 0xFFFFFFFFu) != 0xFFFFFFFFu; }
inline uint32_t n_second(uint64_t n) { return static_cast<uint32_t>(n & 0xFFFFFFFFu) >> 2; }

// kdtree.cpp:420-467: DFS order, left child = next node, leaf runs of id pairs,
// run terminated by a one-id leaf or an all-zero inner "sentinel".
inline std::vector<uint64_t> flatten(const BNode* root) {
    std::vector<uint64_t> nodes;
    struct Item {
        const BNode* n;
        uint32_t parent;
    };
    std::vector<Item> stack;
    stack.push_back({root, kInvalidTri});
    while (!stack.empty()) {
        Item it = stack.back();
        stack.pop_back();
        uint32_t idx = static_cast<uint32_t>(nodes.size());
        if (it.parent != kInvalidTri) {
            uint64_t p = nodes[it.parent];
            nodes[it.parent] = (p & 0xFFFFFFFF00000000ull) | static_cast<uint32_t>((idx << 2) | static_cast<uint32_t>(n_axis(p)));
        }
        if (it.n->axis >= 0) {
            nodes.push_back(enc_inner(it.n->axis, it.n->split, kInvalidTri));
            stack.push_back({it.n->r.get(), idx});
            stack.push_back({it.n->l.get(), kInvalidTri});
        } else {
            const auto& ids = it.n->ids;
            size_t i = 1;
            for (; i < ids.size(); i += 2) nodes.push_back(enc_leaf(ids[i - 1], ids[i]));
            if (i - 1 < ids.size()) {
                nodes.push_back(enc_leaf(ids[i - 1], kInvalidTri));
            } else {
                nodes.push_back(0); // sentinel = FlatNode(X, 0, 0)
            }
        }
    }
    return nodes;
}

struct Scene {
    std::vector<Tri> tris;
    Box box;
    std::vector<uint64_t> nodes;
    double build_ms = 0;

    size_t height() const { // kdtree.h:197-218
        std::vector<std::pair<uint32_t, uint32_t>> st;
        st.emplace_back(0u, 0u);
        size_t h = 0;
        while (!st.empty()) {
            auto cur = st.back();
            st.pop_back();
            if (cur.second > h) h = cur.second;
            uint64_t n = nodes[cur.first];
            if (!n_is_leaf(n)) {
                st.emplace_back(cur.first + 1, cur.second + 1);
                st.emplace_back(n_right(n), cur.second + 1);
            }
        }
        return h;
    }
};

// ------------------------------------------------------------------- traversal
struct Counters {
    uint64_t inner = 0, leaf_nodes = 0, tri_tests = 0, queries = 0;
    // of tri_tests: how many re-test a triangle this ray already tested among its last 4 / 8 / all distinct ones
    // (triangles straddling several leaves); informs the device kernel's mailbox size
    uint64_t repeat4 = 0, repeat8 = 0, repeat_any = 0;
    std::vector<uint32_t> seen;
};

struct Traverser {
    const Scene& sc;
    struct Item {
        uint32_t node;
        float tenter, texit;
    };
    std::vector<Item> stack;
    explicit Traverser(const Scene& s) : sc(s) { stack.reserve(128); }

    // leaf run helper, kdtree.cpp:580-607
    uint32_t leaf_run(uint32_t node, const Ray& ray, float& mr, float& ms, float& mt, Counters* c) const {
        mr = kFltMax;
        uint32_t res = kMissId;
        auto test = [&](uint32_t id) {
            float r, s, t;
            if (c) {
                c->tri_tests += 1;
                size_t n = c->seen.size();
                for (size_t k = 0; k < n; ++k)
                    if (c->seen[n - 1 - k] == id) {
                        c->repeat_any += 1;
                        if (k < 8) c->repeat8 += 1;
                        if (k < 4) c->repeat4 += 1;
                        break;
                    }
                c->seen.push_back(id);
            }
            bool hit = ray_triangle(ray, sc.tris[id], r, s, t);
            if (hit && r < mr) {
                mr = r;
                ms = s;
                mt = t;
                res = id;
            }
        };
        for (; n_is_leaf(sc.nodes[node]); ++node) {
            uint64_t n = sc.nodes[node];
            if (c) c->leaf_nodes += 1;
            test(n_first(n));
            if (!n_has_second(n)) break;
            test(n_second(n));
        }
        return res;
    }

    // kdtree.cpp:515-578. early_exit=false is the reference (drains the stack);
    // early_exit=true clamps tenter to 0 and stops once the best hit lies inside the
    // current cell -- the schedule the GPU kernel uses (ids must be identical).
    uint32_t intersect(const Ray& ray, float& r, float& a, float& b, bool early_exit = false, Counters* c = nullptr) {
        Ray fixed = ray; // fix_direction, kdtree.cpp:503-511
        for (int ax = 0; ax < 3; ++ax)
            if (fixed.d[ax] == 0) fixed.d[ax] = kEps;
        if (c) {
            c->queries += 1;
            c->seen.clear();
        }
        float tenter, texit;
        if (!ray_box(fixed, sc.box, tenter, texit)) return kMissId;
        if (early_exit && tenter < 0) tenter = 0;
        stack.clear();
        stack.push_back({0u, tenter, texit});
        V3 dinv = {1 / fixed.d.x, 1 / fixed.d.y, 1 / fixed.d.z};
        uint32_t res = kMissId;
        r = kFltMax;
        while (!stack.empty()) {
            Item it = stack.back();
            stack.pop_back();
            uint32_t node = it.node;
            tenter = it.tenter;
            texit = it.texit;
            while (!n_is_leaf(sc.nodes[node])) {
                uint64_t n = sc.nodes[node];
                if (c) c->inner += 1;
                int ax = n_axis(n);
                float t = (n_split(n) - fixed.o[ax]) * dinv[ax];
                uint32_t near = node + 1, far = n_right(n);
                if (fixed.d[ax] <= 0) std::swap(near, far);
                if (texit < t) {
                    node = near;
                } else if (t < tenter) {
                    node = far;
                } else {
                    stack.push_back({far, t, texit});
                    node = near;
                    texit = t;
                }
            }
            float nr, na = 0, nb = 0;
            uint32_t next = leaf_run(node, ray, nr, na, nb, c);
            if (next != kMissId && nr < r) {
                res = next;
                r = nr;
                a = na;
                b = nb;
            }
            if (early_exit && res != kMissId && r <= texit) break;
        }
        return res;
    }

    // exhaustive scan in id order, strict <: the semantics the reference's result equals
    uint32_t brute_force(const Ray& ray, float& r, float& a, float& b) const {
        uint32_t res = kMissId;
        r = kFltMax;
        for (uint32_t id = 0; id < sc.tris.size(); ++id) {
            float rr, s, t;
            if (ray_triangle(ray, sc.tris[id], rr, s, t) && rr < r) {
                r = rr;
                a = s;
                b = t;
                res = id;
            }
        }
        return res;
    }
};

// ------------------------------------------------------------------------- RNG
struct XorShift64Star { // xorshift.h:27-62
    uint64_t s;
    uint64_t next() {
        s ^= s >> 12;
        s ^= s << 25;
        s ^= s >> 27;
        return s * 2685821657736338717ULL;
    }
    float nextf() { return std::ldexp(static_cast<float>(next() & 0xFFFFFFull), -24); }
};

// Counter-seeded streams (NOT in the reference; the product's parallel replacement for
// sampling.h:12's sequential thread-local stream, specified in DESIGN.md "RNG"):
// one xorshift64* state per path node, keyed by (run seed, primary sample index, node index).
inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline uint64_t node_state(uint64_t seed, uint64_t sample_index, uint64_t node) {
    uint64_t s = splitmix64(splitmix64(seed ^ splitmix64(sample_index)) + node);
    return s ? s : 0x9E3779B97F4A7C15ull;
}

inline float m2pi() { return static_cast<float>(2.f * M_PI); } // sampling.h:15
// sampling.h:20-32; (x,y,z) local direction, returns cos(theta) = u1
inline float hemisphere(float u1, float u2, V3& d) {
    float z = u1;
    float r = sqrtf(static_cast<float>(fmax(0.f, 1.f - z * z)));
    float phi = m2pi() * u2;
    d = {r * cosf(phi), r * sinf(phi), z};
    return u1;
}

// assimp matrix3x3.inl FromToMatrix with from = (0,0,1) (pathtracer.cpp:68-70), general code path kept.
struct M3 {
    float m[3][3];
};
inline M3 from_to(V3 from, V3 to) {
    M3 mt;
    const float e = from.x * to.x + from.y * to.y + from.z * to.z;
    const float f = (e < 0) ? -e : e;
    if (f > 1.0f - 0.00001f) {
        V3 u, v, x;
        x.x = (from.x > 0.0f) ? from.x : -from.x;
        x.y = (from.y > 0.0f) ? from.y : -from.y;
        x.z = (from.z > 0.0f) ? from.z : -from.z;
        if (x.x < x.y) {
            if (x.x < x.z) x = {1, 0, 0};
            else x = {0, 0, 1};
        } else {
            if (x.y < x.z) x = {0, 1, 0};
            else x = {0, 0, 1};
        }
        u = sub(x, from);
        v = sub(x, to);
        const float c1 = 2.0f / dot(u, u);
        const float c2 = 2.0f / dot(v, v);
        const float c3 = c1 * c2 * dot(u, v);
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < 3; j++) mt.m[i][j] = -c1 * u[i] * u[j] - c2 * v[i] * v[j] + c3 * v[i] * u[j];
            mt.m[i][i] += 1.0f;
        }
    } else {
        const V3 v = {from.y * to.z - from.z * to.y, from.z * to.x - from.x * to.z, from.x * to.y - from.y * to.x};
        const float h = 1.0f / (1.0f + e);
        const float hvx = h * v.x, hvz = h * v.z;
        const float hvxy = hvx * v.y, hvxz = hvx * v.z, hvyz = hvz * v.y;
        mt.m[0][0] = e + hvx * v.x;
        mt.m[0][1] = hvxy - v.z;
        mt.m[0][2] = hvxz + v.y;
        mt.m[1][0] = hvxy + v.z;
        mt.m[1][1] = e + h * v.y * v.y;
        mt.m[1][2] = hvyz - v.x;
        mt.m[2][0] = hvxz - v.y;
        mt.m[2][1] = hvyz + v.x;
        mt.m[2][2] = e + hvz * v.z;
    }
    return mt;
}
inline V3 mat_vec(const M3& a, V3 v) {
    return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}

// ----------------------------------------------------------------- integrators
struct RenderCfg {
    int32_t width, height;
    int32_t max_depth, mc_samples, pixel_samples, num_threads;
    int32_t integrator; // 0 pathtracer, 1 raycaster
    int32_t rng_mode;   // 0 reference streams (jitter 42/row, hemisphere thread-local seed 4), 1 counter-seeded
    uint64_t seed;      // rng_mode 1
    int32_t sample_begin, sample_stride; // render pixel samples i = begin, begin+stride, ... (< pixel_samples)
    float bg[4];
    float max_visibility;
    int32_t num_lights;
    float light_pos[3];
    float light_color[4];
    float cam_pos[3];
    float cam_rot[9]; // row-major 3x3 (a1..c3)
    float delta_x, delta_y;
    float shadow_intensity; // raytracer only (config.h:113)
};

struct RenderStats {
    uint64_t num_rays, num_prim_rays, num_shadow_rays;
    double runtime_ms;
};

struct Tracer {
    const Scene& sc;
    const RenderCfg& cfg;
    Traverser tv;
    XorShift64Star* stream; // rng_mode 0: the worker thread's hemisphere stream
    uint64_t rays = 0, shadow_rays = 0;
    Tracer(const Scene& s, const RenderCfg& c, XorShift64Star* st) : sc(s), cfg(c), tv(s), stream(st) {}

    // pathtracer.cpp:14-102
    C4 trace(const Ray& ray, int depth, uint64_t sample_index, uint64_t node) {
        if (depth > cfg.max_depth) return {0, 0, 0, 0};
        rays += 1;
        float dist, s, t;
        uint32_t id = tv.intersect(ray, dist, s, t);
        if (id == kMissId) return {cfg.bg[0], cfg.bg[1], cfg.bg[2], cfg.bg[3]};
        const Tri& tri = sc.tris[id];
        V3 p = add(ray.o, smul(dist, ray.d));
        float br = 1.f - s - t;
        V3 normal = normalize(add(add(smul(br, tri.n[0]), smul(s, tri.n[1])), smul(t, tri.n[2]))); // triangle.h:54-56
        V3 p2 = add(p, smul(0.0001f, normal));

        C4 direct = {0, 0, 0, 0};
        for (int li = 0; li < cfg.num_lights; ++li) {
            V3 lp = {cfg.light_pos[0], cfg.light_pos[1], cfg.light_pos[2]};
            V3 light_dir = normalize(sub(lp, p));
            float dist_to_light = length(sub(lp, p2));
            float dn;
            shadow_rays += 1;
            uint32_t sh = tv.intersect({p2, light_dir}, dn, s, t);
            if (sh == kMissId || dn > dist_to_light) {
                direct = cscale(std::max(0.f, dot(light_dir, normal)),
                                {cfg.light_color[0], cfg.light_color[1], cfg.light_color[2], cfg.light_color[3]});
            }
        }

        C4 indirect = {0, 0, 0, 0};
        M3 frame = from_to({0, 0, 1}, normal);
        for (int run = 0; run < cfg.mc_samples; ++run) {
            float u1, u2;
            uint64_t child = node * static_cast<uint64_t>(cfg.mc_samples) + static_cast<uint64_t>(run) + 1;
            if (cfg.rng_mode == 0) {
                u1 = stream->nextf();
                u2 = stream->nextf();
            } else {
                XorShift64Star g{node_state(cfg.seed, sample_index, child)};
                u1 = g.nextf();
                u2 = g.nextf();
            }
            V3 local;
            float cos_theta = hemisphere(u1, u2, local);
            V3 dir = mat_vec(frame, local);
            C4 li = trace({p2, dir}, depth + 1, sample_index, child);
            indirect = cadd(indirect, cscale(cos_theta, li));
        }
        float m = static_cast<float>(cfg.mc_samples);
        indirect = {indirect.r / m, indirect.g / m, indirect.b / m, indirect.a / m};
        return cmul(tri.diffuse, cadd(cscale(static_cast<float>(M_1_PI), direct), cscale(2.f, indirect)));
    }

    // raytracer.cpp:6-67 (Whitted: direct Lambert + mirror recursion + shadow attenuation); needs one light
    C4 raytrace(const Ray& ray, int depth) {
        rays += 1; // counted before the depth check (raytracer.cpp:9-13)
        if (depth > cfg.max_depth) return {0, 0, 0, 0};
        float dist, s, t;
        uint32_t id = tv.intersect(ray, dist, s, t);
        if (id == kMissId) return {cfg.bg[0], cfg.bg[1], cfg.bg[2], cfg.bg[3]};
        const V3 lp = {cfg.light_pos[0], cfg.light_pos[1], cfg.light_pos[2]};
        const C4 lc = {cfg.light_color[0], cfg.light_color[1], cfg.light_color[2], cfg.light_color[3]};
        V3 p = add(ray.o, smul(dist, ray.d));
        V3 light_dir = normalize(sub(lp, p));
        const Tri& tri = sc.tris[id];
        float br = 1.f - s - t;
        V3 normal = normalize(add(add(smul(br, tri.n[0]), smul(s, tri.n[1])), smul(t, tri.n[2])));
        // lambertian(L, N, C, I) = max(0, L.N) * C * I, lib/lambertian.h:15-19
        C4 direct = cmul(cscale(std::max(0.f, dot(light_dir, normal)), tri.diffuse), lc);
        V3 p2 = add(p, smul(0.0001f, normal));
        C4 color = direct;
        if (tri.reflectivity > 0) {
            V3 rdir = sub(ray.d, smul(2.f * dot(normal, ray.d), normal));
            C4 rc = raytrace({p2, rdir}, depth + 1);
            color = cadd(cscale(1.f - tri.reflectivity, direct), cmul(cscale(tri.reflectivity, tri.reflective), rc));
        }
        light_dir = normalize(sub(lp, p2));
        float dist_to_light = length(sub(lp, p2));
        float dn;
        shadow_rays += 1;
        uint32_t sh = tv.intersect({p2, light_dir}, dn, s, t);
        if (sh != kMissId && dn < dist_to_light) {
            C4 d = cscale(cfg.shadow_intensity, color);
            color = {color.r - d.r, color.g - d.g, color.b - d.b, color.a - d.a};
        }
        return color;
    }

    // raycaster.cpp:7-24
    C4 raycast(const Ray& ray) {
        float dist, s, t;
        uint32_t id = tv.intersect(ray, dist, s, t);
        if (id == kMissId) return {cfg.bg[0], cfg.bg[1], cfg.bg[2], cfg.bg[3]};
        rays += 1;
        C4 res = sc.tris[id].diffuse;
        float a = 1.f - (dist / cfg.max_visibility);
        res.a = a < 0.f ? 0.f : (1.f < a ? 1.f : a);
        return res;
    }
};

// lib/types.h:119-123 with aiMatrix3x3 * aiVector3D operand order
inline V3 raster2cam(const RenderCfg& c, float px, float py) {
    float w = static_cast<float>(c.width), h = static_cast<float>(c.height);
    V3 v = {-c.delta_x * (1 - 2 * px / w), c.delta_y * (1 - 2 * py / h), -1.f};
    const float* m = c.cam_rot;
    return {m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z};
}

} // namespace orc

// ======================================================================= C ABI
extern "C" {

using orc::Scene;

// reflective (n*4) / reflectivity (n) may be null (= 0): only the raytracer integrator reads them
void* orc_scene_create_ex(const float* verts, const float* normals, const float* diffuse, const float* reflective,
                          const float* reflectivity, uint32_t n);
void* orc_scene_create(const float* verts, const float* normals, const float* diffuse, uint32_t n) {
    return orc_scene_create_ex(verts, normals, diffuse, nullptr, nullptr, n);
}
void* orc_scene_create_ex(const float* verts, const float* normals, const float* diffuse, const float* reflective,
                          const float* reflectivity, uint32_t n) {
    auto* s = new Scene;
    s->tris.reserve(n);
    for (uint32_t i = 0; i < n; ++i)
        s->tris.push_back(orc::make_tri(verts + 9 * size_t(i), normals + 9 * size_t(i), diffuse + 4 * size_t(i),
                                        reflective ? reflective + 4 * size_t(i) : nullptr, reflectivity ? reflectivity[i] : 0.f));
    auto t0 = std::chrono::steady_clock::now();
    // KDTree::KDTree, kdtree.cpp:474-490
    s->box = orc::tri_bbox(s->tris[0]);
    std::vector<uint32_t> ids(n, 0);
    for (uint32_t i = 1; i < n; ++i) {
        s->box = orc::box_union(s->box, orc::tri_bbox(s->tris[i]));
        ids[i] = i;
    }
    orc::Builder b(s->tris);
    auto root = b.build(std::move(ids), s->box);
    s->nodes = orc::flatten(root.get());
    s->build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return s;
}

// adopt a node array built elsewhere (must be in the reference's FlatNode encoding)
void* orc_scene_create_prebuilt(const float* verts, const float* normals, const float* diffuse, uint32_t n,
                                const uint64_t* nodes, uint64_t num_nodes, const float* box6) {
    auto* s = new Scene;
    s->tris.reserve(n);
    for (uint32_t i = 0; i < n; ++i) s->tris.push_back(orc::make_tri(verts + 9 * size_t(i), normals + 9 * size_t(i), diffuse + 4 * size_t(i)));
    s->box = {{box6[0], box6[1], box6[2]}, {box6[3], box6[4], box6[5]}};
    s->nodes.assign(nodes, nodes + num_nodes);
    return s;
}

void orc_scene_destroy(void* h) { delete static_cast<Scene*>(h); }

void orc_scene_info(void* h, uint64_t* num_nodes, uint64_t* height, uint64_t* num_tris, float* box6, double* build_ms) {
    auto* s = static_cast<Scene*>(h);
    *num_nodes = s->nodes.size();
    *height = s->height();
    *num_tris = s->tris.size();
    box6[0] = s->box.lo.x; box6[1] = s->box.lo.y; box6[2] = s->box.lo.z;
    box6[3] = s->box.hi.x; box6[4] = s->box.hi.y; box6[5] = s->box.hi.z;
    if (build_ms) *build_ms = s->build_ms;
}

void orc_scene_nodes(void* h, uint64_t* out) {
    auto* s = static_cast<Scene*>(h);
    std::memcpy(out, s->nodes.data(), s->nodes.size() * 8);
}

// 48 floats per triangle, the reference's member order (triangle.h:95-118)
void orc_triangle_fields(void* h, uint32_t id, float* out48) {
    auto* s = static_cast<Scene*>(h);
    std::memcpy(out48, &s->tris[id], 192);
}

// mode 0: reference traversal; 1: early-exit schedule; 2: brute force over all triangles.
// counters (may be null): inner, leaf_nodes, tri_tests, queries
void orc_intersect(void* h, const float* o, const float* d, uint64_t n, int32_t mode, uint32_t* ids, float* rst,
                   uint64_t* counters) {
    auto* s = static_cast<Scene*>(h);
    orc::Traverser tv(*s);
    orc::Counters c;
    for (uint64_t i = 0; i < n; ++i) {
        orc::Ray ray{{o[3 * i], o[3 * i + 1], o[3 * i + 2]}, {d[3 * i], d[3 * i + 1], d[3 * i + 2]}};
        float r = 0, a = 0, b = 0;
        uint32_t id;
        if (mode == 2) id = tv.brute_force(ray, r, a, b);
        else id = tv.intersect(ray, r, a, b, mode == 1, counters ? &c : nullptr);
        ids[i] = id;
        if (id == orc::kMissId) r = a = b = 0.f;
        rst[3 * i] = r; rst[3 * i + 1] = a; rst[3 * i + 2] = b;
    }
    if (counters) {
        counters[0] = c.inner; counters[1] = c.leaf_nodes; counters[2] = c.tri_tests; counters[3] = c.queries;
        counters[4] = c.repeat4; counters[5] = c.repeat8; counters[6] = c.repeat_any;
    }
}

int orc_ray_box(const float* o, const float* d, const float* box6, float* tmin, float* tmax) {
    orc::Ray ray{{o[0], o[1], o[2]}, {d[0], d[1], d[2]}};
    orc::Box b{{box6[0], box6[1], box6[2]}, {box6[3], box6[4], box6[5]}};
    return orc::ray_box(ray, b, *tmin, *tmax) ? 1 : 0;
}

int orc_ray_triangle(const float* verts9, const float* o, const float* d, float* rst) {
    float zeros[9] = {0}, col[4] = {0};
    orc::Tri t = orc::make_tri(verts9, zeros, col);
    orc::Ray ray{{o[0], o[1], o[2]}, {d[0], d[1], d[2]}};
    return orc::ray_triangle(ray, t, rst[0], rst[1], rst[2]) ? 1 : 0;
}

void orc_clipped_box(const float* verts9, const float* box6, float* out6) {
    float zeros[9] = {0}, col[4] = {0};
    orc::Tri t = orc::make_tri(verts9, zeros, col);
    orc::Box b{{box6[0], box6[1], box6[2]}, {box6[3], box6[4], box6[5]}};
    orc::Box c = orc::clipped_tri_box(t, b);
    out6[0] = c.lo.x; out6[1] = c.lo.y; out6[2] = c.lo.z; out6[3] = c.hi.x; out6[4] = c.hi.y; out6[5] = c.hi.z;
}

void orc_xorshift_float(uint64_t seed, uint64_t n, float* out) {
    orc::XorShift64Star g{seed};
    for (uint64_t i = 0; i < n; ++i) out[i] = g.nextf();
}
void orc_xorshift_u64(uint64_t seed, uint64_t n, uint64_t* out) {
    orc::XorShift64Star g{seed};
    for (uint64_t i = 0; i < n; ++i) out[i] = g.next();
}
uint64_t orc_node_state(uint64_t seed, uint64_t sample_index, uint64_t node) { return orc::node_state(seed, sample_index, node); }

// first n draws of the reference's hemisphere stream (seed 4): out n*4 (x,y,z,cos)
void orc_hemisphere(uint64_t n, float* out) {
    orc::XorShift64Star g{4};
    for (uint64_t i = 0; i < n; ++i) {
        float u1 = g.nextf(), u2 = g.nextf();
        orc::V3 d;
        float c = orc::hemisphere(u1, u2, d);
        out[4 * i] = d.x; out[4 * i + 1] = d.y; out[4 * i + 2] = d.z; out[4 * i + 3] = c;
    }
}

// FromToMatrix((0,0,1) -> n) * v
void orc_frame_apply(const float* n3, const float* v3, float* out3) {
    orc::M3 m = orc::from_to({0, 0, 1}, {n3[0], n3[1], n3[2]});
    orc::V3 r = orc::mat_vec(m, {v3[0], v3[1], v3[2]});
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}

// Camera(trafo, aiCamera) ctor, lib/types.h:92-105: position = trafo * 0, 3x3 part, delta_x = tan(hfov) (double tan,
// rounded: the unqualified tan() there resolves to ::tan(double)), delta_y = delta_x / aspect; height = width / aspect.
void orc_camera_setup(const float* trafo4x4, float hfov, float aspect, int32_t width, float* cam_pos3, float* cam_rot9,
                      float* delta_xy2, int32_t* height) {
    const float* m = trafo4x4;
    cam_pos3[0] = m[0] * 0.f + m[1] * 0.f + m[2] * 0.f + m[3];
    cam_pos3[1] = m[4] * 0.f + m[5] * 0.f + m[6] * 0.f + m[7];
    cam_pos3[2] = m[8] * 0.f + m[9] * 0.f + m[10] * 0.f + m[11];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) cam_rot9[3 * r + c] = m[4 * r + c];
    delta_xy2[0] = static_cast<float>(tan(static_cast<double>(hfov)));
    delta_xy2[1] = delta_xy2[0] / aspect;
    *height = static_cast<int32_t>(width / aspect); // main.cpp:178-179 (int <- float)
}

// primary ray directions exactly as main.cpp:201-209 produces them, dirs[((y*W+x)*pps+i)*3]
void orc_primary_dirs(const orc::RenderCfg* cfg, float* dirs) {
    for (int y = 0; y < cfg->height; ++y) {
        orc::XorShift64Star gen{42};
        for (int x = 0; x < cfg->width; ++x)
            for (int i = 0; i < cfg->pixel_samples; ++i) {
                float dx = gen.nextf();
                float dy = gen.nextf();
                orc::V3 d = orc::raster2cam(*cfg, x + dx, y + dy);
                size_t k = (size_t(y) * cfg->width + x) * cfg->pixel_samples + i;
                dirs[3 * k] = d.x; dirs[3 * k + 1] = d.y; dirs[3 * k + 2] = d.z;
            }
    }
}

// Render loop, main.cpp:181-236. out_sum = per-pixel SUM over the rendered pixel samples of trace()
// (linear, before /pps + exposure + gamma); out_sumsq (nullable) = per-pixel sum of squares.
int orc_render(void* h, const orc::RenderCfg* cfgp, float* out_sum, float* out_sumsq, orc::RenderStats* stats) {
    auto* s = static_cast<Scene*>(h);
    const orc::RenderCfg cfg = *cfgp;
    std::atomic<int> next_row{0};
    std::atomic<uint64_t> rays{0}, prim{0}, shadow{0};
    auto t0 = std::chrono::steady_clock::now();
    auto worker = [&]() {
        orc::XorShift64Star hemi{4}; // sampling.h:12, one stream per (fresh) worker thread
        for (;;) {
            int y = next_row.fetch_add(1);
            if (y >= cfg.height) break;
            orc::Tracer tr(*s, cfg, &hemi); // one KDTreeIntersection per row task, main.cpp:197-198
            orc::XorShift64Star gen{42};    // main.cpp:201
            uint64_t nprim = 0;
            for (int x = 0; x < cfg.width; ++x) {
                orc::C4 sum = {0, 0, 0, 0}, sq = {0, 0, 0, 0};
                for (int i = 0; i < cfg.pixel_samples; ++i) {
                    float dx = gen.nextf();
                    float dy = gen.nextf();
                    if (i < cfg.sample_begin || (i - cfg.sample_begin) % cfg.sample_stride != 0) continue;
                    orc::V3 dir = orc::raster2cam(cfg, x + dx, y + dy);
                    orc::Ray ray{{cfg.cam_pos[0], cfg.cam_pos[1], cfg.cam_pos[2]}, dir};
                    nprim += 1;
                    uint64_t sample_index = (uint64_t(y) * cfg.width + x) * cfg.pixel_samples + i;
                    orc::C4 c = cfg.integrator == 0 ? tr.trace(ray, 0, sample_index, 0)
                                                    : (cfg.integrator == 1 ? tr.raycast(ray) : tr.raytrace(ray, 0));
                    sum = orc::cadd(sum, c);
                    sq = orc::cadd(sq, orc::cmul(c, c));
                }
                size_t k = (size_t(y) * cfg.width + x) * 4;
                out_sum[k] = sum.r; out_sum[k + 1] = sum.g; out_sum[k + 2] = sum.b; out_sum[k + 3] = sum.a;
                if (out_sumsq) {
                    out_sumsq[k] = sq.r; out_sumsq[k + 1] = sq.g; out_sumsq[k + 2] = sq.b; out_sumsq[k + 3] = sq.a;
                }
            }
            rays += tr.rays;
            shadow += tr.shadow_rays;
            prim += nprim;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < cfg.num_threads; ++t) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
    stats->runtime_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    stats->num_rays = rays;
    stats->num_prim_rays = prim;
    stats->num_shadow_rays = shadow;
    return 0;
}

// main.cpp:216-223 + effects.h:15-48: mean over pps, exposure, optional gamma (alpha untouched)
void orc_tonemap(const float* sum_rgba, uint64_t npix, int32_t pps, float exposure, int32_t gamma_enabled,
                 float inverse_gamma, float* out_rgba) {
    float n = static_cast<float>(pps);
    for (uint64_t i = 0; i < npix; ++i) {
        float c[4];
        for (int k = 0; k < 4; ++k) c[k] = sum_rgba[4 * i + k] / n;
        for (int k = 0; k < 3; ++k) c[k] = 1 - expf(-c[k] * exposure);
        if (gamma_enabled)
            for (int k = 0; k < 3; ++k) c[k] = powf(c[k], inverse_gamma);
        for (int k = 0; k < 4; ++k) out_rgba[4 * i + k] = c[k];
    }
}

// lib/raster.h:79-100 + main.cpp:242. Returns bytes needed; writes up to cap.
uint64_t orc_write_p3(const float* rgba, int32_t width, int32_t height, char* buf, uint64_t cap) {
    std::string out = "P3\n" + std::to_string(width) + " " + std::to_string(height) + "\n255";
    char tmp[64];
    size_t i = 0;
    auto q = [](float v) {
        float c = v < 0.f ? 0.f : (255.f < v ? 255.f : v);
        return static_cast<int>(c);
    };
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x, ++i) {
            const float* p = rgba + 4 * i;
            out += (i % width == 0) ? "\n" : " ";
            std::snprintf(tmp, sizeof tmp, "%3d %3d %3d", q(255 * p[0] * p[3]), q(255 * p[1] * p[3]), q(255 * p[2] * p[3]));
            out += tmp;
        }
    out += "\n";
    if (buf && cap) std::memcpy(buf, out.data(), out.size() < cap ? out.size() : cap);
    return out.size();
}

} // extern "C"
