// Host kd-tree builder (see kdtree_build.h). Same decisions, in the same fp32
// arithmetic, as the reference's KDTreeBuildAlgorithm (lib/kdtree.cpp:124-408),
// clip_triangle_at_aabb (lib/clipping.h:199-235) and flatten (lib/kdtree.cpp:420-467);
// different machinery: fixed-size polygon buffers, per-axis plane tests written
// out for axis-aligned planes, subtree tasks on a bounded set of threads.
#include "kdtree_build.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>

namespace trn {
namespace {

constexpr float kEps = 0.00001f; // lib/types.h:13
constexpr float kMax = std::numeric_limits<float>::max();
constexpr float kLowest = std::numeric_limits<float>::lowest();
constexpr uint32_t kInvalid = 0xFFFFFFFFu >> 2; // lib/kdtree.h:67

struct Aabb {
    float lo[3], hi[3];
};

inline float area_of(const Aabb& b) { // geometry.h:1044-1047
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return 2 * (dx * dy + dx * dz + dy * dz);
}
inline float pick_min(float a, float b) { return b < a ? b : a; } // std::min
inline float pick_max(float a, float b) { return a < b ? b : a; } // std::max
// Bbox3::split + the min/max-normalising Bbox3(p1,p2) ctor, geometry.h:998-999,1064-1076
inline void split_box(const Aabb& b, int ax, float pos, Aabb& l, Aabb& r) {
    l = b;
    r = b;
    l.lo[ax] = pick_min(b.lo[ax], pos);
    l.hi[ax] = pick_max(b.lo[ax], pos);
    r.lo[ax] = pick_min(pos, b.hi[ax]);
    r.hi[ax] = pick_max(pos, b.hi[ax]);
}

struct Pt {
    float c[3];
};

inline int side_class(float dist) { return dist > kEps ? 1 : (dist < -kEps ? -1 : 0); }

// One Sutherland-Hodgman pass against the thick plane  c[ax] >= bound (upper=false)
// or c[ax] <= bound (upper=true); lib/clipping.h:135-187 specialised to axis planes.
inline int clip_pass(const Pt* in, int n, int ax, bool upper, float bound, Pt* out) {
    int m = 0;
    Pt a = in[n - 1];
    int sa = side_class(upper ? bound - a.c[ax] : a.c[ax] - bound);
    for (int i = 0; i < n; ++i) {
        const Pt b = in[i];
        int sb = side_class(upper ? bound - b.c[ax] : b.c[ax] - bound);
        if ((sb == 1 && sa == -1) || (sb == -1 && sa == 1)) {
            float t = upper ? (bound - a.c[ax]) / (b.c[ax] - a.c[ax]) : (bound - a.c[ax]) / (b.c[ax] - a.c[ax]);
            Pt x;
            for (int k = 0; k < 3; ++k) x.c[k] = a.c[k] + t * (b.c[k] - a.c[k]);
            out[m++] = x;
        }
        if (sb != -1) out[m++] = b;
        a = b;
        sa = sb;
    }
    return m;
}

// Bounding box of the triangle clipped to `box`. Fewer than two points left => the
// "everything" box (the reference returns a default-constructed Bbox3 there).
inline Aabb clipped_bounds(const float* tri9, const Aabb& box) {
    Pt bufa[16], bufb[16];
    Pt* cur = bufa;
    Pt* nxt = bufb;
    for (int k = 0; k < 3; ++k)
        for (int c = 0; c < 3; ++c) cur[k].c[c] = tri9[3 * k + c];
    int n = 3;
    Aabb res;
    for (int ax = 0; ax < 3; ++ax) {
        for (int side = 0; side < 2; ++side) {
            n = clip_pass(cur, n, ax, side == 1, side == 0 ? box.lo[ax] : box.hi[ax], nxt);
            std::swap(cur, nxt);
            if (n < 2) {
                for (int c = 0; c < 3; ++c) {
                    res.lo[c] = kLowest;
                    res.hi[c] = kMax;
                }
                return res;
            }
        }
    }
    for (int c = 0; c < 3; ++c) {
        float mn = kMax, mx = kLowest;
        for (int i = 0; i < n; ++i) {
            if (cur[i].c[c] < mn) mn = cur[i].c[c];
            if (mx < cur[i].c[c]) mx = cur[i].c[c];
        }
        res.lo[c] = pick_min(mn, mx);
        res.hi[c] = pick_max(mn, mx);
    }
    return res;
}

struct Event {
    uint32_t id;
    float point;
    float aux;
    int type; // 0 end, 1 planar, 2 start -- sort order at equal position (lib/kdtree.cpp:254-256,311-316)
};

inline uint32_t float_bits(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}

struct BuildNode {
    int axis = -1;
    float split = 0;
    bool cut = false; // plane with nothing on one side: dropped by the reference layout, kept by the device layout
    BuildNode* left = nullptr;
    BuildNode* right = nullptr;
    std::vector<uint32_t> ids;
};

struct Context {
    const float* verts; // n*9
    std::atomic<int> spare_threads{0};
};

inline float bias(size_t nl, size_t nr) { return (nl == 0 || nr == 0) ? 0.8f : 1.f; } // lib/kdtree.cpp:182-187
inline float split_cost(float lr, float rr, size_t nl, size_t nr) {                     // lib/kdtree.cpp:197-203
    return bias(nl, nr) * (15 + 20 * (lr * nl + rr * nr));
}

struct PlaneChoice {
    float cost = kMax;
    int axis = 0;
    float pos = 0;
    bool planar_left = true;
};

void sweep_axis(std::vector<Event>& ev, int ax, const Aabb& box, float area, size_t total, PlaneChoice& best) {
    std::sort(ev.begin(), ev.end(), [](const Event& a, const Event& b) {
        return a.point < b.point || (a.point == b.point && a.type < b.type);
    });
    size_t nl = 0, nr = total;
    const size_t n = ev.size();
    for (size_t i = 0; i < n;) {
        const float p = ev[i].point;
        size_t ending = 0, planar = 0, starting = 0;
        while (i < n && ev[i].point == p && ev[i].type == 0) { ++ending; ++i; }
        while (i < n && ev[i].point == p && ev[i].type == 1) { ++planar; ++i; }
        while (i < n && ev[i].point == p && ev[i].type == 2) { ++starting; ++i; }
        if (ending + planar + starting == 0) { ++i; continue; } // NaN guard; inputs are finite
        nr -= planar + ending;
        Aabb lb, rb;
        split_box(box, ax, p, lb, rb);
        const float lr = area_of(lb) / area;
        const float rr = area_of(rb) / area;
        const float cl = split_cost(lr, rr, nl + planar, nr);
        const float cr = split_cost(lr, rr, nl, planar + nr);
        const bool left = cl < cr;
        const float c = left ? cl : cr;
        if (c < best.cost) {
            best.cost = c;
            best.axis = ax;
            best.pos = p;
            best.planar_left = left;
        }
        nl += starting + planar;
    }
}

BuildNode* make_leaf(std::vector<uint32_t>&& ids) {
    BuildNode* n = new BuildNode;
    n->ids = std::move(ids);
    return n;
}

BuildNode* build_rec(Context& ctx, std::vector<uint32_t> ids, const Aabb& box) {
    if (ids.empty()) return nullptr;
    if (ids.size() <= 3) return make_leaf(std::move(ids));
    const float area = area_of(box);
    if (area == 0) return make_leaf(std::move(ids));

    // events from clipped boxes
    std::vector<Event> ev[3];
    for (auto& e : ev) e.reserve(ids.size() * 2);
    size_t total = 0;
    for (uint32_t id : ids) {
        const Aabb cb = clipped_bounds(ctx.verts + 9 * size_t(id), box);
        if (cb.hi[0] <= cb.lo[0] && cb.hi[1] <= cb.lo[1] && cb.hi[2] <= cb.lo[2]) continue; // Bbox3::empty()
        ++total;
        for (int ax = 0; ax < 3; ++ax) {
            if (std::abs(cb.hi[ax] - cb.lo[ax]) < kEps) { // Bbox3::planar()
                ev[ax].push_back({id, cb.lo[ax], cb.lo[ax], 1});
            } else {
                ev[ax].push_back({id, cb.lo[ax], cb.hi[ax], 2});
                ev[ax].push_back({id, cb.hi[ax], cb.lo[ax], 0});
            }
        }
    }

    PlaneChoice best;
    std::vector<uint32_t> lt, rt;
    if (total > 0) {
        // the three axis sweeps are independent; the winner is the first strict minimum in
        // X, Y, Z order, which merging per-axis minima in that order reproduces.
        PlaneChoice per_axis[3];
        const bool big = ids.size() >= 200000;
        std::thread helpers[2];
        int used = 0;
        for (int ax = 0; ax < 3; ++ax) {
            bool spawned = false;
            if (big && ax < 2 && ctx.spare_threads.fetch_sub(1) > 0) {
                helpers[used++] = std::thread([&, ax]() { sweep_axis(ev[ax], ax, box, area, total, per_axis[ax]); });
                spawned = true;
            } else if (big && ax < 2) {
                ctx.spare_threads.fetch_add(1);
            }
            if (!spawned) sweep_axis(ev[ax], ax, box, area, total, per_axis[ax]);
        }
        for (int i = 0; i < used; ++i) {
            helpers[i].join();
            ctx.spare_threads.fetch_add(1);
        }
        for (int ax = 0; ax < 3; ++ax)
            if (per_axis[ax].cost < best.cost) best = per_axis[ax];

        // classification along the chosen axis, in event order (lib/kdtree.cpp:370-399)
        for (const Event& e : ev[best.axis]) {
            if (e.point < best.pos) {
                if (e.type != 2) lt.push_back(e.id);
                else if (best.pos < e.aux) { lt.push_back(e.id); rt.push_back(e.id); }
            } else if (e.point == best.pos) {
                if (e.type == 0) lt.push_back(e.id);
                else if (e.type == 1) (best.planar_left ? lt : rt).push_back(e.id);
                else rt.push_back(e.id);
            } else if (e.type != 0) {
                rt.push_back(e.id);
            }
        }
    }
    for (auto& e : ev) std::vector<Event>().swap(e);

    // automatic termination (lib/kdtree.cpp:151-155)
    if (20 * ids.size() * bias(lt.size(), rt.size()) < best.cost) return make_leaf(std::move(ids));
    std::vector<uint32_t>().swap(ids);

    Aabb lb, rb;
    split_box(box, best.axis, best.pos, lb, rb);

    BuildNode* l = nullptr;
    BuildNode* r = nullptr;
    if (lt.size() + rt.size() >= 4096 && ctx.spare_threads.fetch_sub(1) > 0) {
        std::thread th([&]() { l = build_rec(ctx, std::move(lt), lb); });
        r = build_rec(ctx, std::move(rt), rb);
        th.join();
        ctx.spare_threads.fetch_add(1);
    } else {
        if (lt.size() + rt.size() >= 4096) ctx.spare_threads.fetch_add(1);
        l = build_rec(ctx, std::move(lt), lb);
        r = build_rec(ctx, std::move(rt), rb);
    }
    if (!l && !r) return nullptr;
    BuildNode* n = new BuildNode;
    n->axis = best.axis;
    n->split = best.pos;
    n->left = l;
    n->right = r;
    n->cut = !l || !r; // lib/kdtree.cpp:168-172 would return the surviving child here
    return n;
}

inline BuildNode* skip_cuts(BuildNode* n) {
    while (n && n->cut) n = n->left ? n->left : n->right;
    return n;
}

// Sibling-pair device layout, cuts included (see kdtree_build.h).
void flatten_pairs(BuildNode* root, KdTree& out) {
    float scale = 0.f;
    for (int c = 0; c < 6; ++c) scale = std::fmax(scale, std::fabs(out.box[c]));
    auto& nodes = out.pair_nodes;
    auto& refs = out.pair_leaf_refs;
    nodes.assign(2, 0);
    nodes[1] = 3; // padding sibling of the root: an empty leaf
    struct Item {
        BuildNode* node;
        uint32_t slot;
    };
    out.num_cut_nodes = 0;
    out.num_pair_refs = 0;
    auto emit = [&](const Item& it, std::vector<Item>& children) {
        BuildNode* n = it.node;
        if (!n) {
            nodes[it.slot] = static_cast<uint64_t>(3u) << 32; // empty leaf: count 0
        } else if (n->axis >= 0) {
            if (n->cut) ++out.num_cut_nodes;
            const uint32_t pair = static_cast<uint32_t>(nodes.size());
            nodes.push_back(0);
            nodes.push_back(0);
            float split = n->split;
            if (n->cut) {
                // A cut plane exists only in this layout (the reference walks on into the surviving child with the
                // parent's interval). Move it a few ulp of the scene scale INTO the void, so that a ray whose interval
                // ends or starts within rounding noise of the plane still visits the solid side.
                const float shift = 4e-6f * std::fmax(std::fabs(split), scale);
                split += n->left ? shift : -shift; // solid left: void is above the plane, and vice versa
            }
            nodes[it.slot] = (static_cast<uint64_t>((pair << 2) | static_cast<uint32_t>(n->axis)) << 32) | float_bits(split);
            children.push_back({n->left, pair});
            children.push_back({n->right, pair + 1});
        } else {
            // every leaf's run starts at a multiple of 4 references (16 bytes), so that a kernel can fetch four ids with
            // one vector load; the padding repeats the previous run's last id (a valid triangle, never counted)
            while (refs.size() % 4 != 0) refs.push_back(refs.back());
            const uint32_t first = static_cast<uint32_t>(refs.size());
            refs.insert(refs.end(), n->ids.begin(), n->ids.end());
            out.num_pair_refs += n->ids.size();
            nodes[it.slot] = (static_cast<uint64_t>((static_cast<uint32_t>(n->ids.size()) << 2) | 3u) << 32) | first;
        }
    };
    // top of the tree breadth-first, so that the first kTreeletNodes nodes are the top treelet (stageable as one block)
    std::vector<Item> level{{root, 0}}, next_level;
    while (!level.empty() && nodes.size() + 2 * level.size() <= KdTree::kTreeletNodes) {
        next_level.clear();
        for (const Item& it : level) emit(it, next_level);
        level.swap(next_level);
    }
    // the rest depth-first (a subtree stays contiguous)
    std::vector<Item> stack(level.rbegin(), level.rend()), kids;
    while (!stack.empty()) {
        Item it = stack.back();
        stack.pop_back();
        kids.clear();
        emit(it, kids);
        for (auto k = kids.rbegin(); k != kids.rend(); ++k) stack.push_back(*k);
    }
    for (int k = 0; k < 4; ++k) refs.push_back(refs.empty() ? 0u : refs.back()); // a 4-wide read of the last chunk stays inside
}

// DFS layout of lib/kdtree.cpp:420-467 in the FlatNode encoding of lib/kdtree.h:62-154.
void flatten_tree(BuildNode* root_with_cuts, KdTree& out) {
    BuildNode* root = skip_cuts(root_with_cuts);
    struct Item {
        BuildNode* node;
        uint32_t parent;
        uint32_t level;
    };
    std::vector<Item> stack;
    stack.push_back({root, kInvalid, 0});
    auto& nodes = out.nodes;
    out.height = 0;
    out.num_leaf_refs = 0;
    while (!stack.empty()) {
        Item it = stack.back();
        stack.pop_back();
        if (it.level > out.height) out.height = it.level;
        const uint32_t idx = static_cast<uint32_t>(nodes.size());
        if (it.parent != kInvalid) {
            uint64_t p = nodes[it.parent];
            nodes[it.parent] = (p & 0xFFFFFFFF00000000ull) | static_cast<uint32_t>((idx << 2) | (static_cast<uint32_t>(p) & 3u));
        }
        BuildNode* n = it.node;
        if (n->axis >= 0) {
            nodes.push_back((static_cast<uint64_t>(float_bits(n->split)) << 32) |
                            static_cast<uint32_t>((kInvalid << 2) | static_cast<uint32_t>(n->axis)));
            stack.push_back({skip_cuts(n->right), idx, it.level + 1});
            stack.push_back({skip_cuts(n->left), kInvalid, it.level + 1});
        } else {
            const auto& ids = n->ids;
            out.num_leaf_refs += ids.size();
            size_t i = 1;
            for (; i < ids.size(); i += 2)
                nodes.push_back((static_cast<uint64_t>(ids[i - 1]) << 32) | static_cast<uint32_t>((ids[i] << 2) | 3u));
            if (i - 1 < ids.size())
                nodes.push_back((static_cast<uint64_t>(ids[i - 1]) << 32) | 0xFFFFFFFFull);
            else
                nodes.push_back(0); // all-zero inner node terminates the leaf run
        }
    }
}

void free_tree(BuildNode* root) {
    std::vector<BuildNode*> stack{root};
    while (!stack.empty()) {
        BuildNode* n = stack.back();
        stack.pop_back();
        if (!n) continue;
        stack.push_back(n->left);
        stack.push_back(n->right);
        delete n;
    }
}

} // namespace

void precompute_triangles(const float* verts, const float* normals, const float* diffuse, uint32_t n, HostTriangles& out,
                          const float* reflective, const float* reflectivity) {
    out.count = n;
    out.mirror.assign(size_t(n) * 4, 0.f);
    if (reflective) std::memcpy(out.mirror.data(), reflective, size_t(n) * 4 * sizeof(float));
    out.verts.assign(verts, verts + size_t(n) * 9);
    out.isect.resize(size_t(n) * 16);
    out.shade.resize(size_t(n) * 16);
    for (uint32_t i = 0; i < n; ++i) {
        const float* p = verts + size_t(i) * 9;
        float* r = &out.isect[size_t(i) * 16];
        // lib/triangle.h:33-39 with geometry.h's conventions: edges, cross in double rounded
        // once, normalize as (1/len)*v, dots as left-to-right sums.
        const float ux = p[3] - p[0], uy = p[4] - p[1], uz = p[5] - p[2];
        const float vx = p[6] - p[0], vy = p[7] - p[1], vz = p[8] - p[2];
        const double dux = ux, duy = uy, duz = uz, dvx = vx, dvy = vy, dvz = vz;
        const float cx = static_cast<float>((duy * dvz) - (duz * dvy));
        const float cy = static_cast<float>((duz * dvx) - (dux * dvz));
        const float cz = static_cast<float>((dux * dvy) - (duy * dvx));
        const float inv = 1 / std::sqrt(cx * cx + cy * cy + cz * cz);
        r[0] = p[0]; r[1] = p[1]; r[2] = p[2];
        r[3] = inv * cx; r[4] = inv * cy; r[5] = inv * cz;
        r[6] = ux; r[7] = uy; r[8] = uz;
        r[9] = vx; r[10] = vy; r[11] = vz;
        const float uv = ux * vx + uy * vy + uz * vz;
        const float vv = vx * vx + vy * vy + vz * vz;
        const float uu = ux * ux + uy * uy + uz * uz;
        r[12] = uv; r[13] = vv; r[14] = uu;
        r[15] = uv * uv - uu * vv;
        float* s = &out.shade[size_t(i) * 16];
        std::memcpy(s, normals + size_t(i) * 9, 9 * sizeof(float));
        s[9] = reflectivity ? reflectivity[i] : 0.f;
        s[10] = s[11] = 0.f;
        std::memcpy(s + 12, diffuse + size_t(i) * 4, 4 * sizeof(float));
    }
}

void reference_shape_from_pairs(KdTree& tree) {
    const auto& pn = tree.pair_nodes;
    const auto& refs = tree.pair_leaf_refs;
    auto word_x = [&](uint32_t i) { return static_cast<uint32_t>(pn[i]); };
    auto word_y = [&](uint32_t i) { return static_cast<uint32_t>(pn[i] >> 32); };
    auto is_void = [&](uint32_t i) { return word_y(i) == 3u; };
    // follow empty-space cuts: an inner node with a void child is its other child (lib/kdtree.cpp:168-172)
    auto resolve = [&](uint32_t i) {
        for (;;) {
            const uint32_t y = word_y(i);
            if ((y & 3u) == 3u) return i;
            const uint32_t pair = y >> 2;
            if (is_void(pair)) i = pair + 1;
            else if (is_void(pair + 1)) i = pair;
            else return i;
        }
    };
    struct Item {
        uint32_t node, parent, level;
    };
    auto& nodes = tree.nodes;
    nodes.clear();
    tree.height = 0;
    tree.num_leaf_refs = 0;
    std::vector<Item> stack;
    stack.push_back({resolve(0), kInvalid, 0});
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        if (it.level > tree.height) tree.height = it.level;
        const uint32_t idx = static_cast<uint32_t>(nodes.size());
        if (it.parent != kInvalid) {
            const uint64_t p = nodes[it.parent];
            nodes[it.parent] = (p & 0xFFFFFFFF00000000ull) | static_cast<uint32_t>((idx << 2) | (static_cast<uint32_t>(p) & 3u));
        }
        const uint32_t y = word_y(it.node);
        if ((y & 3u) != 3u) {
            const uint32_t pair = y >> 2;
            nodes.push_back((static_cast<uint64_t>(word_x(it.node)) << 32) | static_cast<uint32_t>((kInvalid << 2) | (y & 3u)));
            stack.push_back({resolve(pair + 1), idx, it.level + 1});
            stack.push_back({resolve(pair), kInvalid, it.level + 1});
        } else {
            const uint32_t first = word_x(it.node), cnt = y >> 2;
            tree.num_leaf_refs += cnt;
            uint32_t i = 1;
            for (; i < cnt; i += 2)
                nodes.push_back((static_cast<uint64_t>(refs[first + i - 1]) << 32) | static_cast<uint32_t>((refs[first + i] << 2) | 3u));
            if (i - 1 < cnt)
                nodes.push_back((static_cast<uint64_t>(refs[first + i - 1]) << 32) | 0xFFFFFFFFull);
            else
                nodes.push_back(0); // all-zero inner node terminates the leaf run
        }
    }
}

void build_kdtree(const HostTriangles& tris, KdTree& out, int num_threads) {
    auto t0 = std::chrono::steady_clock::now();
    const uint32_t n = tris.count;
    const float* v = tris.verts.data();
    // KDTree::KDTree, lib/kdtree.cpp:474-490 (triangle bbox via fmin/fmax, union via min/max)
    Aabb box;
    for (uint32_t i = 0; i < n; ++i) {
        const float* p = v + size_t(i) * 9;
        for (int c = 0; c < 3; ++c) {
            float mn = std::fmin(p[c], std::min(p[3 + c], p[6 + c]));
            float mx = std::fmax(p[c], std::max(p[3 + c], p[6 + c]));
            float lo = pick_min(mn, mx), hi = pick_max(mn, mx);
            if (i == 0) {
                box.lo[c] = lo;
                box.hi[c] = hi;
            } else {
                box.lo[c] = pick_min(box.lo[c], lo);
                box.hi[c] = pick_max(box.hi[c], hi);
            }
        }
    }
    std::vector<uint32_t> ids(n);
    for (uint32_t i = 0; i < n; ++i) ids[i] = i;
    Context ctx;
    ctx.verts = v;
    int hw = num_threads > 0 ? num_threads : static_cast<int>(std::thread::hardware_concurrency());
    if (hw < 1) hw = 1;
    ctx.spare_threads = hw - 1;
    BuildNode* root = build_rec(ctx, std::move(ids), box);
    out.nodes.clear();
    for (int c = 0; c < 3; ++c) {
        out.box[c] = box.lo[c];
        out.box[3 + c] = box.hi[c];
    }
    flatten_tree(root, out);
    flatten_pairs(root, out);
    free_tree(root);
    out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

} // namespace trn
