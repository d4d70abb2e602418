// TEST INFRASTRUCTURE ONLY. Never linked into, imported by or executed from the
// product path (turner_b200/). Allowed users: tests/, __graft_entry__.smoke(),
// and bench.py's cpu_baseline / --impl reference legs.
//
// C-ABI driver around the UNMODIFIED reference hot path. It is compiled by
// oracle/build_ref.sh together with the reference's own translation units
//   /root/reference/lib/kdtree.cpp  and  /root/reference/{pathtracer,raycaster}.cpp
// (read where they lie; nothing is copied into the repo) into
//   oracle/_ref/libturner_ref_{pathtracer,raycaster}.so
// This file contains no reference code: it only *calls* the reference's
// KDTree / KDTreeIntersection / trace() / Camera / xorshift64star / exposure /
// gamma / Image writer, and re-states the ~25-line render loop body of
// main.cpp:181-236 (which cannot be compiled here: main.cpp needs assimp's
// importer, docopt, cereal and ThreadPool, all network-fetched).
#include "trace.h" // reference: declares trace(), pulls config.h, lib/kdtree.h, lib/types.h

#include "lib/effects.h"
#include "lib/intersection.h"
#include "lib/raster.h"
#include "lib/runtime.h"
#include "lib/sampling.h"
#include "lib/stats.h"
#include "lib/xorshift.h"

#include <atomic>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>
#include <thread>
#include <fstream>
#include <cereal/archives/portable_binary.hpp>
#include <vector>

// The reference TUs define USAGE; keep the linker happy if they are absent.
extern const char* USAGE;

namespace {

// In-memory "archive" for the reference's cereal serialize() hooks
// (lib/kdtree.h:228-230, lib/triangle.h:89-92, lib/kdtree.h:136-138). It does
// what main.cpp:147-152 does with a PortableBinaryInputArchive: hands out the
// members so a pre-built tree can be loaded instead of running the builder.
struct MemberGrabber {
    Triangles* tris = nullptr;
    Bbox3f* box = nullptr;
    std::vector<detail::FlatNode>* nodes = nullptr;
    void take(Triangles& t) { tris = &t; }
    void take(Bbox3f& b) { box = &b; }
    void take(std::vector<detail::FlatNode>& n) { nodes = &n; }
    template <class... A> void operator()(A&... a) {
        int dummy[] = {(take(a), 0)...};
        (void)dummy;
    }
};

struct FloatGrabber {
    std::vector<float> vals;
    void take(float& f) { vals.push_back(f); }
    void take(Vector3f& v) { vals.push_back(v.x); vals.push_back(v.y); vals.push_back(v.z); }
    void take(Normal3f& v) { vals.push_back(v.x); vals.push_back(v.y); vals.push_back(v.z); }
    void take(aiColor4D& c) { vals.push_back(c.r); vals.push_back(c.g); vals.push_back(c.b); vals.push_back(c.a); }
    void take(std::array<Point3f, 3>& a) { for (auto& p : a) { vals.push_back(p.x); vals.push_back(p.y); vals.push_back(p.z); } }
    void take(std::array<Normal3f, 3>& a) { for (auto& p : a) { vals.push_back(p.x); vals.push_back(p.y); vals.push_back(p.z); } }
    template <class... A> void operator()(A&... a) {
        int dummy[] = {(take(a), 0)...};
        (void)dummy;
    }
};

struct RefScene {
    KDTree tree;
};

Triangles make_triangles(const float* verts, const float* normals, const float* diffuse, uint32_t n,
                         const float* reflective = nullptr, const float* reflectivity = nullptr) {
    Triangles tris;
    tris.reserve(n);
    for (uint32_t i = 0; i < n; ++i) {
        const float* v = verts + 9 * size_t(i);
        const float* nn = normals + 9 * size_t(i);
        const float* c = diffuse + 4 * size_t(i);
        aiColor4D dif(c[0], c[1], c[2], c[3]);
        tris.push_back(Triangle{{Point3f(v[0], v[1], v[2]), Point3f(v[3], v[4], v[5]), Point3f(v[6], v[7], v[8])},
                                {Normal3f(nn[0], nn[1], nn[2]), Normal3f(nn[3], nn[4], nn[5]),
                                 Normal3f(nn[6], nn[7], nn[8])},
                                aiColor4D(),
                                dif,
                                dif,
                                reflective ? aiColor4D(reflective[4 * size_t(i)], reflective[4 * size_t(i) + 1],
                                                       reflective[4 * size_t(i) + 2], reflective[4 * size_t(i) + 3])
                                           : aiColor4D(),
                                reflectivity ? reflectivity[i] : 0.f});
    }
    return tris;
}

} // namespace

extern "C" {

struct RefCamera {
    float trafo4x4[16]; // row-major a1..d4 (assimp order), node transformation of the camera
    float hfov;         // aiCamera::mHorizontalFOV as the loader would set it
    float aspect;       // aiCamera::mAspect after main.cpp:112-116
};

struct RefConfig {
    int32_t width;
    int32_t max_depth;
    int32_t mc_samples;
    int32_t pixel_samples;
    int32_t num_threads;
    int32_t gamma_enabled;
    float bg[4];
    float exposure;
    float inverse_gamma;
    float max_visibility;
    int32_t num_lights; // 0 or 1
    float light_pos[3];
    float light_color[4];
    float shadow_intensity; // raytracer only (config.h:113)
    // timing sampler of bench.py's reference arm: render only rows y = row_begin + k * row_stride (whole rows, every
    // pixel sample of them -- a row is the reference's unit of work, main.cpp:194). 0 / 0 or 0 / 1 = the whole image.
    int32_t row_begin;
    int32_t row_stride;
};

struct RefStats {
    uint64_t num_rays;
    uint64_t num_prim_rays;
    uint64_t runtime_ms;
    uint64_t width;
    uint64_t height;
};

const char* ref_usage() { return USAGE; }

void* ref_scene_create(const float* verts, const float* normals, const float* diffuse, uint32_t n) {
    auto* s = new RefScene;
    s->tree = KDTree(make_triangles(verts, normals, diffuse, n)); // reference builder, lib/kdtree.cpp:474-490
    return s;
}

void* ref_scene_create_ex(const float* verts, const float* normals, const float* diffuse, const float* reflective,
                          const float* reflectivity, uint32_t n) {
    auto* s = new RefScene;
    s->tree = KDTree(make_triangles(verts, normals, diffuse, n, reflective, reflectivity));
    return s;
}

// Load a tree built elsewhere through the reference's own serialize() hook
// (the same entry main.cpp:147-152 uses for kdtree.cache).
void* ref_scene_create_prebuilt(const float* verts, const float* normals, const float* diffuse, uint32_t n,
                                const uint64_t* nodes, uint64_t num_nodes, const float* box6) {
    auto* s = new RefScene;
    MemberGrabber g;
    s->tree.serialize(g);
    *g.tris = make_triangles(verts, normals, diffuse, n);
    *g.box = Bbox3f(Point3f(box6[0], box6[1], box6[2]), Point3f(box6[3], box6[4], box6[5]));
    g.nodes->resize(num_nodes);
    static_assert(sizeof(detail::FlatNode) == 8, "FlatNode is 8 bytes");
    std::memcpy(static_cast<void*>(g.nodes->data()), nodes, num_nodes * 8);
    return s;
}

// kdtree.cache, as main.cpp:158-165 writes it and main.cpp:147-152 reads it: the reference's KDTree::serialize() through
// a PortableBinary archive (oracle/ref_shims/cereal/archives/portable_binary.hpp restates cereal's byte layer).
int ref_cache_write(void* h, const char* path) {
    auto* s = static_cast<RefScene*>(h);
    std::ofstream output_file;
    output_file.open(path, std::ios::out | std::ios::binary);
    if (!output_file.is_open()) return -1;
    cereal::PortableBinaryOutputArchive oarchive(output_file);
    oarchive(s->tree);
    return output_file.good() ? 0 : -1;
}

void* ref_cache_read(const char* path) {
    std::ifstream kdtree_cache(path, std::ios::in | std::ios::binary);
    if (!kdtree_cache.is_open()) return nullptr;
    auto* s = new RefScene;
    try {
        cereal::PortableBinaryInputArchive iarchive(kdtree_cache);
        iarchive(s->tree);
    } catch (const std::exception&) {
        delete s;
        return nullptr;
    }
    return s;
}

void ref_scene_destroy(void* h) { delete static_cast<RefScene*>(h); }

void ref_scene_info(void* h, uint64_t* num_nodes, uint64_t* height, uint64_t* num_tris, float* box6) {
    auto* s = static_cast<RefScene*>(h);
    *num_nodes = s->tree.num_nodes();
    *height = s->tree.height();
    *num_tris = s->tree.num_triangles();
    const auto& b = s->tree.box();
    box6[0] = b.p_min.x; box6[1] = b.p_min.y; box6[2] = b.p_min.z;
    box6[3] = b.p_max.x; box6[4] = b.p_max.y; box6[5] = b.p_max.z;
}

void ref_scene_nodes(void* h, uint64_t* out) {
    auto* s = static_cast<RefScene*>(h);
    MemberGrabber g;
    s->tree.serialize(g);
    std::memcpy(out, static_cast<const void*>(g.nodes->data()), g.nodes->size() * 8);
}

// 48 floats per triangle in serialize() order: vertices(9) normals(9) ambient(4)
// diffuse(4) emissive(4) reflective(4) reflectivity(1) u(3) v(3) normal(3) uv vv uu denom
void ref_triangle_fields(void* h, uint32_t id, float* out48) {
    auto* s = static_cast<RefScene*>(h);
    Triangle t = s->tree[id];
    FloatGrabber g;
    t.serialize(g);
    std::memcpy(out48, g.vals.data(), sizeof(float) * (g.vals.size() < 48 ? g.vals.size() : 48));
}

// KDTreeIntersection::intersect(ray, r, a, b), lib/kdtree.cpp:515-578. ids: miss = 1<<30.
void ref_intersect(void* h, const float* o, const float* d, uint64_t n, uint32_t* ids, float* rst) {
    auto* s = static_cast<RefScene*>(h);
    KDTreeIntersection ti(s->tree);
    for (uint64_t i = 0; i < n; ++i) {
        Ray ray(Point3f(o[3 * i], o[3 * i + 1], o[3 * i + 2]), Vector3f(d[3 * i], d[3 * i + 1], d[3 * i + 2]));
        float r = 0, a = 0, b = 0;
        auto id = ti.intersect(ray, r, a, b);
        if (id) {
            ids[i] = static_cast<uint32_t>(static_cast<KDTree::TriangleId>(id));
            rst[3 * i] = r; rst[3 * i + 1] = a; rst[3 * i + 2] = b;
        } else {
            ids[i] = 1u << 30;
            rst[3 * i] = rst[3 * i + 1] = rst[3 * i + 2] = 0.f;
        }
    }
}

int ref_intersect_ray_box(const float* o, const float* d, const float* box6, float* tmin, float* tmax) {
    Ray ray(Point3f(o[0], o[1], o[2]), Vector3f(d[0], d[1], d[2]));
    Bbox3f box(Point3f(box6[0], box6[1], box6[2]), Point3f(box6[3], box6[4], box6[5]));
    return intersect_ray_box(ray, box, *tmin, *tmax) ? 1 : 0;
}

static Camera make_camera(const RefCamera& rc) {
    aiMatrix4x4 m;
    std::memcpy(&m.a1, rc.trafo4x4, sizeof(float) * 16);
    aiCamera cam;
    cam.mPosition = aiVector3D(0, 0, 0);
    cam.mUp = aiVector3D(0, 1, 0);
    cam.mLookAt = aiVector3D(0, 0, -1);
    cam.mHorizontalFOV = rc.hfov;
    cam.mAspect = rc.aspect;
    return Camera(m, cam);
}

// primary directions exactly as main.cpp:201-209 produces them: dirs[(y*W+x)*pps+i]
void ref_primary_dirs(const RefCamera* rc, int32_t width, int32_t pps, float* cam_pos3, float* dirs,
                      int32_t* height_out) {
    Camera cam = make_camera(*rc);
    int height = width / cam.mAspect;
    *height_out = height;
    cam_pos3[0] = cam.mPosition.x; cam_pos3[1] = cam.mPosition.y; cam_pos3[2] = cam.mPosition.z;
    if (!dirs) return;
    for (int y = 0; y < height; ++y) {
        xorshift64star<float> gen(42);
        for (int x = 0; x < width; ++x) {
            for (int i = 0; i < pps; ++i) {
                float dx = gen();
                float dy = gen();
                auto dir = cam.raster2cam({x + dx, y + dy}, width, height);
                size_t k = (size_t(y) * width + x) * pps + i;
                dirs[3 * k] = dir.x; dirs[3 * k + 1] = dir.y; dirs[3 * k + 2] = dir.z;
            }
        }
    }
}

// The render loop body of main.cpp:181-236 around the reference's trace().
// out_linear_sum : W*H*4, per-pixel SUM over pixel samples of trace() (before /pps, exposure, gamma)
// out_linear_sumsq: W*H*4 or null, per-pixel sum of squares of the per-sample values
// out_final      : W*H*4 or null, after /pps, exposure, gamma (what the P3 writer sees)
int ref_render(void* h, const RefCamera* rc, const RefConfig* cfg, float* out_linear_sum, float* out_linear_sumsq,
               float* out_final, RefStats* stats) {
    auto* s = static_cast<RefScene*>(h);
    const KDTree& tree = s->tree;
    Camera cam = make_camera(*rc);

    Config common;
    common.aspect = rc->aspect;
    common.width = cfg->width;
    common.num_threads = cfg->num_threads;
    common.inverse_gamma = cfg->inverse_gamma;
    common.exposure = cfg->exposure;
    common.bg_color = Color(cfg->bg[0], cfg->bg[1], cfg->bg[2], cfg->bg[3]);
    common.gamma_correction_enabled = cfg->gamma_enabled != 0;
    TracerConfig conf(common);
    conf.max_recursion_depth = cfg->max_depth;
    conf.max_visibility = cfg->max_visibility;
    conf.num_pixel_samples = cfg->pixel_samples;
    conf.num_monte_carlo_samples = cfg->mc_samples;
    conf.shadow_intensity = cfg->shadow_intensity;

    std::vector<Light> lights;
    if (cfg->num_lights == 1) {
        lights.push_back({{cfg->light_pos[0], cfg->light_pos[1], cfg->light_pos[2]},
                          aiColor4D{cfg->light_color[0], cfg->light_color[1], cfg->light_color[2], cfg->light_color[3]}});
    }

    int width = conf.width;
    int height = width / cam.mAspect;
    Image image(width, height);

    Stats::instance().num_rays = 0;
    Stats::instance().num_prim_rays = 0;
    Stats::instance().runtime_ms = 0;
    {
        Runtime rt(Stats::instance().runtime_ms);
        Point3f cam_pos = Point3f(cam.mPosition.x, cam.mPosition.y, cam.mPosition.z);

        // Row-per-task FIFO over num_threads fresh worker threads (what
        // ThreadPool(conf.num_threads) + enqueue-per-row amounts to). Fresh
        // threads matter: sampling.h:12's thread-local stream restarts at seed 4.
        const int row_stride = cfg->row_stride > 0 ? cfg->row_stride : 1;
        std::atomic<int> next_row{cfg->row_begin > 0 ? cfg->row_begin : 0};
        auto worker = [&]() {
            for (;;) {
                int y = next_row.fetch_add(row_stride);
                if (y >= height) break;
                KDTreeIntersection tree_intersection(tree);
                float dx, dy;
                xorshift64star<float> gen(42);
                for (int x = 0; x < width; ++x) {
                    Color sumsq;
                    for (int i = 0; i < conf.num_pixel_samples; ++i) {
                        dx = gen();
                        dy = gen();
                        auto cam_dir = cam.raster2cam({x + dx, y + dy}, width, height);
                        Stats::instance().num_prim_rays += 1;
                        Color c = trace({cam_pos, cam_dir}, tree_intersection, lights, 0, conf);
                        image(x, y) += c;
                        sumsq += c * c;
                    }
                    size_t k = (size_t(y) * width + x) * 4;
                    const Color& sum = image(x, y);
                    out_linear_sum[k] = sum.r; out_linear_sum[k + 1] = sum.g;
                    out_linear_sum[k + 2] = sum.b; out_linear_sum[k + 3] = sum.a;
                    if (out_linear_sumsq) {
                        out_linear_sumsq[k] = sumsq.r; out_linear_sumsq[k + 1] = sumsq.g;
                        out_linear_sumsq[k + 2] = sumsq.b; out_linear_sumsq[k + 3] = sumsq.a;
                    }
                    image(x, y) /= static_cast<float>(conf.num_pixel_samples);
                    image(x, y) = exposure(image(x, y), conf.exposure);
                    if (conf.gamma_correction_enabled) {
                        image(x, y) = gamma(image(x, y), conf.inverse_gamma);
                    }
                    if (out_final) {
                        const Color& f = image(x, y);
                        out_final[k] = f.r; out_final[k + 1] = f.g; out_final[k + 2] = f.b; out_final[k + 3] = f.a;
                    }
                }
            }
        };
        std::vector<std::thread> pool;
        for (int t = 0; t < cfg->num_threads; ++t) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
    }
    stats->num_rays = Stats::instance().num_rays;
    stats->num_prim_rays = Stats::instance().num_prim_rays;
    stats->runtime_ms = Stats::instance().runtime_ms;
    stats->width = width;
    stats->height = height;
    return 0;
}

// operator<<(ostream, Image), lib/raster.h:79-100, on a caller-supplied RGBA float image.
// Returns the number of bytes needed; writes up to cap bytes into buf.
uint64_t ref_write_p3(const float* rgba, int32_t width, int32_t height, char* buf, uint64_t cap) {
    Image img(width, height);
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            const float* p = rgba + (size_t(y) * width + x) * 4;
            img(x, y) = Color(p[0], p[1], p[2], p[3]);
        }
    std::ostringstream os;
    os << img << std::endl; // main.cpp:242
    std::string str = os.str();
    if (buf && cap) std::memcpy(buf, str.data(), str.size() < cap ? str.size() : cap);
    return str.size();
}

void ref_xorshift_float(uint64_t seed, uint64_t n, float* out) {
    xorshift64star<float> gen(seed);
    for (uint64_t i = 0; i < n; ++i) out[i] = gen();
}

void ref_xorshift_u64(uint64_t seed, uint64_t n, uint64_t* out) {
    xorshift64star<uint64_t> gen(seed);
    for (uint64_t i = 0; i < n; ++i) out[i] = gen();
}

// sampling::hemisphere(), lib/sampling.h:20-32, drawn on a fresh thread so the
// thread-local stream (seed 4, sampling.h:12) starts from its first value.
// out: n * 4 floats (x, y, z, cos_theta).
void ref_hemisphere(uint64_t n, float* out) {
    std::thread t([&]() {
        for (uint64_t i = 0; i < n; ++i) {
            auto s = sampling::hemisphere();
            out[4 * i] = s.first.x; out[4 * i + 1] = s.first.y; out[4 * i + 2] = s.first.z; out[4 * i + 3] = s.second;
        }
    });
    t.join();
}

void ref_tonemap(const float* rgba_in, uint64_t npix, float exposure_v, int gamma_enabled, float inverse_gamma,
                 float* rgba_out) {
    for (uint64_t i = 0; i < npix; ++i) {
        Color c(rgba_in[4 * i], rgba_in[4 * i + 1], rgba_in[4 * i + 2], rgba_in[4 * i + 3]);
        c = exposure(c, exposure_v);
        if (gamma_enabled) c = gamma(c, inverse_gamma);
        rgba_out[4 * i] = c.r; rgba_out[4 * i + 1] = c.g; rgba_out[4 * i + 2] = c.b; rgba_out[4 * i + 3] = c.a;
    }
}

} // extern "C"
