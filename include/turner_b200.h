/* turner_b200 -- C ABI of the B200-native path-tracing hot path.
 *
 * Drop-in boundary: the body of the reference's render loop, main.cpp:181-236
 * (inputs: triangles, camera, lights, TracerConfig; output: the per-pixel RGBA
 * image + Stats). The reference's own plug point is the link-time symbol
 *   Color trace(const Ray&, KDTreeIntersection&, const std::vector<Light>&, int, const TracerConfig&)
 * (trace.h:23-25, called from main.cpp:212-214 one ray at a time); a wavefront
 * GPU renderer cannot live behind a per-ray call, so the boundary sits one level
 * up. INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * All entry points: plain pointers and sizes, no C++/torch types, return 0 on
 * success or a negative trn_status; trn_last_error() gives the message of the
 * last failure on the calling thread. Nothing here aborts or throws. There is NO
 * CPU fallback: compute entry points fail with TRN_ERR_CUDA when no sm_100 device
 * is usable.
 *
 * Threads: a trn_scene may be used from several threads; compute calls on the
 * same scene AND device serialise on a per-device lock (one set of wave buffers
 * per device), calls on different devices run concurrently.
 *
 * Citations are relative to the reference repository root.
 */
#ifndef TURNER_B200_H
#define TURNER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRN_MISS_ID 0x40000000u /* OptionalId miss value, lib/kdtree.h:156-161 */

typedef enum trn_status {
    TRN_OK = 0,
    TRN_ERR_INVALID = -1, /* bad argument (the reference would assert, config.h:29-33,119-126) */
    TRN_ERR_CUDA = -2,    /* CUDA runtime / no usable device */
    TRN_ERR_NCCL = -3,    /* NCCL failure or libnccl not loadable (multi-GPU only) */
    TRN_ERR_IO = -4,      /* scene file unreadable / unsupported (.blend loader) */
    TRN_ERR_LIMIT = -5    /* workload exceeds an implementation limit (stated in the message) */
} trn_status;

typedef struct trn_scene trn_scene; /* opaque: triangles + kd-tree, host copy + per-device copies */

/* Camera as the render loop uses it: lib/types.h:92-123 (Camera ctor + raster2cam). */
typedef struct trn_camera {
    float pos[3];  /* mPosition after `trafo * mPosition` (types.h:101) */
    float rot[9];  /* aiMatrix3x3 trafo_, row-major a1..c3 (types.h:141) */
    float delta_x; /* tan(mHorizontalFOV) (types.h:103) */
    float delta_y; /* delta_x / mAspect   (types.h:104) */
} trn_camera;

/* Light, lib/types.h:81-84. */
typedef struct trn_light {
    float pos[3];
    float rgba[4];
} trn_light;

enum {
    TRN_PATHTRACER = 0, /* pathtracer.cpp:14-102 */
    TRN_RAYCASTER = 1,  /* raycaster.cpp:7-24 */
    TRN_RAYTRACER = 2   /* raytracer.cpp:6-67 (Whitted: Lambert + mirror recursion + shadow attenuation; needs one light) */
};

/* TracerConfig (config.h:103-153) + image size + the sample split. */
typedef struct trn_render_config {
    int32_t width;         /* --width (config.h:16) */
    int32_t height;        /* width / aspect (main.cpp:178-179) */
    int32_t max_depth;     /* --max-depth > 0 (config.h:107,121) */
    int32_t mc_samples;    /* --monte-carlo-samples >= 1 here (config.h:117; 0 divides by zero in the reference) */
    int32_t pixel_samples; /* --pixel-samples >= 1 (config.h:116,124) */
    int32_t integrator;    /* TRN_PATHTRACER | TRN_RAYCASTER */
    float bg_rgba[4];      /* --background, alpha 1 (config.h:45-60) */
    float max_visibility;  /* raycaster only (config.h:110) */
    int32_t num_lights;    /* 0 or 1 (main.cpp:123) */
    trn_light light;
    uint64_t seed;         /* run seed of the counter-seeded hemisphere streams (DESIGN.md "RNG") */
    /* Sample split (multi-GPU): this call renders pixel samples
     * i = sample_begin, sample_begin + sample_stride, ... < pixel_samples. 0/1 = all. */
    int32_t sample_begin;
    int32_t sample_stride;
    float shadow_intensity; /* raytracer only: --shadow in [0,1] (config.h:113,123) */
} trn_render_config;

typedef struct trn_stats {
    uint64_t rays;        /* Stats::num_rays: trace() calls that pass the depth check (pathtracer.cpp:17-21) */
    uint64_t prim_rays;   /* Stats::num_prim_rays (main.cpp:211) */
    uint64_t shadow_rays; /* shadow queries (pathtracer.cpp:49-50), not counted by the reference */
    uint64_t launches;    /* kernels launched by this call */
    double ms_render;     /* device time of the call (CUDA events), ms; host-buffer variants include the D2H copy */
    double ms_trace;      /* summed device time of the closest-hit traversal kernel (only when profiling is on) */
    double ms_shadow;     /* ... of the any-hit traversal kernel */
    double ms_shade;      /* ... of the shade/bounce kernel */
    double ms_other;      /* ... raygen + bookkeeping */
    uint64_t trace_launches;  /* closest-hit traversal launches / rays they processed */
    uint64_t trace_queries;
    uint64_t shadow_launches; /* any-hit traversal launches (queries = shadow_rays) */
    /* Visit counts of the two traversal kernels, filled only while trn_set_counting(1) is on (instrumented
     * kernels, slower): inner-node visits, reference leaf nodes (8-byte id pairs, lib/kdtree.h:62-154) and
     * triangle tests -- the n_* of the algorithmic-bytes formula in DESIGN.md "Roofline". */
    uint64_t trace_inner, trace_leaf_nodes, trace_tri_tests;
    uint64_t shadow_inner, shadow_leaf_nodes, shadow_tri_tests;
    /* ... and what the production kernels really visit on the device layout (sibling pairs + empty-space cuts) */
    uint64_t trace_actual_inner, trace_actual_leaf_nodes, trace_actual_tri_tests;
    uint64_t shadow_actual_inner, shadow_actual_leaf_nodes, shadow_actual_tri_tests;
    /* host-buffer / multi-GPU variants: device time of the ncclReduce(sum) of the accumulation buffers and of the
     * device->host copy of the image; both are part of ms_render */
    double ms_reduce;
    double ms_d2h;
    /* counting mode, pooled traversal kernel (trees with >= 1024 leaves): what the production schedule itself issues --
     * [0] walk steps (one 16-byte node-pair load each), [1] chunks (four 16-byte plane records = one 64-byte block; + one 16-byte id vector if a triangle survives the pre-filter),
     * [2] triangle pre-tests (valid ids of the chunks), [3] exact tests (32-byte hot record), [4] of those with the
     * 32-byte cold record, [5] stack pushes, [6] pops (16 bytes of local memory each), [7] leaves, [8] walk steps at
     * empty-space cuts (a subset of [0]), [9] unused */
    uint64_t trace_pooled[10];
    uint64_t shadow_pooled[10];
    /* one-leaf scenes (the brute-force kernel, traverse_flat.cuh): scan records every query of this call was tested against
     * -- a record is one triangle or one coplanar pair; 0 when another traversal kernel ran */
    uint64_t flat_records;
} trn_stats;

typedef struct trn_scene_info {
    uint64_t num_triangles; /* Stats::num_triangles */
    uint64_t num_nodes;     /* KDTree::num_nodes(), lib/kdtree.h:219 */
    uint64_t kdtree_height; /* KDTree::height(), lib/kdtree.h:197-218 */
    uint64_t num_leaf_refs; /* triangle references in leaves */
    uint64_t num_cut_nodes; /* empty-space cuts kept in the device layout (dropped by lib/kdtree.cpp:168-172) */
    float box[6];           /* KDTree::box(): min xyz, max xyz */
    double build_ms;        /* kd-tree build time (host builder, or device builder incl. its copies) */
    double upload_ms;       /* layout + H2D */
} trn_scene_info;

const char* trn_last_error(void);
/* number of usable CUDA devices (0 when there is none; never fails) */
int32_t trn_device_count(void);

/* ---- scene: replaces `KDTree tree(triangles_from_scene(scene))`, main.cpp:25-82,156-157.
 * verts/normals: n*9 floats (v0 v1 v2 / n0 n1 n2, world space), diffuse: n*4 rgba. Copies its inputs.
 * Builds the kd-tree on the host (same SAH build as lib/kdtree.cpp:124-467, node-for-node) and keeps a host
 * copy; device copies are created lazily per device on first use. */
int32_t trn_scene_create(const float* verts, const float* normals, const float* diffuse, uint32_t n, trn_scene** out);
/* same plus the mirror material the raytracer integrator reads (main.cpp:44-47): reflective n*4 rgba
 * (AI_MATKEY_COLOR_REFLECTIVE) and reflectivity n (AI_MATKEY_REFLECTIVITY); either may be NULL (= 0) */
int32_t trn_scene_create_ex(const float* verts, const float* normals, const float* diffuse, const float* reflective,
                            const float* reflectivity, uint32_t n, trn_scene** out);
/* Same, but the kd-tree is built ON THE DEVICE (lib/kdtree.cpp:124-467 replaced by a level-synchronous binned-SAH build:
 * same cost function and termination rules, 32 candidate planes per axis instead of the reference's event sweep, triangles
 * clipped at every split). The tree has a different shape than the reference's -- closest-hit ids do not depend on the
 * shape except for which of several triangles reports an EXACT tie in r -- and is ready ~20x sooner at 1 M triangles.
 * device < 0: the current device. trn_scene_get_nodes / trn_scene_save_cache derive the reference's FlatNode view from it
 * on first use. Setting TRN_BUILDER=gpu makes trn_scene_create(_ex) take this path as well. */
int32_t trn_scene_create_gpu(const float* verts, const float* normals, const float* diffuse, const float* reflective,
                             const float* reflectivity, uint32_t n, int32_t device, trn_scene** out);
void trn_scene_destroy(trn_scene* scene);
int32_t trn_scene_get_info(const trn_scene* scene, trn_scene_info* info);
/* the flattened tree in the reference's FlatNode encoding (lib/kdtree.h:62-154), num_nodes uint64 values */
int32_t trn_scene_get_nodes(const trn_scene* scene, uint64_t* out_nodes);

/* ---- closest hit for arbitrary rays: replaces KDTreeIntersection::intersect(ray, r, a, b),
 * lib/kdtree.cpp:515-578. Host buffers: origins/dirs n*3 floats in; ids n (TRN_MISS_ID on miss), rst n*3 out. */
int32_t trn_intersect(trn_scene* scene, int32_t device, const float* origins, const float* dirs, uint64_t n,
                      uint32_t* ids, float* rst);

/* ---- primary-hit parity hook (BASELINE config 2): the closest hit of every primary ray the render loop would
 * shoot (main.cpp:201-214), index (y*width + x)*pixel_samples + i. ids / rst (3 per ray) are host buffers. */
int32_t trn_primary_hits(trn_scene* scene, int32_t device, const trn_camera* cam, const trn_render_config* cfg,
                         uint32_t* ids, float* rst);

/* ---- render: replaces the loop main.cpp:187-236 up to (not including) `/= pps`, exposure and gamma.
 * out_rgba_sum: width*height*4 floats, HOST memory, per-pixel SUM over the rendered pixel samples of trace().
 * Blocking; includes the device->host copy. */
int32_t trn_render(trn_scene* scene, int32_t device, const trn_camera* cam, const trn_render_config* cfg,
                   float* out_rgba_sum, trn_stats* stats);

/* Same, but ADDS into a caller-owned DEVICE buffer (width*height*4 floats on `device`), launching on the given CUDA
 * stream (cudaStream_t passed as void*; NULL = legacy default stream). For callers that own device memory / streams /
 * a process-per-GPU reduce (torch.distributed). The host drives the wavefront depth by depth (it reads each wave's size
 * back), so the call returns when the last kernels are enqueued; it waits for them only if `stats` is non-NULL
 * (ms_render is then measured with events on that stream). */
int32_t trn_render_device(trn_scene* scene, int32_t device, const trn_camera* cam, const trn_render_config* cfg,
                          float* d_accum_rgba, void* cuda_stream, trn_stats* stats);

/* Single-process multi-GPU render (sample split over `num_devices` GPUs, scene replicated, one ncclReduce(sum)
 * of the float accumulation buffers onto devices[0]). libnccl is dlopen()ed on first use. */
int32_t trn_render_multi(trn_scene* scene, const int32_t* devices, int32_t num_devices, const trn_camera* cam,
                         const trn_render_config* cfg, float* out_rgba_sum, trn_stats* stats);

/* ---- process-per-GPU jobs (one rank per process: torchrun, mpirun ...). The image reduce is this library's own
 * ncclReduce; the launcher only has to carry the 128-byte id from rank 0 to the other ranks.
 *   trn_comm_unique_id  rank 0: ncclGetUniqueId
 *   trn_comm_init_rank  every rank: ncclCommInitRank on `device` (nranks == 1: no NCCL involved, id128 is ignored)
 *   trn_render_rank     every rank: renders its share of the pixel samples (i = sample_begin + sample_stride * (rank +
 *                       nranks * k), scene replicated), then ONE ncclReduce(sum) onto rank 0; rank 0 copies the summed
 *                       image to out_rgba_sum (host, width*height*4 floats; NULL = leave it on the device, ignored on the
 *                       other ranks). stats are this rank's (rays of its share; ms_reduce / ms_d2h filled). Blocking. */
typedef struct trn_comm trn_comm;
int32_t trn_comm_unique_id(uint8_t* id128);
int32_t trn_comm_init_rank(const uint8_t* id128, int32_t nranks, int32_t rank, int32_t device, trn_comm** out);
void trn_comm_destroy(trn_comm* comm);
int32_t trn_render_rank(trn_scene* scene, trn_comm* comm, const trn_camera* cam, const trn_render_config* cfg,
                        float* out_rgba_sum, trn_stats* stats);

/* ---- asynchronous frames: trn_render_async returns at once; trn_wait blocks until out_rgba_sum (host, ideally
 * pinned) holds the frame, fills stats and frees the job. Frames of one scene+device render one after the other, but the
 * device->host copy of frame k runs next to the render of frame k+1 (two accumulation buffers, a copy stream). Every job
 * must be waited for (trn_wait releases it) before its scene is destroyed. */
typedef struct trn_job trn_job;
int32_t trn_render_async(trn_scene* scene, int32_t device, const trn_camera* cam, const trn_render_config* cfg,
                         float* out_rgba_sum, trn_job** job);
int32_t trn_wait(trn_job* job, trn_stats* stats);

/* ---- occlusion parity hook: the shadow predicate of pathtracer.cpp:49-53 for arbitrary rays --
 * occluded[i] = 1 iff some triangle is accepted by intersect_ray_triangle (lib/intersection.h:63-89) with
 * 0 <= r <= tmax[i], i.e. NOT (!hit || r_closest > dist_to_light). Runs the production any-hit kernels. */
int32_t trn_occluded(trn_scene* scene, int32_t device, const float* origins, const float* dirs, const float* tmax, uint64_t n,
                     uint8_t* occluded);

/* ---- measured ceilings for the roofline of the traversal kernels: GB/s of independent random 16-byte gathers (the
 * kernels' access shape) over a working set of set_bytes; mode 0 = long run for a set that fits L1, 1 = L2 / HBM sets */
int32_t trn_measure_gather_peak(int32_t device, uint64_t set_bytes, int32_t mode, double* gbps);

/* per-kernel timing inside trn_render* (adds an event pair per launch); off by default */
void trn_set_profiling(int32_t enabled);
/* run the instrumented traversal kernels and fill the visit counts of trn_stats; off by default */
void trn_set_counting(int32_t enabled);
/* like trn_intersect, additionally returns in counts6 the batch's {inner visits, reference leaf nodes, triangle
 * tests} of the reference-shaped schedule (the B_alg definition) followed by the same three for the device layout */
int32_t trn_intersect_counted(trn_scene* scene, int32_t device, const float* origins, const float* dirs, uint64_t n,
                              uint32_t* ids, float* rst, uint64_t* counts6);

/* ---- host-side pieces of the reference's main() that the CLI keeps (no GPU involved) ------------------- */

/* Camera(trafo, aiCamera) + height, lib/types.h:92-105, main.cpp:109-119,178-179.
 * trafo4x4: the camera node's transformation, row-major a1..d4. */
int32_t trn_camera_setup(const float* trafo4x4, float hfov, float aspect, int32_t width, trn_camera* cam,
                         int32_t* height);
/* image(x,y) /= pps; exposure; gamma -- main.cpp:216-223, lib/effects.h:15-48. in/out: npix*4 floats. */
int32_t trn_tonemap(const float* rgba_sum, uint64_t npix, int32_t pixel_samples, float exposure,
                    int32_t gamma_enabled, float inverse_gamma, float* rgba_out);
/* operator<<(ostream, Image) + std::endl, lib/raster.h:79-100, main.cpp:242. Returns the byte count needed;
 * writes at most cap bytes. */
uint64_t trn_write_p3(const float* rgba, int32_t width, int32_t height, char* buf, uint64_t cap);

/* .blend scene file -> triangle arrays + camera + light (replaces Assimp::Importer::ReadFile + triangles_from_scene,
 * main.cpp:25-82,96-136). Arrays are malloc()ed; release with trn_loaded_scene_free. */
typedef struct trn_loaded_scene {
    uint32_t num_triangles;
    float* verts;   /* n*9 */
    float* normals; /* n*9 */
    float* diffuse; /* n*4 */
    float* reflective;   /* n*4 */
    float* reflectivity; /* n */
    int32_t has_camera;
    float cam_trafo4x4[16];
    float cam_hfov;
    float cam_aspect; /* 0 when the file does not fix one (Blender importer leaves mAspect = 0) */
    int32_t num_lights;
    trn_light light;
} trn_loaded_scene;
int32_t trn_load_blend(const char* path, trn_loaded_scene* out);
/* neutral triangle-soup text file (format in turner_b200/csrc/blend_loader.cpp) -> same structure */
int32_t trn_load_soup(const char* path, trn_loaded_scene* out);
void trn_loaded_scene_free(trn_loaded_scene* s);

/* ---- kdtree.cache (main.cpp:142-167): the reference stores its KDTree (triangles with their precomputed fields, the
 * scene box, the FlatNode array) through cereal's PortableBinary archive and, when ./kdtree.cache exists, loads it INSTEAD
 * of the scene's triangles -- whatever scene was asked for. Here the file is explicit and checked:
 *   trn_scene_save_cache  writes the same byte layout (1 flag byte; u64 triangle count; 48 f32 per triangle in the order
 *                         of Triangle::serialize, lib/triangle.h:89-92 -- ambient is written as 0, it is not part of the
 *                         scene arrays and no tracer reads it, emissive = diffuse as main.cpp:43 loads it; 6 f32 box;
 *                         u64 node count; 8-byte FlatNodes, lib/kdtree.h:62-154).
 *   trn_scene_load_cache  reads such a file (written by this library or by the reference), rebuilds the tree from the
 *                         cached triangles and REFUSES the file (TRN_ERR_INVALID) when the cached node array or box is not
 *                         what these triangles produce: a stale or foreign cache cannot be rendered silently. */
int32_t trn_scene_save_cache(const trn_scene* scene, const char* path);
int32_t trn_scene_load_cache(const char* path, trn_scene** out);

#ifdef __cplusplus
}
#endif
#endif /* TURNER_B200_H */
