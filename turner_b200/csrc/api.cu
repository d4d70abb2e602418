// C ABI (include/turner_b200.h) + the host side of the wavefront: device scene
// layout, wave buffers, the depth-by-depth launch schedule, sample-split
// multi-GPU. No CPU compute path exists here: every compute entry point needs a
// CUDA device and fails with TRN_ERR_CUDA otherwise.
#include "../../include/turner_b200.h"

#include "host_util.h"
#include "kdtree_build.h"
#include "kdtree_build_gpu.h"
#include "kernels.cuh"
#include "traverse_persistent.cuh"
#include "traverse_pooled.cuh"
#include "traverse_flat.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

// launch trace_persistent_ww_kernel<MODE, two_pass>
#define TRN_LAUNCH_WW(MODE, two_pass, grid, stream, ...)                                              \
    do {                                                                                              \
        if (two_pass) trace_persistent_ww_kernel<MODE, true><<<(grid), 128, 0, (stream)>>>(__VA_ARGS__);  \
        else trace_persistent_ww_kernel<MODE, false><<<(grid), 128, 0, (stream)>>>(__VA_ARGS__);          \
    } while (0)

namespace trn {

constexpr int kVisitSlots = 32; // 12 of the per-ray twins (reference-shaped + device layout) + 2 x 10 PooledCounts
thread_local std::string g_last_error;
static std::atomic<int> g_profiling{0};
static std::atomic<int> g_counting{0};

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                                          \
    do {                                                                                                        \
        cudaError_t e__ = (expr);                                                                               \
        if (e__ != cudaSuccess)                                                                                 \
            return fail(TRN_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__) + " (" __FILE__ ":" + \
                                          std::to_string(__LINE__) + ")");                                      \
    } while (0)

// extern "C" bodies: nothing may unwind through the C ABI (std::bad_alloc on a huge scene, a failed std::thread)
#define TRN_GUARD_BEGIN try {
#define TRN_GUARD_END                                                                            \
    }                                                                                            \
    catch (const std::bad_alloc&) { return fail(TRN_ERR_LIMIT, "out of host memory"); }          \
    catch (const std::exception& e) { return fail(TRN_ERR_LIMIT, std::string("host error: ") + e.what()); }

// scoped CUDA events (destroyed on every return path)
template <int N> struct Events {
    cudaEvent_t e[N] = {};
    ~Events() {
        for (cudaEvent_t x : e)
            if (x) cudaEventDestroy(x);
    }
    cudaError_t create() {
        for (cudaEvent_t& x : e) {
            cudaError_t rc = cudaEventCreate(&x);
            if (rc != cudaSuccess) return rc;
        }
        return cudaSuccess;
    }
};

// scoped device allocation (freed on every return path)
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

// ------------------------------------------------------------ per-device state
struct DeviceScene {
    int device = -1;
    DevScene dev{};
    void* d_nodes = nullptr;
    void* d_refs = nullptr;
    void* d_pnodes = nullptr;
    void* d_prefs = nullptr;
    void* d_isect_hot = nullptr;
    void* d_isect_cold = nullptr;
    void* d_tri_box = nullptr;
    void* d_planes = nullptr; // float4 per leaf REFERENCE: unit normal, n.v0 of its triangle -- the pre-filter record of the pooled kernels
    void* d_shade = nullptr;
    void* d_mirror = nullptr;
    uint32_t* d_child_slot = nullptr; // raytracer: child-ray slot of every shadow query
    // work buffers (sized lazily, reused across calls)
    uint64_t wave_cap = 0;        // rays per wave buffer
    std::vector<RayWave> waves;   // one per depth level in use
    std::vector<void*> wave_mem;  // backing allocations of `waves`
    ShadowWave shadow{};
    ShadowWave shadow_alt{};        // second shadow wave: shadow waves run on stream_b next to the following closest-hit wave
    void* shadow_mem = nullptr;
    cudaStream_t stream_b = nullptr;
    cudaEvent_t ev_shaded = nullptr, ev_shadow_done[2] = {nullptr, nullptr};
    bool shadow_pending[2] = {false, false};
    uint4* d_hits = nullptr;
    uint32_t *d_keys = nullptr, *d_keys_alt = nullptr, *d_order = nullptr, *d_order_alt = nullptr; // ray sorting
    void* d_sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    WaveCounters* d_counters = nullptr;  // ring of counters, one per (launch) use
    WaveCounters* h_counters = nullptr;  // pinned mirror
    uint32_t counter_slots = 0;
    uint32_t ring_pos = 0;
    uint32_t treelet_pairs = 0; // node pairs of the top treelet present in pnodes (<= kTreeletNodes / 2)
    bool two_pass = true; // leaf evaluation schedule of the persistent kernels (small leaves: two-pass)
    bool pooled = true;   // pooled kernel for both kinds of wave (a real tree: many leaves of moderate size)
    bool flat = false;    // the tree is ONE leaf of <= kFlatMaxTris triangles (cornell_box): brute-force kernel, no walk
    uint32_t flat_first = 0, flat_count = 0; // that leaf's references
    uint32_t flat_groups = 0;                // scan records of the brute-force kernel (a triangle or a coplanar pair each)
    FlatParams flat_params{};                // its pre-filter records (plane + grown box, visiting order): kernel parameters
    int grid_flat[3] = {0, 0, 0};
    int grid_closest = 0, grid_shadow = 0, grid_plain = 0; // persistent grids: resident CTAs per SM x SMs
    int grid_pooled[3] = {0, 0, 0};                        // same for trace_pooled_kernel<MODE>
    unsigned long long* d_hitcount = nullptr;
    unsigned long long* d_visits = nullptr; // [6]: closest inner/leaf/tri, shadow inner/leaf/tri
    float2* d_jitter = nullptr;
    size_t jitter_elems = 0;
    int jitter_w = 0, jitter_pps = 0;
    float4* d_accum = nullptr; // internal accumulation buffer for host-buffer renders
    size_t accum_pixels = 0;
    float4* d_async[2] = {nullptr, nullptr}; // trn_render_async: two frames in flight (render k+1 next to the D2H of k)
    size_t async_pixels[2] = {0, 0};
    cudaEvent_t ev_async[2] = {nullptr, nullptr};
    uint32_t async_use = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_copy = nullptr; // D2H of a finished frame next to the following render (trn_render_async)
    cudaEvent_t ev_sync = nullptr;
    double upload_ms = 0;
    // Every compute entry point holds this for its whole duration: the wave buffers, counters ring, shadow buffers and
    // streams above are one set of work state per device (include/turner_b200.h: calls on one scene+device serialise).
    std::recursive_mutex mu;

    DeviceScene() = default;
    DeviceScene(const DeviceScene&) = delete;
    DeviceScene& operator=(const DeviceScene&) = delete;
    ~DeviceScene() { release(); }

    void release() {
        if (device < 0) return;
        cudaSetDevice(device);
        cudaFree(d_nodes);
        cudaFree(d_refs);
        cudaFree(d_pnodes);
        cudaFree(d_prefs);
        cudaFree(d_isect_hot);
        cudaFree(d_isect_cold);
        cudaFree(d_tri_box);
        cudaFree(d_planes);
        cudaFree(d_shade);
        cudaFree(d_mirror);
        cudaFree(d_child_slot);
        for (void* p : wave_mem) cudaFree(p);
        cudaFree(shadow_mem);
        cudaFree(d_hits);
        cudaFree(d_keys);
        cudaFree(d_keys_alt);
        cudaFree(d_order);
        cudaFree(d_order_alt);
        cudaFree(d_sort_tmp);
        cudaFree(d_counters);
        if (h_counters) cudaFreeHost(h_counters);
        cudaFree(d_hitcount);
        cudaFree(d_visits);
        cudaFree(d_jitter);
        cudaFree(d_accum);
        for (int k = 0; k < 2; ++k) {
            cudaFree(d_async[k]);
            if (ev_async[k]) cudaEventDestroy(ev_async[k]);
        }
        if (ev_sync) cudaEventDestroy(ev_sync);
        if (ev_shaded) cudaEventDestroy(ev_shaded);
        for (cudaEvent_t e : ev_shadow_done)
            if (e) cudaEventDestroy(e);
        if (stream_b) cudaStreamDestroy(stream_b);
        if (stream_copy) cudaStreamDestroy(stream_copy);
        if (stream) cudaStreamDestroy(stream);
        device = -1;
    }
};

} // namespace trn

struct trn_scene {
    trn::HostTriangles tris;
    trn::KdTree tree;
    // device layout, built once on the host
    std::vector<uint2> gpu_nodes;
    std::vector<uint32_t> leaf_refs;
    bool reference_shape = false; // tree.nodes / gpu_nodes / leaf_refs are filled (always after a host build)
    bool gpu_built = false;
    std::map<int, std::unique_ptr<trn::DeviceScene>> devices;
    std::mutex mu;
    // NCCL communicators of the last device set used by trn_render_multi (ncclCommInitAll costs seconds)
    std::vector<int> nccl_devs;
    std::vector<void*> nccl_comms;
    int (*nccl_destroy)(void*) = nullptr;
};

// one rank of a process-per-GPU job (trn_comm_init_rank)
struct trn_comm {
    void* comm = nullptr; // ncclComm_t
    int nranks = 1, rank = 0, device = 0;
};

namespace trn {

// Rewrite the reference-format node array for the GPU: same indexing (so `right` links stay
// valid), but the head of every leaf run becomes (first ref offset, count<<2 | 3) into a flat
// triangle-reference list, so a leaf costs one 8-byte node load + `count` 4-byte id loads.
static void make_gpu_layout(trn_scene& sc) {
    const auto& nodes = sc.tree.nodes;
    sc.gpu_nodes.resize(nodes.size());
    for (size_t i = 0; i < nodes.size(); ++i)
        sc.gpu_nodes[i] = make_uint2(static_cast<uint32_t>(nodes[i] >> 32), static_cast<uint32_t>(nodes[i]));
    sc.leaf_refs.clear();
    sc.leaf_refs.reserve(sc.tree.num_leaf_refs);
    std::vector<uint32_t> stack;
    stack.push_back(0);
    while (!stack.empty()) {
        uint32_t idx = stack.back();
        stack.pop_back();
        uint64_t n = nodes[idx];
        if ((n & 3) != 3) {
            stack.push_back(static_cast<uint32_t>(n) >> 2);
            stack.push_back(idx + 1);
            continue;
        }
        const uint32_t first = static_cast<uint32_t>(sc.leaf_refs.size());
        uint32_t count = 0;
        for (uint32_t k = idx; (nodes[k] & 3) == 3; ++k) { // lib/kdtree.cpp:599-605
            sc.leaf_refs.push_back(static_cast<uint32_t>(nodes[k] >> 32));
            ++count;
            if (static_cast<uint32_t>(nodes[k]) == 0xFFFFFFFFu) break;
            sc.leaf_refs.push_back(static_cast<uint32_t>(nodes[k] & 0xFFFFFFFFu) >> 2);
            ++count;
        }
        sc.gpu_nodes[idx] = make_uint2(first, (count << 2) | 3u);
    }
}

// the reference's view of the tree (FlatNode array, leaf runs): the host builder makes it itself, a device-built scene
// derives it from the pair layout the first time somebody asks (trn_scene_get_nodes, kdtree.cache, the counting twins)
static void ensure_reference_shape(trn_scene* sc) {
    std::lock_guard<std::mutex> lock(sc->mu);
    if (sc->reference_shape) return;
    reference_shape_from_pairs(sc->tree);
    make_gpu_layout(*sc);
    sc->reference_shape = true;
}

static uint64_t env_u64(const char* name, uint64_t dflt);

static int get_device_scene(trn_scene* sc, int device, DeviceScene** out) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(TRN_ERR_CUDA, "no CUDA device available (turner_b200 has no CPU fallback)");
    if (device < 0) CUDA_TRY(cudaGetDevice(&device));
    if (device >= ndev) return fail(TRN_ERR_INVALID, "device ordinal out of range");
    std::lock_guard<std::mutex> lock(sc->mu);
    auto it = sc->devices.find(device);
    if (it != sc->devices.end()) {
        *out = it->second.get();
        return TRN_OK;
    }
    auto t0 = std::chrono::steady_clock::now();
    CUDA_TRY(cudaSetDevice(device));
    std::unique_ptr<DeviceScene> ds(new DeviceScene);
    ds->device = device;
    auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, std::max<size_t>(bytes, 16));
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    if (sc->reference_shape) { // only the instrumented reference-schedule twins read these (upload_reference_shape otherwise)
        CUDA_TRY(up(&ds->d_nodes, sc->gpu_nodes.data(), sc->gpu_nodes.size() * sizeof(uint2)));
        CUDA_TRY(up(&ds->d_refs, sc->leaf_refs.data(), sc->leaf_refs.size() * sizeof(uint32_t)));
    }
    CUDA_TRY(up(&ds->d_pnodes, sc->tree.pair_nodes.data(), sc->tree.pair_nodes.size() * sizeof(uint64_t)));
    CUDA_TRY(up(&ds->d_prefs, sc->tree.pair_leaf_refs.data(), sc->tree.pair_leaf_refs.size() * sizeof(uint32_t)));
    std::vector<float> flat_planes; // host copy for the one-leaf kernel's parameter block
    {
        const size_t nt = sc->tris.count;
        std::vector<float> hot(nt * 8), cold(nt * 8);
        for (size_t i = 0; i < nt; ++i) {
            std::memcpy(&hot[i * 8], &sc->tris.isect[i * 16], 8 * sizeof(float));
            std::memcpy(&cold[i * 8], &sc->tris.isect[i * 16 + 8], 8 * sizeof(float));
        }
        CUDA_TRY(up(&ds->d_isect_hot, hot.data(), hot.size() * sizeof(float)));
        CUDA_TRY(up(&ds->d_isect_cold, cold.data(), cold.size() * sizeof(float)));
        float scale = 0.f;
        for (int c = 0; c < 6; ++c) scale = std::max(scale, std::fabs(sc->tree.box[c]));
        const float grow = 1e-4f * scale;
        std::vector<float> tb(nt * 8, 0.f);
        for (size_t i = 0; i < nt; ++i) {
            const float* v = &sc->tris.verts[i * 9];
            for (int c = 0; c < 3; ++c) {
                tb[i * 8 + c] = std::min(v[c], std::min(v[3 + c], v[6 + c])) - grow;
                tb[i * 8 + 4 + c] = std::max(v[c], std::max(v[3 + c], v[6 + c])) + grow;
            }
        }
        CUDA_TRY(up(&ds->d_tri_box, tb.data(), tb.size() * sizeof(float)));
        // plane records of the pooled kernels' pre-filter: (n, dp), dp = n.v0 accumulated in double and rounded once
        std::vector<float> pl(nt * 4);
        for (size_t i = 0; i < nt; ++i) {
            const float* q = &sc->tris.isect[i * 16]; // v0.xyz n.xyz ...
            pl[i * 4] = q[3];
            pl[i * 4 + 1] = q[4];
            pl[i * 4 + 2] = q[5];
            pl[i * 4 + 3] = static_cast<float>(static_cast<double>(q[3]) * q[0] + static_cast<double>(q[4]) * q[1] + static_cast<double>(q[5]) * q[2]);
        }
#if TRN_PQ_REFPLANES
        { // one plane record per leaf REFERENCE, in the order of the (padded) reference array: the pooled kernel reads a chunk's
          // four planes as one 64-byte block and the ids only for pre-filter survivors (16 B x 3.06 M references = 49 MB on the
          // benchmark's device-built tree; +4 % whole-job throughput over the per-triangle array, profiles/README.md)
            const auto& refs = sc->tree.pair_leaf_refs;
            std::vector<float> pr(refs.size() * 4);
            for (size_t j = 0; j < refs.size(); ++j) std::memcpy(&pr[j * 4], &pl[static_cast<size_t>(refs[j]) * 4], 16);
            CUDA_TRY(up(&ds->d_planes, pr.data(), pr.size() * sizeof(float)));
        }
#else
        CUDA_TRY(up(&ds->d_planes, pl.data(), pl.size() * sizeof(float)));
#endif
        if (nt <= static_cast<size_t>(kFlatMaxTris)) flat_planes = pl;
    }
    CUDA_TRY(up(&ds->d_shade, sc->tris.shade.data(), sc->tris.shade.size() * sizeof(float)));
    CUDA_TRY(up(&ds->d_mirror, sc->tris.mirror.data(), sc->tris.mirror.size() * sizeof(float)));
    ds->dev.nodes = static_cast<const uint2*>(ds->d_nodes);
    ds->dev.leaf_refs = static_cast<const uint32_t*>(ds->d_refs);
    ds->dev.pnodes = static_cast<const uint2*>(ds->d_pnodes);
    ds->dev.prefs = static_cast<const uint32_t*>(ds->d_prefs);
    ds->dev.isect_hot = static_cast<const float4*>(ds->d_isect_hot);
    ds->dev.isect_cold = static_cast<const float4*>(ds->d_isect_cold);
    ds->dev.tri_box = static_cast<const float4*>(ds->d_tri_box);
    ds->dev.shade = static_cast<const float4*>(ds->d_shade);
    ds->dev.mirror = static_cast<const float4*>(ds->d_mirror);
    ds->dev.treelet_pairs = static_cast<uint32_t>(std::min<size_t>(sc->tree.pair_nodes.size(), KdTree::kTreeletNodes) / 2);
    float extent = 0.f;
    for (int c = 0; c < 3; ++c) {
        ds->dev.lo[c] = sc->tree.box[c];
        ds->dev.hi[c] = sc->tree.box[3 + c];
        extent = std::max(extent, sc->tree.box[3 + c] - sc->tree.box[c]);
    }
    ds->dev.verbatim = extent < 0.1f ? 1u : 0u; // see DevScene::verbatim
    {
        // average triangles per non-empty leaf decides the leaf schedule (a one-leaf scene like cornell_box: one pass)
        uint64_t leaves = 0;
        for (uint64_t nd : sc->tree.pair_nodes) {
            const uint32_t y = static_cast<uint32_t>(nd >> 32);
            if ((y & 3u) == 3u && (y >> 2) > 0) ++leaves;
        }
        ds->treelet_pairs = static_cast<uint32_t>(std::min<size_t>(sc->tree.pair_nodes.size(), KdTree::kTreeletNodes) / 2);
        const double refs_per_leaf = leaves > 0 ? static_cast<double>(sc->tree.num_pair_refs) / static_cast<double>(leaves) : 0.0;
        ds->two_pass = leaves > 0 && refs_per_leaf <= 6.0;
        // measured (profiles/README.md): the pooled kernel wins from a few thousand leaves up (5000-triangle soup +11 %,
        // 110k-triangle mesh +6 %, 1M-triangle mesh +17 %), ties around 12k leaves and loses on trees that are a handful of
        // big leaves (furnace_test 155 leaves: -4 %; cornell_box, one 36-triangle leaf: -44 %)
        ds->pooled = leaves >= 1024 && refs_per_leaf <= 16.0;
        if (leaves == 1 && sc->tree.num_pair_refs <= static_cast<uint64_t>(kFlatMaxTris) && !flat_planes.empty()) {
            for (uint64_t nd : sc->tree.pair_nodes) {
                const uint32_t y = static_cast<uint32_t>(nd >> 32);
                if ((y & 3u) == 3u && (y >> 2) > 0) {
                    ds->flat = true;
                    ds->flat_first = static_cast<uint32_t>(nd);
                    ds->flat_count = y >> 2;
                    // Scan groups (traverse_flat.cuh): one record per triangle, or per PAIR of coplanar triangles with nearly the
                    // same box (the two halves of a quad -- every face of cornell_box). A pair shares the leader's plane: its
                    // partner's normal / plane offset may differ by at most 8 u / 8 u x scale (u = 2^-24), which the kernel's
                    // error bounds E, F include.
                    float scale = 0.f;
                    for (int c = 0; c < 6; ++c) scale = std::max(scale, std::fabs(sc->tree.box[c]));
                    const float u8 = 8.f * 5.9604645e-8f;
                    auto leaf_id = [&](uint32_t k) { return sc->tree.pair_leaf_refs[ds->flat_first + k]; };
                    auto tri_box = [&](uint32_t id, float* lo, float* hi) {
                        const float* v = &sc->tris.verts[static_cast<size_t>(id) * 9];
                        for (int c = 0; c < 3; ++c) {
                            lo[c] = std::min(v[c], std::min(v[3 + c], v[6 + c]));
                            hi[c] = std::max(v[c], std::max(v[3 + c], v[6 + c]));
                        }
                    };
                    std::vector<char> grouped(ds->flat_count, 0);
                    uint32_t ng = 0;
                    for (uint32_t k = 0; k < ds->flat_count; ++k) {
                        if (grouped[k]) continue;
                        grouped[k] = 1;
                        const uint32_t id = leaf_id(k);
                        const float* pk = &flat_planes[static_cast<size_t>(id) * 4];
                        float lo[3], hi[3];
                        tri_box(id, lo, hi);
                        uint32_t members = k + 1u;
                        if (env_u64("TRN_FLAT_PAIRS", 1) != 0) {
                            for (uint32_t j = k + 1; j < ds->flat_count; ++j) {
                                if (grouped[j]) continue;
                                const uint32_t idj = leaf_id(j);
                                const float* pj = &flat_planes[static_cast<size_t>(idj) * 4];
                                bool match = false;
                                for (float sgn : {1.f, -1.f}) {
                                    bool m = std::fabs(pk[3] - sgn * pj[3]) <= u8 * scale;
                                    for (int c = 0; c < 3; ++c) m = m && std::fabs(pk[c] - sgn * pj[c]) <= u8;
                                    match = match || m;
                                }
                                if (!match) continue;
                                // the union box must not be much larger than either box (else the shared box filters poorly)
                                float lj[3], hj[3];
                                tri_box(idj, lj, hj);
                                bool close = true;
                                for (int c = 0; c < 3; ++c) {
                                    const float eu = std::max(hi[c], hj[c]) - std::min(lo[c], lj[c]);
                                    close = close && eu <= 1.5f * std::min(hi[c] - lo[c], hj[c] - lj[c]) + 1e-3f * scale;
                                }
                                if (!close) continue;
                                for (int c = 0; c < 3; ++c) {
                                    lo[c] = std::min(lo[c], lj[c]);
                                    hi[c] = std::max(hi[c], hj[c]);
                                }
                                grouped[j] = 1;
                                members |= (j + 1u) << 8;
                                break;
                            }
                        }
                        FlatTri& t = ds->flat_params.t[ng];
                        t.nx = pk[0];
                        t.ny = pk[1];
                        t.nz = pk[2];
                        t.dp = pk[3];
                        // the scan's own box: grown by kFlatBoxGrow x scene scale, 50x the exact path's tri_box, so that the
                        // approximate hit point is trusted for all but grazing rays (traverse_flat.cuh)
                        for (int c = 0; c < 3; ++c) {
                            t.blo[c] = lo[c] - kFlatBoxGrow * scale;
                            t.bhi[c] = hi[c] + kFlatBoxGrow * scale;
                        }
                        ds->flat_params.members[ng] = static_cast<uint16_t>(members);
                        ++ng;
                    }
                    ds->flat_groups = ng;
                }
            }
        }
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&ds->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&ds->stream_b, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&ds->stream_copy, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&ds->ev_shaded, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ds->ev_shadow_done[0], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ds->ev_shadow_done[1], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ds->ev_sync, cudaEventDisableTiming));
    ds->counter_slots = 4096;
    CUDA_TRY(cudaMalloc(&ds->d_counters, ds->counter_slots * sizeof(WaveCounters)));
    CUDA_TRY(cudaMallocHost(&ds->h_counters, ds->counter_slots * sizeof(WaveCounters)));
    CUDA_TRY(cudaMalloc(&ds->d_hitcount, sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&ds->d_visits, kVisitSlots * sizeof(unsigned long long)));
    {
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, device));
        int b0 = 1 << 30, b1 = 1 << 30, b2 = 1 << 30;
        int w0 = 0, w1 = 0, w2 = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w0, trace_persistent_ww_kernel<0, true>, 128, 0));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w1, trace_persistent_ww_kernel<1, true>, 128, 0));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w2, trace_persistent_ww_kernel<2, true>, 128, 0));
        b0 = std::min(b0, w0);
        b1 = std::min(b1, w1);
        b2 = std::min(b2, w2);
        int p0 = 0, p1 = 0, p2 = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p0, trace_pooled_kernel<0>, 128, 0));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p1, trace_pooled_kernel<1>, 128, 0));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p2, trace_pooled_kernel<2>, 128, 0));
        int f0 = 0, f1 = 0, f2 = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&f0, trace_flat_kernel<0, kFlatMaxTris>, 128, 0));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&f1, trace_flat_kernel<1, kFlatMaxTris>, 128, 0));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&f2, trace_flat_kernel<2, kFlatMaxTris>, 128, 0));
        ds->grid_flat[0] = std::max(1, f0) * prop.multiProcessorCount;
        ds->grid_flat[1] = std::max(1, f1) * prop.multiProcessorCount;
        ds->grid_flat[2] = std::max(1, f2) * prop.multiProcessorCount;
        ds->grid_pooled[0] = std::max(1, p0) * prop.multiProcessorCount;
        ds->grid_pooled[1] = std::max(1, p1) * prop.multiProcessorCount;
        ds->grid_pooled[2] = std::max(1, p2) * prop.multiProcessorCount;
        ds->grid_closest = std::max(1, b0) * prop.multiProcessorCount;
        ds->grid_shadow = std::max(1, b1) * prop.multiProcessorCount;
        ds->grid_plain = std::max(1, b2) * prop.multiProcessorCount;
    }
    ds->upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out = ds.get();
    sc->devices[device] = std::move(ds);
    return TRN_OK;
}

// the reference-shaped arrays on the device, for the instrumented reference-schedule twins of a device-built scene
static int upload_reference_shape(trn_scene* sc, DeviceScene* ds) {
    if (ds->d_nodes) return TRN_OK;
    ensure_reference_shape(sc);
    CUDA_TRY(cudaMalloc(&ds->d_nodes, std::max<size_t>(sc->gpu_nodes.size() * sizeof(uint2), 16)));
    CUDA_TRY(cudaMemcpy(ds->d_nodes, sc->gpu_nodes.data(), sc->gpu_nodes.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&ds->d_refs, std::max<size_t>(sc->leaf_refs.size() * sizeof(uint32_t), 16)));
    CUDA_TRY(cudaMemcpy(ds->d_refs, sc->leaf_refs.data(), sc->leaf_refs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    ds->dev.nodes = static_cast<const uint2*>(ds->d_nodes);
    ds->dev.leaf_refs = static_cast<const uint32_t*>(ds->d_refs);
    return TRN_OK;
}

// one WaveCounters slot (wave counts + work cursors) from the ring; the ring is zeroed in bulk when it wraps
static int alloc_slot(DeviceScene* ds, cudaStream_t stream, uint32_t* out) {
    if (ds->ring_pos >= ds->counter_slots) {
        CUDA_TRY(cudaDeviceSynchronize()); // every slot handed out so far has been consumed
        ds->ring_pos = 0;
    }
    if (ds->ring_pos == 0) CUDA_TRY(cudaMemsetAsync(ds->d_counters, 0, ds->counter_slots * sizeof(WaveCounters), stream));
    *out = ds->ring_pos++;
    return TRN_OK;
}

// rays a warp of a persistent kernel reserves per atomicAdd on the work cursor
static uint64_t env_u64(const char* name, uint64_t dflt);
static inline uint32_t pool_chunk_for(const struct DeviceScene* ds, uint64_t n);

static inline unsigned persistent_grid(int full, uint64_t n) {
    const uint64_t need = (n + 127) / 128;
    return static_cast<unsigned>(std::max<uint64_t>(1, std::min<uint64_t>(static_cast<uint64_t>(full), need)));
}

// Scheduling of rays onto lanes (results are identical in every mode; profiles/README.md has the A/B numbers):
//   3  pooled kernel (traverse_pooled.cuh): speculative walk + warp-pooled triangle tests -- production for trees with
//      small leaves, closest-hit and shadow waves alike
//   2  persistent warps with lane refill, while-while quantum (traverse_persistent.cuh) -- closest-hit waves of scenes
//      that are a few big leaves (cornell_box)
//   0  one thread per ray (kernels.cuh) -- shadow waves of those scenes
//   4  brute force over the single leaf of a one-leaf tree (traverse_flat.cuh) -- cornell_box, both kinds of wave
// TRN_PERSISTENT=0|2|3 forces one mode for both kinds of wave.
static int persistent_mode(const struct DeviceScene* ds, bool shadow);

static inline uint32_t pool_chunk_for(const DeviceScene* ds, uint64_t n) {
    (void)ds;
    (void)n;
    // measured (profiles/README.md): 32 beats 64 on the mesh, and 64 beats 188..512 on cornell_box's 16 M-ray waves --
    // the cursor atomic is not a hot spot, a balanced tail is what matters
    return static_cast<uint32_t>(env_u64("TRN_POOL_CHUNK", 32));
}

static uint64_t env_u64(const char* name, uint64_t dflt) {
    const char* v = std::getenv(name);
    if (!v || !*v) return dflt;
    return std::strtoull(v, nullptr, 10);
}

static inline unsigned blocks_for(uint64_t n, unsigned bs) { return static_cast<unsigned>((n + bs - 1) / bs); }

static int persistent_mode(const DeviceScene* ds, bool shadow) {
    const char* v = std::getenv("TRN_PERSISTENT");
    if (v && *v) {
        const int m = std::atoi(v);
        if (m == 4) return ds->flat ? 4 : (shadow ? 0 : 2);
        return m == 3 ? 3 : (m != 0 ? 2 : 0);
    }
    if (ds->dev.verbatim) return 0; // every ray takes the reference's schedule: the one-thread-per-ray kernels do exactly that
    // one-leaf scenes (cornell_box): the brute-force kernel, both kinds of wave (TRN_FLAT=0: the while-while / per-ray kernels, A/B)
    if (ds->flat && env_u64("TRN_FLAT", 1) != 0) return 4;
    if (ds->pooled) return 3;
    return shadow ? 0 : 2;
}

// brute-force kernel of a one-leaf scene: the instantiation whose unrolled scan covers the leaf's scan groups (multiples of 4)
template <int MODE, int NT>
static void launch_flat_nt(DeviceScene* ds, unsigned grid, cudaStream_t stream, const float4* ra, const float4* rb, const float4* rc,
                           const float* po, const float* pd, uint32_t n, const uint32_t* count_ptr, uint4* hits, float4* acc) {
    if constexpr (NT > kFlatMaxTris) {
        (void)ds; (void)grid; (void)stream; (void)ra; (void)rb; (void)rc; (void)po; (void)pd; (void)n; (void)count_ptr; (void)hits; (void)acc;
    } else {
        if (ds->flat_groups > static_cast<uint32_t>(NT))
            launch_flat_nt<MODE, NT + 4>(ds, grid, stream, ra, rb, rc, po, pd, n, count_ptr, hits, acc);
        else
            trace_flat_kernel<MODE, NT><<<grid, 128, 0, stream>>>(ds->dev, ds->flat_params, ds->flat_first, ds->flat_count, ds->flat_groups, ra, rb,
                                                                  rc, po, pd, n, count_ptr, hits, acc);
    }
}
template <int MODE>
static void launch_flat(DeviceScene* ds, unsigned grid, cudaStream_t stream, const float4* ra, const float4* rb, const float4* rc,
                        const float* po, const float* pd, uint32_t n, const uint32_t* count_ptr, uint4* hits, float4* acc) {
    launch_flat_nt<MODE, 4>(ds, grid, stream, ra, rb, rc, po, pd, n, count_ptr, hits, acc);
}

// closest-hit traversal of n rays (wave arrays ra/rb, or plain o/d arrays) into hits, in the given scheduling mode
static void launch_closest(DeviceScene* ds, int mode, cudaStream_t stream, const float4* ra, const float4* rb, const float* po,
                           const float* pd, uint32_t n, uint32_t* cursor, uint4* hits, const uint32_t* order = nullptr) {
    const bool plain = po != nullptr;
    if (mode == 4) {
        if (plain)
            launch_flat<2>(ds, persistent_grid(ds->grid_flat[2], n), stream, nullptr, nullptr, nullptr, po, pd, n, nullptr, hits, nullptr);
        else
            launch_flat<0>(ds, persistent_grid(ds->grid_flat[0], n), stream, ra, rb, nullptr, nullptr, nullptr, n, nullptr, hits, nullptr);
    } else if (mode == 3) {
        const int refill = static_cast<int>(env_u64("TRN_PQ_REFILL", 28)), iters = static_cast<int>(env_u64("TRN_PQ_WALK", 12));
        if (plain)
            trace_pooled_kernel<2><<<persistent_grid(ds->grid_pooled[2], n), 128, 0, stream>>>(
                ds->dev, static_cast<const float4*>(ds->d_planes), nullptr, nullptr, nullptr, po, pd, n, nullptr, cursor, hits, nullptr,
                refill, iters, pool_chunk_for(ds, n), static_cast<int>(env_u64("TRN_PQ_GATE", 10)), nullptr);
        else
            trace_pooled_kernel<0><<<persistent_grid(ds->grid_pooled[0], n), 128, 0, stream>>>(
                ds->dev, static_cast<const float4*>(ds->d_planes), ra, rb, nullptr, nullptr, nullptr, n, nullptr, cursor, hits, nullptr,
                refill, iters, pool_chunk_for(ds, n), static_cast<int>(env_u64("TRN_PQ_GATE", 10)), nullptr);
    } else if (mode == 2) {
        const int refill = static_cast<int>(env_u64("TRN_REFILL", 28)), quanta = static_cast<int>(env_u64("TRN_QUANTA", 2));
        if (plain)
            TRN_LAUNCH_WW(2, ds->two_pass, persistent_grid(ds->grid_plain, n), stream, ds->dev, nullptr, nullptr, nullptr, po, pd, n,
                          nullptr, cursor, hits, nullptr, refill, quanta, nullptr, ds->treelet_pairs, pool_chunk_for(ds, n));
        else
            TRN_LAUNCH_WW(0, ds->two_pass, persistent_grid(ds->grid_closest, n), stream, ds->dev, ra, rb, nullptr, nullptr, nullptr, n,
                          nullptr, cursor, hits, nullptr, refill, quanta, order, ds->treelet_pairs, pool_chunk_for(ds, n));
    } else if (plain) {
        trace_closest_plain_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(ds->dev, po, pd, n, hits);
    } else {
        trace_closest_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(ds->dev, ra, rb, n, hits);
    }
}

// any-hit traversal of the shadow wave built by shade_bounce_kernel (count in counters->shadow_count, at most n_max)
static void launch_shadow(DeviceScene* ds, int mode, cudaStream_t stream, const ShadowWave& sw, uint32_t n_max, WaveCounters* counters,
                          float4* acc) {
    if (mode == 4) {
        launch_flat<1>(ds, persistent_grid(ds->grid_flat[1], n_max), stream, sw.a, sw.b, sw.c, nullptr, nullptr, 0, &counters->shadow_count, nullptr, acc);
    } else if (mode == 3) {
        trace_pooled_kernel<1><<<persistent_grid(ds->grid_pooled[1], n_max), 128, 0, stream>>>(
            ds->dev, static_cast<const float4*>(ds->d_planes), sw.a, sw.b, sw.c, nullptr, nullptr, 0,
            &counters->shadow_count, &counters->shadow_cursor, nullptr, acc, static_cast<int>(env_u64("TRN_PQ_REFILL", 28)),
            static_cast<int>(env_u64("TRN_PQ_WALK", 12)), pool_chunk_for(ds, n_max), static_cast<int>(env_u64("TRN_PQ_GATE", 10)), nullptr);
    } else if (mode == 2) {
        TRN_LAUNCH_WW(1, ds->two_pass, persistent_grid(ds->grid_shadow, n_max), stream, ds->dev, sw.a, sw.b,
                      sw.c, nullptr, nullptr, 0, &counters->shadow_count, &counters->shadow_cursor, nullptr, acc,
                      static_cast<int>(env_u64("TRN_REFILL", 26)), static_cast<int>(env_u64("TRN_QUANTA", 2)), nullptr,
                      ds->treelet_pairs, pool_chunk_for(ds, n_max));
    } else {
        trace_shadow_kernel<<<blocks_for(n_max, 128), 128, 0, stream>>>(ds->dev, sw, counters, acc);
    }
}

// wave buffers: `levels` ray waves of `cap` rays + two shadow waves + one hit buffer (+ the raytracer's child-slot array
// and the ray-sorting scratch only when those paths are in use)
static int ensure_waves(DeviceScene* ds, uint64_t cap, int levels, bool need_child_slot, bool need_sort) {
    if (ds->wave_cap != cap) {
        for (void* p : ds->wave_mem) cudaFree(p);
        ds->wave_mem.clear();
        ds->waves.clear();
        cudaFree(ds->shadow_mem);
        ds->shadow_mem = nullptr;
        cudaFree(ds->d_hits);
        ds->d_hits = nullptr;
        cudaFree(ds->d_child_slot);
        ds->d_child_slot = nullptr;
        cudaFree(ds->d_keys); cudaFree(ds->d_keys_alt); cudaFree(ds->d_order); cudaFree(ds->d_order_alt); cudaFree(ds->d_sort_tmp);
        ds->d_keys = ds->d_keys_alt = ds->d_order = ds->d_order_alt = nullptr;
        ds->d_sort_tmp = nullptr;
        ds->wave_cap = cap;
    }
    if (!ds->d_hits) CUDA_TRY(cudaMalloc(&ds->d_hits, cap * sizeof(uint4)));
    if (need_child_slot && !ds->d_child_slot) CUDA_TRY(cudaMalloc(&ds->d_child_slot, cap * sizeof(uint32_t)));
    if (need_sort && !ds->d_keys) {
        CUDA_TRY(cudaMalloc(&ds->d_keys, cap * 4));
        CUDA_TRY(cudaMalloc(&ds->d_keys_alt, cap * 4));
        CUDA_TRY(cudaMalloc(&ds->d_order, cap * 4));
        CUDA_TRY(cudaMalloc(&ds->d_order_alt, cap * 4));
        cub::DoubleBuffer<uint32_t> k(ds->d_keys, ds->d_keys_alt), v(ds->d_order, ds->d_order_alt);
        ds->sort_tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, ds->sort_tmp_bytes, k, v, static_cast<int>(std::min<uint64_t>(cap, 0x7fffffffull)), 0, 18);
        CUDA_TRY(cudaMalloc(&ds->d_sort_tmp, std::max<size_t>(ds->sort_tmp_bytes, 16)));
    }
    if (!ds->shadow_mem) {
        CUDA_TRY(cudaMalloc(&ds->shadow_mem, cap * 6 * sizeof(float4)));
        float4* p = static_cast<float4*>(ds->shadow_mem);
        ds->shadow = ShadowWave{p, p + cap, p + 2 * cap};
        ds->shadow_alt = ShadowWave{p + 3 * cap, p + 4 * cap, p + 5 * cap};
    }
    while (static_cast<int>(ds->waves.size()) < levels) {
        void* mem = nullptr;
        CUDA_TRY(cudaMalloc(&mem, cap * 3 * sizeof(float4)));
        float4* p = static_cast<float4*>(mem);
        ds->wave_mem.push_back(mem);
        ds->waves.push_back(RayWave{p, p + cap, p + 2 * cap});
    }
    return TRN_OK;
}

// Rays per wave buffer. Bigger waves amortise the tails of the persistent kernels and the per-wave launches (whole-job
// throughput on the 1M mesh: 8 Mi rays 1142, 16 Mi 1283, 32 Mi 1357, 64 Mi 1404 Mrays/s, profiles/README.md), and a B200 has
// the memory for them: at least 16 Mi; a call with more primaries than that grows the buffers up to 128 Mi rays (48 B per
// ray and depth level + 112 B of shadow and hit buffers: 39 GB at max-depth 3), within 60 % of the free device memory.
// Buffers only grow. TRN_WAVE_CAP fixes the size (tests: results do not depend on it).
static uint64_t choose_wave_cap(const DeviceScene* ds, uint64_t primaries, int levels) {
    const uint64_t forced = env_u64("TRN_WAVE_CAP", 0);
    if (forced) return std::max<uint64_t>(forced, 1024);
    uint64_t want = 16ull << 20;
    while (want < primaries && want < (128ull << 20)) want <<= 1;
    want = std::max(want, ds->wave_cap);
    size_t free_b = 0, total_b = 0;
    if (want > ds->wave_cap && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
        const uint64_t per_ray = 48ull * static_cast<uint64_t>(levels) + 112ull;
        const uint64_t have = ds->wave_cap * per_ray; // freed when the buffers are re-allocated
        while (want > (16ull << 20) && want > ds->wave_cap && want * per_ray > (free_b + have) * 6 / 10) want >>= 1;
        want = std::max(want, ds->wave_cap);
    }
    return want;
}

// the reference's per-row jitter stream (main.cpp:201-206): xorshift64star<float>(42), two draws per
// pixel sample, restarted every row -> one [width][pps] table of (dx, dy)
static int ensure_jitter(DeviceScene* ds, int width, int pps) {
    if (ds->d_jitter && ds->jitter_w == width && ds->jitter_pps == pps) return TRN_OK;
    cudaFree(ds->d_jitter);
    ds->d_jitter = nullptr;
    std::vector<float2> tab(static_cast<size_t>(width) * pps);
    uint64_t s = 42;
    auto next = [&]() {
        s ^= s >> 12;
        s ^= s << 25;
        s ^= s >> 27;
        uint64_t v = s * 2685821657736338717ULL;
        return std::ldexp(static_cast<float>(v & 0xFFFFFFull), -24);
    };
    for (size_t k = 0; k < tab.size(); ++k) {
        float dx = next();
        float dy = next();
        tab[k] = make_float2(dx, dy);
    }
    CUDA_TRY(cudaMalloc(&ds->d_jitter, tab.size() * sizeof(float2)));
    CUDA_TRY(cudaMemcpy(ds->d_jitter, tab.data(), tab.size() * sizeof(float2), cudaMemcpyHostToDevice));
    ds->jitter_w = width;
    ds->jitter_pps = pps;
    return TRN_OK;
}

static int validate(const trn_camera* cam, const trn_render_config* cfg) {
    if (!cam || !cfg) return fail(TRN_ERR_INVALID, "null camera/config");
    if (cfg->width < 1 || cfg->height < 1) return fail(TRN_ERR_INVALID, "width/height must be >= 1");
    if (cfg->pixel_samples < 1) return fail(TRN_ERR_INVALID, "pixel_samples must be >= 1 (config.h:124)");
    if (cfg->integrator != TRN_PATHTRACER && cfg->integrator != TRN_RAYCASTER && cfg->integrator != TRN_RAYTRACER)
        return fail(TRN_ERR_INVALID, "unknown integrator");
    if (cfg->integrator == TRN_RAYTRACER) {
        if (cfg->max_depth < 1) return fail(TRN_ERR_INVALID, "max_depth must be > 0 (config.h:121)");
        if (cfg->num_lights != 1) return fail(TRN_ERR_INVALID, "the raytracer needs exactly one light (raytracer.cpp:15 takes lights.front())");
        if (!(cfg->shadow_intensity >= 0.f && cfg->shadow_intensity <= 1.f))
            return fail(TRN_ERR_INVALID, "shadow_intensity must be in [0,1] (config.h:123)");
    }
    if (cfg->integrator == TRN_PATHTRACER) {
        if (cfg->max_depth < 1) return fail(TRN_ERR_INVALID, "max_depth must be > 0 (config.h:121)");
        if (cfg->mc_samples < 1) return fail(TRN_ERR_INVALID, "mc_samples must be >= 1");
        double nodes = 0, p = 1;
        for (int d = 0; d <= cfg->max_depth; ++d) {
            nodes += p;
            p *= cfg->mc_samples;
        }
        if (nodes >= 4294967295.0)
            return fail(TRN_ERR_LIMIT, "ray tree has >= 2^32 nodes per primary sample (sum of mc_samples^d, d<=max_depth)");
    }
    if (cfg->num_lights < 0 || cfg->num_lights > 1) return fail(TRN_ERR_INVALID, "0 or 1 lights (main.cpp:123)");
    if (cfg->sample_stride < 0 || cfg->sample_begin < 0) return fail(TRN_ERR_INVALID, "bad sample split");
    if (static_cast<double>(cfg->width) * cfg->height >= 4294967295.0) return fail(TRN_ERR_LIMIT, "more than 2^32 pixels");
    return TRN_OK;
}

static FrameParams make_frame(const trn_camera* cam, const trn_render_config* cfg) {
    FrameParams fp{};
    std::memcpy(fp.cam.pos, cam->pos, sizeof fp.cam.pos);
    std::memcpy(fp.cam.rot, cam->rot, sizeof fp.cam.rot);
    fp.cam.delta_x = cam->delta_x;
    fp.cam.delta_y = cam->delta_y;
    fp.width = cfg->width;
    fp.height = cfg->height;
    fp.pps = cfg->pixel_samples;
    fp.sample_stride = cfg->sample_stride > 0 ? cfg->sample_stride : 1;
    fp.sample_begin = cfg->sample_begin;
    fp.n_local = fp.sample_begin < fp.pps ? (fp.pps - fp.sample_begin + fp.sample_stride - 1) / fp.sample_stride : 0;
    fp.mc_samples = cfg->mc_samples;
    fp.max_depth = cfg->max_depth;
    std::memcpy(fp.bg, cfg->bg_rgba, sizeof fp.bg);
    fp.has_light = cfg->num_lights;
    fp.child_major = static_cast<int32_t>(env_u64("TRN_CHILD_MAJOR", 1));
    std::memcpy(fp.light_pos, cfg->light.pos, sizeof fp.light_pos);
    std::memcpy(fp.light_rgba, cfg->light.rgba, sizeof fp.light_rgba);
    fp.max_visibility = cfg->max_visibility;
    fp.shadow_intensity = cfg->shadow_intensity;
    fp.seed = cfg->seed;
    return fp;
}

struct KernelTimer {
    cudaStream_t stream;
    bool on;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> spans;
    cudaStream_t cur = nullptr; // stream of the span that is open
    explicit KernelTimer(cudaStream_t s) : stream(s), on(g_profiling.load() != 0) {}
    void begin(int kind) { begin(kind, stream); }
    void begin(int kind, cudaStream_t s) {
        if (!on) return;
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, s);
        cur = s;
        spans.push_back({kind, {a, b}});
    }
    void end() {
        if (!on) return;
        cudaEventRecord(spans.back().second.second, cur);
    }
    void collect(trn_stats* st) {
        for (auto& s : spans) {
            float ms = 0;
            cudaEventSynchronize(s.second.second);
            cudaEventElapsedTime(&ms, s.second.first, s.second.second);
            if (st) {
                if (s.first == 0) st->ms_trace += ms;
                else if (s.first == 1) st->ms_shadow += ms;
                else if (s.first == 2) st->ms_shade += ms;
                else st->ms_other += ms;
            }
            cudaEventDestroy(s.second.first);
            cudaEventDestroy(s.second.second);
        }
        spans.clear();
    }
};


// The wavefront over one device: primaries in batches; per batch a depth-first walk over waves.
// A wave at depth d is processed in chunks small enough that its children (<= m per ray) fit the next
// wave buffer, so arbitrary (m, max_depth) work with fixed memory. After each shade launch the wave
// counters come back to the host over a pinned buffer while the shadow kernel of the same chunk is
// already running, so the GPU does not idle on the round trip.
struct Renderer {
    DeviceScene* ds;
    FrameParams fp;
    int integrator;
    float4* acc;
    cudaStream_t stream;
    KernelTimer timer;
    uint64_t rays = 0, prim = 0, shadow = 0, launches = 0;
    uint64_t trace_launches = 0, trace_queries = 0, shadow_launches = 0;
    bool counting = g_counting.load() != 0;
    int mode_closest, mode_shadow;
    bool sort_rays = env_u64("TRN_SORT", 0) != 0 && ds->two_pass; // experiment: (octant, Morton) order for secondary waves; measured no gain (profiles/README.md)
    uint64_t cap;
    bool shadow_overlap = env_u64("TRN_SHADOW_OVERLAP", 1) != 0;
    uint32_t shadow_use = 0;

    Renderer(DeviceScene* d, const FrameParams& f, int integ, float4* a, cudaStream_t s)
        : ds(d), fp(f), integrator(integ), acc(a), stream(s), timer(s), mode_closest(persistent_mode(d, false)),
          mode_shadow(persistent_mode(d, true)), cap(d->wave_cap) {}

    int next_slot(uint32_t* out) { return alloc_slot(ds, stream, out); }

    int process(int depth, const RayWave& wave, uint32_t count, uint64_t first_local_index) {
        const int m = fp.mc_samples;
        if (integrator == TRN_RAYTRACER) return process_raytrace(depth, wave, count, first_local_index);
        const bool spawn = integrator == TRN_PATHTRACER && depth < fp.max_depth;
        const uint64_t chunk_max = spawn ? std::max<uint64_t>(1, cap / static_cast<uint64_t>(m)) : cap;
        for (uint64_t off = 0; off < count; off += chunk_max) {
            const uint32_t n = static_cast<uint32_t>(std::min<uint64_t>(chunk_max, count - off));
            RayWave w{wave.a + off, wave.b + off, wave.T + off};
            uint32_t cs;
            int rc = next_slot(&cs);
            if (rc) return rc;
            // secondary waves of a real tree are consumed in (octant, Morton) order; primaries are coherent as generated
            const uint32_t* order = nullptr;
            if (!counting && mode_closest == 2 && depth > 0 && sort_rays && n >= 4096) {
                timer.begin(3);
                ray_sort_keys_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(ds->dev, w.a, w.b, n, ds->d_keys, ds->d_order);
                cub::DoubleBuffer<uint32_t> k(ds->d_keys, ds->d_keys_alt), v(ds->d_order, ds->d_order_alt);
                size_t tmp = ds->sort_tmp_bytes;
                CUDA_TRY(cub::DeviceRadixSort::SortPairs(ds->d_sort_tmp, tmp, k, v, static_cast<int>(n), 0, 18, stream));
                order = v.Current();
                timer.end();
                launches += 4;
            }
            timer.begin(0);
            if (counting) {
                trace_closest_count_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(ds->dev, w.a, w.b, n, ds->d_hits, ds->d_visits);
                if (mode_closest == 3) // and what the production schedule itself requests (same hits)
                    trace_pooled_kernel<0, true><<<persistent_grid(ds->grid_pooled[0], n), 128, 0, stream>>>(
                        ds->dev, static_cast<const float4*>(ds->d_planes), w.a, w.b, nullptr, nullptr, nullptr, n, nullptr,
                        &ds->d_counters[cs].trace_cursor, ds->d_hits, nullptr, static_cast<int>(env_u64("TRN_PQ_REFILL", 28)),
                        static_cast<int>(env_u64("TRN_PQ_WALK", 12)), pool_chunk_for(ds, n), static_cast<int>(env_u64("TRN_PQ_GATE", 10)),
                        ds->d_visits + 12);
            } else
                launch_closest(ds, mode_closest, stream, w.a, w.b, nullptr, nullptr, n, &ds->d_counters[cs].trace_cursor, ds->d_hits, order);
            timer.end();
            ++launches;
            ++trace_launches;
            trace_queries += n;
            if (integrator == TRN_RAYCASTER) {
                timer.begin(2);
                shade_raycast_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(ds->dev, fp, first_local_index, w, ds->d_hits, n,
                                                                            ds->d_hitcount, acc);
                timer.end();
                ++launches;
                continue;
            }
            rays += n;
            RayWave next = spawn ? ds->waves[depth + 1] : RayWave{nullptr, nullptr, nullptr};
            // The shadow wave of this chunk is traced on a second stream, next to the closest-hit wave of the next depth
            // (both only depend on this shade kernel): the tail of one persistent kernel is filled by the head of the
            // other. Two shadow buffers alternate; a buffer is rewritten only after its last shadow kernel is done.
            // (not while every launch is timed with its own event pair: the spans of the two streams would overlap)
            const bool overlap = shadow_overlap && !counting && !timer.on && fp.has_light;
            const int sb = overlap ? static_cast<int>(shadow_use++ & 1u) : 0;
            const ShadowWave& sw = sb ? ds->shadow_alt : ds->shadow;
            if (overlap && ds->shadow_pending[sb]) CUDA_TRY(cudaStreamWaitEvent(stream, ds->ev_shadow_done[sb], 0));
            timer.begin(2);
            shade_bounce_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(ds->dev, fp, first_local_index, w, ds->d_hits, n, depth, next,
                                                                       sw, ds->d_counters + cs, acc);
            timer.end();
            ++launches;
            CUDA_TRY(cudaMemcpyAsync(ds->h_counters + cs, ds->d_counters + cs, sizeof(WaveCounters), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaEventRecord(ds->ev_sync, stream));
            if (fp.has_light) {
                if (counting) {
                    timer.begin(1);
                    trace_shadow_count_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(ds->dev, sw, ds->d_counters + cs,
                                                                                       mode_shadow == 3 ? nullptr : acc, ds->d_visits + 3);
                    if (mode_shadow == 3)
                        trace_pooled_kernel<1, true><<<persistent_grid(ds->grid_pooled[1], n), 128, 0, stream>>>(
                            ds->dev, static_cast<const float4*>(ds->d_planes), sw.a, sw.b, sw.c, nullptr, nullptr, 0,
                            &ds->d_counters[cs].shadow_count, &ds->d_counters[cs].shadow_cursor, nullptr, acc,
                            static_cast<int>(env_u64("TRN_PQ_REFILL", 28)), static_cast<int>(env_u64("TRN_PQ_WALK", 12)),
                            pool_chunk_for(ds, n), static_cast<int>(env_u64("TRN_PQ_GATE", 10)), ds->d_visits + 22);
                    timer.end();
                } else if (overlap) {
                    CUDA_TRY(cudaEventRecord(ds->ev_shaded, stream));
                    CUDA_TRY(cudaStreamWaitEvent(ds->stream_b, ds->ev_shaded, 0));
                    timer.begin(1, ds->stream_b);
                    launch_shadow(ds, mode_shadow, ds->stream_b, sw, n, ds->d_counters + cs, acc);
                    timer.end();
                    CUDA_TRY(cudaEventRecord(ds->ev_shadow_done[sb], ds->stream_b));
                    ds->shadow_pending[sb] = true;
                } else {
                    timer.begin(1);
                    launch_shadow(ds, mode_shadow, stream, sw, n, ds->d_counters + cs, acc);
                    timer.end();
                }
                ++launches;
                ++shadow_launches;
            }
            CUDA_TRY(cudaEventSynchronize(ds->ev_sync));
            const WaveCounters wc = ds->h_counters[cs];
            shadow += wc.shadow_count;
            if (spawn && wc.next_count) {
                rc = process(depth + 1, next, wc.next_count, first_local_index);
                if (rc) return rc;
            }
        }
        return TRN_OK;
    }

    // raytracer.cpp:6-67: fan-out <= 1 (mirror), the shadow query also scales the child's throughput
    int process_raytrace(int depth, const RayWave& wave, uint32_t n, uint64_t first_local_index) {
        uint32_t cs;
        int rc = next_slot(&cs);
        if (rc) return rc;
        timer.begin(0);
        launch_closest(ds, mode_closest, stream, wave.a, wave.b, nullptr, nullptr, n, &ds->d_counters[cs].trace_cursor, ds->d_hits);
        timer.end();
        ++launches;
        ++trace_launches;
        trace_queries += n;
        rays += n;
        const bool may_spawn = depth < fp.max_depth;
        RayWave next = may_spawn ? ds->waves[depth + 1] : RayWave{nullptr, nullptr, nullptr};
        timer.begin(2);
        shade_raytrace_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(ds->dev, fp, first_local_index, wave, ds->d_hits, n, depth, next,
                                                                     ds->shadow, ds->d_child_slot, ds->d_counters + cs, ds->d_hitcount,
                                                                     acc);
        timer.end();
        ++launches;
        CUDA_TRY(cudaMemcpyAsync(ds->h_counters + cs, ds->d_counters + cs, sizeof(WaveCounters), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaEventRecord(ds->ev_sync, stream));
        timer.begin(1);
        trace_shadow_raytrace_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(ds->dev, ds->shadow, ds->d_child_slot, ds->d_counters + cs,
                                                                            fp.shadow_intensity, next.T, acc);
        timer.end();
        ++launches;
        ++shadow_launches;
        CUDA_TRY(cudaEventSynchronize(ds->ev_sync));
        const WaveCounters wc = ds->h_counters[cs];
        shadow += wc.shadow_count;
        if (may_spawn && wc.next_count) return process_raytrace(depth + 1, next, wc.next_count, first_local_index);
        return TRN_OK;
    }

    int run() {
        const uint64_t total = static_cast<uint64_t>(fp.width) * fp.height * static_cast<uint64_t>(fp.n_local);
        if (integrator != TRN_PATHTRACER) CUDA_TRY(cudaMemsetAsync(ds->d_hitcount, 0, sizeof(unsigned long long), stream));
        if (counting) CUDA_TRY(cudaMemsetAsync(ds->d_visits, 0, kVisitSlots * sizeof(unsigned long long), stream));
        // primaries per batch: as many as fit one wave
        const uint64_t batch = cap;
        for (uint64_t first = 0; first < total; first += batch) {
            const uint32_t n = static_cast<uint32_t>(std::min<uint64_t>(batch, total - first));
            timer.begin(3);
            raygen_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(fp, ds->d_jitter, first, n, ds->waves[0]);
            timer.end();
            ++launches;
            prim += n;
            int rc = process(0, ds->waves[0], n, first);
            if (rc) return rc;
        }
        for (int sb = 0; sb < 2; ++sb) // the caller's stream sees the frame complete only after the last shadow waves
            if (ds->shadow_pending[sb]) {
                CUDA_TRY(cudaStreamWaitEvent(stream, ds->ev_shadow_done[sb], 0));
                ds->shadow_pending[sb] = false;
            }
        if (integrator == TRN_RAYCASTER) {
            unsigned long long hc = 0;
            CUDA_TRY(cudaMemcpyAsync(&hc, ds->d_hitcount, sizeof hc, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            rays = hc; // raycaster.cpp:17 counts a ray only when it hits
        }
        if (integrator == TRN_RAYTRACER) {
            unsigned long long cut = 0;
            CUDA_TRY(cudaMemcpyAsync(&cut, ds->d_hitcount, sizeof cut, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            rays += cut; // calls the depth check rejects are counted too (raytracer.cpp:9-13)
        }
        CUDA_TRY(cudaGetLastError());
        return TRN_OK;
    }
};

static int render_on_device(trn_scene* scene, int device, const trn_camera* cam, const trn_render_config* cfg,
                            float4* d_accum, cudaStream_t user_stream, bool use_user_stream, trn_stats* stats,
                            DeviceScene** ds_out) {
    int rc = validate(cam, cfg);
    if (rc) return rc;
    DeviceScene* ds = nullptr;
    rc = get_device_scene(scene, device, &ds);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> guard(ds->mu);
    CUDA_TRY(cudaSetDevice(ds->device));
    if (ds_out) *ds_out = ds;
    if (g_counting.load() != 0) {
        rc = upload_reference_shape(scene, ds);
        if (rc) return rc;
    }
    FrameParams fp = make_frame(cam, cfg);
    const int levels = cfg->integrator == TRN_RAYCASTER ? 1 : cfg->max_depth + 1;
    const uint64_t cap = choose_wave_cap(ds, static_cast<uint64_t>(fp.width) * fp.height * static_cast<uint64_t>(fp.n_local),
                                         std::max(levels, static_cast<int>(ds->waves.size())));
    rc = ensure_waves(ds, cap, levels, cfg->integrator == TRN_RAYTRACER, env_u64("TRN_SORT", 0) != 0);
    if (rc) return rc;
    rc = ensure_jitter(ds, cfg->width, cfg->pixel_samples);
    if (rc) return rc;
    cudaStream_t stream = use_user_stream ? user_stream : ds->stream;
    Events<2> ev;
    cudaEvent_t& e0 = ev.e[0];
    cudaEvent_t& e1 = ev.e[1];
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        CUDA_TRY(cudaEventCreate(&e0));
        CUDA_TRY(cudaEventCreate(&e1));
        CUDA_TRY(cudaEventRecord(e0, stream));
    }
    Renderer r(ds, fp, cfg->integrator, d_accum, stream);
    rc = r.run();
    if (rc) { // leave no shadow wave pending behind a failed frame
        cudaStreamSynchronize(ds->stream_b);
        ds->shadow_pending[0] = ds->shadow_pending[1] = false;
    }
    if (stats) {
        cudaEventRecord(e1, stream);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        stats->ms_render = ms;
        stats->rays = r.rays;
        stats->prim_rays = r.prim;
        stats->shadow_rays = r.shadow;
        stats->launches = r.launches;
        stats->trace_launches = r.trace_launches;
        stats->trace_queries = r.trace_queries;
        stats->shadow_launches = r.shadow_launches;
        stats->flat_records = (r.mode_closest == 4 && !r.counting) ? ds->flat_groups : 0;
        if (r.counting) {
            unsigned long long v[kVisitSlots];
            cudaMemcpy(v, ds->d_visits, sizeof v, cudaMemcpyDeviceToHost);
            for (int k = 0; k < 10; ++k) {
                stats->trace_pooled[k] = v[12 + k];
                stats->shadow_pooled[k] = v[22 + k];
            }
            stats->trace_inner = v[0]; stats->trace_leaf_nodes = v[1]; stats->trace_tri_tests = v[2];
            stats->shadow_inner = v[3]; stats->shadow_leaf_nodes = v[4]; stats->shadow_tri_tests = v[5];
            stats->trace_actual_inner = v[6]; stats->trace_actual_leaf_nodes = v[7]; stats->trace_actual_tri_tests = v[8];
            stats->shadow_actual_inner = v[9]; stats->shadow_actual_leaf_nodes = v[10]; stats->shadow_actual_tri_tests = v[11];
        }
        r.timer.collect(stats);
    } else {
        r.timer.collect(nullptr);
    }
    return rc;
}

static int ensure_accum(DeviceScene* ds, size_t pixels) {
    if (ds->accum_pixels < pixels) {
        cudaFree(ds->d_accum);
        ds->d_accum = nullptr;
        CUDA_TRY(cudaMalloc(&ds->d_accum, pixels * sizeof(float4)));
        ds->accum_pixels = pixels;
    }
    return TRN_OK;
}

// ------------------------------------------------------------------ NCCL (dlopen)
struct NcclUniqueId {
    char internal[128]; // ncclUniqueId (NCCL_UNIQUE_ID_BYTES)
};
struct NcclApi {
    void* lib = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

static int load_nccl(NcclApi& api) {
    // NCCL's version/debug banner goes to stdout by default; stdout is the image (main.cpp:242)
    setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
    // A libnccl.so.2 the process already holds (e.g. the one bundled with torch, when torch.distributed runs next to this
    // library) is reused -- dlopen matches a resident object by its SONAME -- so that one process never runs two NCCLs.
    const char* names[] = {std::getenv("TRN_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n || !*n) continue;
        api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (api.lib) break;
    }
    if (!api.lib) return fail(TRN_ERR_NCCL, std::string("cannot dlopen libnccl: ") + dlerror());
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(dlsym(api.lib, "ncclCommInitAll"));
    api.Reduce = reinterpret_cast<decltype(api.Reduce)>(dlsym(api.lib, "ncclReduce"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(api.lib, "ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(api.lib, "ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.lib, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.lib, "ncclCommInitRank"));
    if (!api.CommInitAll || !api.Reduce || !api.CommDestroy || !api.GroupStart || !api.GroupEnd || !api.GetUniqueId || !api.CommInitRank) {
        api.lib = nullptr;
        return fail(TRN_ERR_NCCL, "libnccl lacks a required symbol");
    }
    return TRN_OK;
}

static NcclApi g_nccl;
static std::mutex g_nccl_mu;
static int ensure_nccl() {
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    if (g_nccl.lib) return TRN_OK;
    return load_nccl(g_nccl);
}

} // namespace trn

// =========================================================================== C ABI
using namespace trn;

extern "C" {

const char* trn_last_error(void) { return g_last_error.c_str(); }

int32_t trn_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void trn_set_profiling(int32_t enabled) { g_profiling = enabled; }
void trn_set_counting(int32_t enabled) { g_counting = enabled; }

int32_t trn_scene_create(const float* verts, const float* normals, const float* diffuse, uint32_t n, trn_scene** out) {
    return trn_scene_create_ex(verts, normals, diffuse, nullptr, nullptr, n, out);
}

static int32_t scene_create_impl(const float* verts, const float* normals, const float* diffuse, const float* reflective,
                                 const float* reflectivity, uint32_t n, bool on_device, int32_t device, trn_scene** out) {
    if (!verts || !normals || !diffuse || !out) return fail(TRN_ERR_INVALID, "null argument");
    if (n == 0) return fail(TRN_ERR_INVALID, "scene needs at least one triangle (lib/kdtree.cpp:475)");
    if (n >= TRN_MISS_ID) return fail(TRN_ERR_LIMIT, "triangle count must be < 2^30 (lib/kdtree.cpp:476)");
    TRN_GUARD_BEGIN
    for (size_t i = 0; i < size_t(n) * 9; ++i)
        if (!std::isfinite(verts[i])) return fail(TRN_ERR_INVALID, "non-finite vertex coordinate");
    std::unique_ptr<trn_scene> sc(new trn_scene);
    precompute_triangles(verts, normals, diffuse, n, sc->tris, reflective, reflectivity);
    if (on_device) {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            return fail(TRN_ERR_CUDA, "no CUDA device available for the device kd-tree build (use trn_scene_create for the host build)");
        if (device < 0) CUDA_TRY(cudaGetDevice(&device));
        if (device >= ndev) return fail(TRN_ERR_INVALID, "device ordinal out of range");
        std::string err;
        const int rc = build_kdtree_gpu(sc->tris, sc->tree, device, err);
        if (rc) return fail(rc, err);
        sc->gpu_built = true;
    } else {
        build_kdtree(sc->tris, sc->tree, static_cast<int>(env_u64("TRN_BUILD_THREADS", 0)));
        if (env_u64("TRN_REDERIVE_NODES", 0)) { // test hook: the reference-shaped array must follow from the pair layout alone
            sc->tree.nodes.clear();
            reference_shape_from_pairs(sc->tree);
        }
        make_gpu_layout(*sc);
        sc->reference_shape = true;
    }
    if (sc->tree.height + 1 > static_cast<uint64_t>(kStackDepth))
        return fail(TRN_ERR_LIMIT, "kd-tree height " + std::to_string(sc->tree.height) + " exceeds the traversal stack (" +
                                       std::to_string(kStackDepth) + ")");
    *out = sc.release();
    return TRN_OK;
    TRN_GUARD_END
}

int32_t trn_scene_create_ex(const float* verts, const float* normals, const float* diffuse, const float* reflective,
                            const float* reflectivity, uint32_t n, trn_scene** out) {
    // TRN_BUILDER=gpu: run a whole test suite / CLI session on device-built trees without touching the callers
    const char* b = std::getenv("TRN_BUILDER");
    const bool on_device = b && std::strcmp(b, "gpu") == 0;
    return scene_create_impl(verts, normals, diffuse, reflective, reflectivity, n, on_device, -1, out);
}

int32_t trn_scene_create_gpu(const float* verts, const float* normals, const float* diffuse, const float* reflective,
                             const float* reflectivity, uint32_t n, int32_t device, trn_scene** out) {
    return scene_create_impl(verts, normals, diffuse, reflective, reflectivity, n, true, device, out);
}

void trn_scene_destroy(trn_scene* scene) {
    if (!scene) return;
    if (scene->nccl_destroy)
        for (void* c : scene->nccl_comms) scene->nccl_destroy(c);
    for (auto& kv : scene->devices) kv.second->release();
    delete scene;
}

int32_t trn_scene_get_info(const trn_scene* scene, trn_scene_info* info) {
    if (!scene || !info) return fail(TRN_ERR_INVALID, "null argument");
    info->num_triangles = scene->tris.count;
    info->num_nodes = scene->reference_shape ? scene->tree.nodes.size() : scene->tree.expected_nodes;
    info->kdtree_height = scene->tree.height;
    info->num_leaf_refs = scene->tree.num_leaf_refs;
    info->num_cut_nodes = scene->tree.num_cut_nodes;
    std::memcpy(info->box, scene->tree.box, sizeof info->box);
    info->build_ms = scene->tree.build_ms;
    info->upload_ms = 0;
    for (auto& kv : scene->devices) info->upload_ms += kv.second->upload_ms;
    return TRN_OK;
}

int32_t trn_scene_get_nodes(const trn_scene* scene, uint64_t* out_nodes) {
    if (!scene || !out_nodes) return fail(TRN_ERR_INVALID, "null argument");
    TRN_GUARD_BEGIN
    ensure_reference_shape(const_cast<trn_scene*>(scene));
    std::memcpy(out_nodes, scene->tree.nodes.data(), scene->tree.nodes.size() * sizeof(uint64_t));
    return TRN_OK;
    TRN_GUARD_END
}

static int32_t intersect_impl(trn_scene* scene, int32_t device, const float* origins, const float* dirs, uint64_t n,
                              uint32_t* ids, float* rst, uint64_t* counts3);

int32_t trn_intersect(trn_scene* scene, int32_t device, const float* origins, const float* dirs, uint64_t n,
                      uint32_t* ids, float* rst) {
    return intersect_impl(scene, device, origins, dirs, n, ids, rst, nullptr);
}

int32_t trn_intersect_counted(trn_scene* scene, int32_t device, const float* origins, const float* dirs, uint64_t n,
                              uint32_t* ids, float* rst, uint64_t* counts3) {
    if (!counts3) return fail(TRN_ERR_INVALID, "null argument");
    return intersect_impl(scene, device, origins, dirs, n, ids, rst, counts3);
}

static int32_t intersect_impl(trn_scene* scene, int32_t device, const float* origins, const float* dirs, uint64_t n,
                              uint32_t* ids, float* rst, uint64_t* counts3) {
    if (!scene || !origins || !dirs || !ids || !rst) return fail(TRN_ERR_INVALID, "null argument");
    DeviceScene* ds = nullptr;
    int rc = get_device_scene(scene, device, &ds);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> guard(ds->mu);
    CUDA_TRY(cudaSetDevice(ds->device));
    if (counts3) {
        rc = upload_reference_shape(scene, ds);
        if (rc) return rc;
    }
    const uint64_t chunk = 8ull << 20;
    const uint64_t cn = std::min<uint64_t>(chunk, std::max<uint64_t>(n, 1));
    DevBuf b_o, b_d, b_rst, b_h, b_ids;
    CUDA_TRY(b_o.alloc(cn * 12));
    CUDA_TRY(b_d.alloc(cn * 12));
    CUDA_TRY(b_rst.alloc(cn * 12));
    CUDA_TRY(b_h.alloc(cn * 16));
    CUDA_TRY(b_ids.alloc(cn * 4));
    float *d_o = b_o.as<float>(), *d_d = b_d.as<float>(), *d_rst = b_rst.as<float>();
    uint4* d_h = b_h.as<uint4>();
    uint32_t* d_ids = b_ids.as<uint32_t>();
    if (counts3) CUDA_TRY(cudaMemsetAsync(ds->d_visits, 0, kVisitSlots * sizeof(unsigned long long), ds->stream));
    for (uint64_t off = 0; off < n; off += chunk) {
        const uint32_t c = static_cast<uint32_t>(std::min<uint64_t>(chunk, n - off));
        CUDA_TRY(cudaMemcpyAsync(d_o, origins + 3 * off, size_t(c) * 12, cudaMemcpyHostToDevice, ds->stream));
        CUDA_TRY(cudaMemcpyAsync(d_d, dirs + 3 * off, size_t(c) * 12, cudaMemcpyHostToDevice, ds->stream));
        if (counts3)
            trace_closest_plain_count_kernel<<<blocks_for(c, 128), 128, 0, ds->stream>>>(ds->dev, d_o, d_d, c, d_h, ds->d_visits);
        else {
            uint32_t cs;
            int rc2 = alloc_slot(ds, ds->stream, &cs);
            if (rc2) return rc2;
            launch_closest(ds, persistent_mode(ds, false), ds->stream, nullptr, nullptr, d_o, d_d, c, &ds->d_counters[cs].trace_cursor, d_h);
        }
        unpack_hits_kernel<<<blocks_for(c, 256), 256, 0, ds->stream>>>(d_h, c, d_ids, d_rst);
        CUDA_TRY(cudaMemcpyAsync(ids + off, d_ids, size_t(c) * 4, cudaMemcpyDeviceToHost, ds->stream));
        CUDA_TRY(cudaMemcpyAsync(rst + 3 * off, d_rst, size_t(c) * 12, cudaMemcpyDeviceToHost, ds->stream));
        CUDA_TRY(cudaStreamSynchronize(ds->stream));
    }
    CUDA_TRY(cudaGetLastError());
    if (counts3) {
        unsigned long long v[9];
        CUDA_TRY(cudaMemcpy(v, ds->d_visits, sizeof v, cudaMemcpyDeviceToHost));
        counts3[0] = v[0]; counts3[1] = v[1]; counts3[2] = v[2];
        counts3[3] = v[6]; counts3[4] = v[7]; counts3[5] = v[8];
    }
    return TRN_OK;
}

int32_t trn_primary_hits(trn_scene* scene, int32_t device, const trn_camera* cam, const trn_render_config* cfg,
                         uint32_t* ids, float* rst) {
    if (!scene || !ids || !rst) return fail(TRN_ERR_INVALID, "null argument");
    int rc = validate(cam, cfg);
    if (rc) return rc;
    DeviceScene* ds = nullptr;
    rc = get_device_scene(scene, device, &ds);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> guard(ds->mu);
    CUDA_TRY(cudaSetDevice(ds->device));
    trn_render_config c2 = *cfg;
    c2.sample_begin = 0;
    c2.sample_stride = 1;
    FrameParams fp = make_frame(cam, &c2);
    const uint64_t cap = choose_wave_cap(ds, 0, std::max(1, static_cast<int>(ds->waves.size())));
    rc = ensure_waves(ds, cap, 1, false, false);
    if (rc) return rc;
    rc = ensure_jitter(ds, cfg->width, cfg->pixel_samples);
    if (rc) return rc;
    const uint64_t total = static_cast<uint64_t>(fp.width) * fp.height * static_cast<uint64_t>(fp.n_local);
    const uint64_t cn = std::min<uint64_t>(cap, total);
    DevBuf b_ids, b_rst;
    CUDA_TRY(b_ids.alloc(cn * 4));
    CUDA_TRY(b_rst.alloc(cn * 12));
    uint32_t* d_ids = b_ids.as<uint32_t>();
    float* d_rst = b_rst.as<float>();
    for (uint64_t first = 0; first < total; first += cap) {
        const uint32_t n = static_cast<uint32_t>(std::min<uint64_t>(cap, total - first));
        raygen_kernel<<<blocks_for(n, 256), 256, 0, ds->stream>>>(fp, ds->d_jitter, first, n, ds->waves[0]);
        {
            uint32_t cs;
            rc = alloc_slot(ds, ds->stream, &cs);
            if (rc) return rc;
            launch_closest(ds, persistent_mode(ds, false), ds->stream, ds->waves[0].a, ds->waves[0].b, nullptr, nullptr, n,
                           &ds->d_counters[cs].trace_cursor, ds->d_hits);
        }
        unpack_hits_kernel<<<blocks_for(n, 256), 256, 0, ds->stream>>>(ds->d_hits, n, d_ids, d_rst);
        CUDA_TRY(cudaMemcpyAsync(ids + first, d_ids, size_t(n) * 4, cudaMemcpyDeviceToHost, ds->stream));
        CUDA_TRY(cudaMemcpyAsync(rst + 3 * first, d_rst, size_t(n) * 12, cudaMemcpyDeviceToHost, ds->stream));
        CUDA_TRY(cudaStreamSynchronize(ds->stream));
    }
    CUDA_TRY(cudaGetLastError());
    return TRN_OK;
}

int32_t trn_render_device(trn_scene* scene, int32_t device, const trn_camera* cam, const trn_render_config* cfg,
                          float* d_accum_rgba, void* cuda_stream, trn_stats* stats) {
    if (!scene || !d_accum_rgba) return fail(TRN_ERR_INVALID, "null argument");
    TRN_GUARD_BEGIN
    return render_on_device(scene, device, cam, cfg, reinterpret_cast<float4*>(d_accum_rgba),
                            static_cast<cudaStream_t>(cuda_stream), true, stats, nullptr);
    TRN_GUARD_END
}

// memset + render + (reduce) + D2H of one device's share, the common body of trn_render / trn_render_rank.
// comm == nullptr: single device. Times the three parts with events on the device's stream.
static int render_share(trn_scene* scene, int device, const trn_camera* cam, const trn_render_config* cfg, trn_comm* comm,
                        float* out_rgba_sum, trn_stats* stats) {
    int rc = validate(cam, cfg);
    if (rc) return rc;
    DeviceScene* ds = nullptr;
    rc = get_device_scene(scene, device, &ds);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> guard(ds->mu);
    CUDA_TRY(cudaSetDevice(ds->device));
    const size_t pixels = static_cast<size_t>(cfg->width) * cfg->height;
    rc = ensure_accum(ds, pixels);
    if (rc) return rc;
    Events<4> ev;
    CUDA_TRY(ev.create());
    CUDA_TRY(cudaEventRecord(ev.e[0], ds->stream));
    CUDA_TRY(cudaMemsetAsync(ds->d_accum, 0, pixels * sizeof(float4), ds->stream));
    trn_render_config c = *cfg;
    if (comm) { // sample split: rank g of G renders i = begin + stride * (g + G k)
        const int base_stride = cfg->sample_stride > 0 ? cfg->sample_stride : 1;
        c.sample_begin = cfg->sample_begin + comm->rank * base_stride;
        c.sample_stride = base_stride * comm->nranks;
    }
    trn_stats local;
    rc = render_on_device(scene, ds->device, cam, &c, ds->d_accum, ds->stream, true, &local, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ev.e[1], ds->stream));
    if (comm && comm->nranks > 1) {
        // ONE ncclReduce(sum) of the W*H*4 float accumulation buffers onto rank 0 (in place on the root)
        const int nrc = g_nccl.Reduce(ds->d_accum, ds->d_accum, pixels * 4, /*ncclFloat32*/ 7, /*ncclSum*/ 0, 0, comm->comm, ds->stream);
        if (nrc != 0) return fail(TRN_ERR_NCCL, std::string("ncclReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "?"));
    }
    CUDA_TRY(cudaEventRecord(ev.e[2], ds->stream));
    const bool root = !comm || comm->rank == 0;
    if (root && out_rgba_sum)
        CUDA_TRY(cudaMemcpyAsync(out_rgba_sum, ds->d_accum, pixels * sizeof(float4), cudaMemcpyDeviceToHost, ds->stream));
    CUDA_TRY(cudaEventRecord(ev.e[3], ds->stream));
    CUDA_TRY(cudaEventSynchronize(ev.e[3]));
    float ms = 0, ms_r = 0, ms_c = 0;
    cudaEventElapsedTime(&ms, ev.e[0], ev.e[3]);
    cudaEventElapsedTime(&ms_r, ev.e[1], ev.e[2]);
    cudaEventElapsedTime(&ms_c, ev.e[2], ev.e[3]);
    local.ms_render = ms;
    local.ms_reduce = ms_r;
    local.ms_d2h = ms_c;
    if (stats) *stats = local;
    return TRN_OK;
}

int32_t trn_render(trn_scene* scene, int32_t device, const trn_camera* cam, const trn_render_config* cfg,
                   float* out_rgba_sum, trn_stats* stats) {
    if (!scene || !out_rgba_sum) return fail(TRN_ERR_INVALID, "null argument");
    TRN_GUARD_BEGIN
    return render_share(scene, device, cam, cfg, nullptr, out_rgba_sum, stats);
    TRN_GUARD_END
}

// ---- process-per-GPU communicator: one rank per process (torchrun / mpirun), NCCL over NVLink
int32_t trn_comm_unique_id(uint8_t* id128) {
    if (!id128) return fail(TRN_ERR_INVALID, "null argument");
    TRN_GUARD_BEGIN
    int rc = ensure_nccl();
    if (rc) return rc;
    NcclUniqueId id;
    const int nrc = g_nccl.GetUniqueId(&id);
    if (nrc != 0) return fail(TRN_ERR_NCCL, std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "?"));
    std::memcpy(id128, id.internal, 128);
    return TRN_OK;
    TRN_GUARD_END
}

int32_t trn_comm_init_rank(const uint8_t* id128, int32_t nranks, int32_t rank, int32_t device, trn_comm** out) {
    if (!id128 || !out || nranks < 1 || rank < 0 || rank >= nranks) return fail(TRN_ERR_INVALID, "bad communicator arguments");
    TRN_GUARD_BEGIN
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(TRN_ERR_CUDA, "no CUDA device available (turner_b200 has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(TRN_ERR_INVALID, "device ordinal out of range");
    CUDA_TRY(cudaSetDevice(device));
    std::unique_ptr<trn_comm> c(new trn_comm);
    c->nranks = nranks;
    c->rank = rank;
    c->device = device;
    if (nranks > 1) { // a one-rank job needs no NCCL (and no libnccl on the machine)
        int rc = ensure_nccl();
        if (rc) return rc;
        NcclUniqueId id;
        std::memcpy(id.internal, id128, 128);
        const int nrc = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
        if (nrc != 0) return fail(TRN_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "?"));
    }
    *out = c.release();
    return TRN_OK;
    TRN_GUARD_END
}

void trn_comm_destroy(trn_comm* comm) {
    if (!comm) return;
    if (comm->comm && g_nccl.CommDestroy) {
        cudaSetDevice(comm->device);
        g_nccl.CommDestroy(comm->comm);
    }
    delete comm;
}

int32_t trn_render_rank(trn_scene* scene, trn_comm* comm, const trn_camera* cam, const trn_render_config* cfg,
                        float* out_rgba_sum, trn_stats* stats) {
    if (!scene || !comm) return fail(TRN_ERR_INVALID, "null argument");
    TRN_GUARD_BEGIN
    return render_share(scene, comm->device, cam, cfg, comm, out_rgba_sum, stats);
    TRN_GUARD_END
}

int32_t trn_render_multi(trn_scene* scene, const int32_t* devices, int32_t num_devices, const trn_camera* cam,
                         const trn_render_config* cfg, float* out_rgba_sum, trn_stats* stats) {
    if (!scene || !devices || num_devices < 1 || !out_rgba_sum) return fail(TRN_ERR_INVALID, "null argument");
    if (num_devices == 1) return trn_render(scene, devices[0], cam, cfg, out_rgba_sum, stats);
    TRN_GUARD_BEGIN
    int rc = validate(cam, cfg);
    if (rc) return rc;
    rc = ensure_nccl();
    if (rc) return rc;
    NcclApi& nccl = g_nccl;
    const size_t pixels = static_cast<size_t>(cfg->width) * cfg->height;
    std::vector<DeviceScene*> dss(num_devices, nullptr);
    for (int g = 0; g < num_devices; ++g) {
        for (int h = 0; h < g; ++h)
            if (devices[h] == devices[g]) return fail(TRN_ERR_INVALID, "trn_render_multi: a device is listed twice");
        rc = get_device_scene(scene, devices[g], &dss[g]);
        if (rc) return rc;
        CUDA_TRY(cudaSetDevice(dss[g]->device));
        rc = ensure_accum(dss[g], pixels);
        if (rc) return rc;
    }
    std::vector<int> devs(devices, devices + num_devices);
    int nrc = 0;
    if (scene->nccl_devs != devs) {
        for (void* c : scene->nccl_comms) nccl.CommDestroy(c);
        scene->nccl_comms.assign(num_devices, nullptr);
        scene->nccl_devs.clear();
        nrc = nccl.CommInitAll(scene->nccl_comms.data(), num_devices, devs.data());
        if (nrc != 0) {
            scene->nccl_comms.clear();
            return fail(TRN_ERR_NCCL, std::string("ncclCommInitAll: ") + (nccl.GetErrorString ? nccl.GetErrorString(nrc) : "?"));
        }
        scene->nccl_devs = devs;
        scene->nccl_destroy = nccl.CommDestroy;
    }
    std::vector<void*>& comms = scene->nccl_comms;

    auto t0 = std::chrono::steady_clock::now();
    std::vector<trn_stats> st(num_devices);
    std::vector<int> rcs(num_devices, 0);
    std::vector<std::string> errs(num_devices);
    std::vector<std::thread> workers;
    const int base_begin = cfg->sample_begin;
    const int base_stride = cfg->sample_stride > 0 ? cfg->sample_stride : 1;
    for (int g = 0; g < num_devices; ++g) {
        workers.emplace_back([&, g]() {
            try {
                cudaSetDevice(dss[g]->device);
                trn_render_config c = *cfg;
                c.sample_begin = base_begin + g * base_stride; // sample split: i = begin + stride*(g + G*k)
                c.sample_stride = base_stride * num_devices;
                cudaMemsetAsync(dss[g]->d_accum, 0, pixels * sizeof(float4), dss[g]->stream);
                rcs[g] = render_on_device(scene, dss[g]->device, cam, &c, dss[g]->d_accum, dss[g]->stream, true, &st[g], nullptr);
                if (rcs[g]) errs[g] = g_last_error;
            } catch (const std::exception& e) {
                rcs[g] = TRN_ERR_LIMIT;
                errs[g] = e.what();
            }
        });
    }
    for (auto& w : workers) w.join();
    for (int g = 0; g < num_devices; ++g)
        if (rcs[g]) return fail(rcs[g], errs[g]);
    auto t1 = std::chrono::steady_clock::now();
    // one ncclReduce(sum) of the float accumulation buffers onto devices[0]
    nccl.GroupStart();
    for (int g = 0; g < num_devices; ++g) {
        cudaSetDevice(dss[g]->device);
        nrc = nccl.Reduce(dss[g]->d_accum, dss[g]->d_accum, pixels * 4, /*ncclFloat32*/ 7, /*ncclSum*/ 0, 0, comms[g], dss[g]->stream);
        if (nrc != 0) break;
    }
    int erc = nccl.GroupEnd();
    if (nrc != 0 || erc != 0) return fail(TRN_ERR_NCCL, "ncclReduce failed");
    for (int g = 0; g < num_devices; ++g) {
        cudaSetDevice(dss[g]->device);
        CUDA_TRY(cudaStreamSynchronize(dss[g]->stream));
    }
    auto t2 = std::chrono::steady_clock::now();
    CUDA_TRY(cudaSetDevice(dss[0]->device));
    CUDA_TRY(cudaMemcpy(out_rgba_sum, dss[0]->d_accum, pixels * sizeof(float4), cudaMemcpyDeviceToHost));
    auto t3 = std::chrono::steady_clock::now();
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        for (auto& s : st) {
            stats->rays += s.rays;
            stats->prim_rays += s.prim_rays;
            stats->shadow_rays += s.shadow_rays;
            stats->launches += s.launches;
            stats->ms_trace = std::max(stats->ms_trace, s.ms_trace);
            stats->ms_shadow = std::max(stats->ms_shadow, s.ms_shadow);
            stats->ms_shade = std::max(stats->ms_shade, s.ms_shade);
            stats->ms_other = std::max(stats->ms_other, s.ms_other);
        }
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double, std::milli>(b - a).count();
        };
        stats->ms_render = ms(t0, t3);
        stats->ms_reduce = ms(t1, t2);
        stats->ms_d2h = ms(t2, t3);
    }
    return TRN_OK;
    TRN_GUARD_END
}

// ---- asynchronous frames: trn_render_async returns at once; the frame is rendered by a worker thread into one of two
// accumulation buffers and copied to the host on a copy stream, so the D2H of frame k runs next to the render of k+1.
struct trn_job {
    std::thread worker;
    int rc = 0;
    std::string err;
    trn_stats stats{};
};

int32_t trn_render_async(trn_scene* scene, int32_t device, const trn_camera* cam, const trn_render_config* cfg,
                         float* out_rgba_sum, trn_job** job_out) {
    if (!scene || !out_rgba_sum || !job_out) return fail(TRN_ERR_INVALID, "null argument");
    TRN_GUARD_BEGIN
    int rc = validate(cam, cfg);
    if (rc) return rc;
    DeviceScene* ds = nullptr;
    rc = get_device_scene(scene, device, &ds);
    if (rc) return rc;
    std::unique_ptr<trn_job> job(new trn_job);
    trn_job* j = job.get();
    const trn_camera cam_c = *cam;
    const trn_render_config cfg_c = *cfg;
    j->worker = std::thread([scene, ds, cam_c, cfg_c, out_rgba_sum, j]() {
        auto body = [&]() -> int {
            const size_t pixels = static_cast<size_t>(cfg_c.width) * cfg_c.height;
            cudaEvent_t done = nullptr;
            float4* buf = nullptr;
            {
                std::lock_guard<std::recursive_mutex> guard(ds->mu); // renders serialise; the copy below does not hold it
                CUDA_TRY(cudaSetDevice(ds->device));
                const int slot = static_cast<int>(ds->async_use++ & 1u);
                if (ds->async_pixels[slot] < pixels) {
                    cudaFree(ds->d_async[slot]);
                    ds->d_async[slot] = nullptr;
                    ds->async_pixels[slot] = 0;
                    CUDA_TRY(cudaMalloc(&ds->d_async[slot], pixels * sizeof(float4)));
                    ds->async_pixels[slot] = pixels;
                }
                if (!ds->ev_async[slot]) CUDA_TRY(cudaEventCreateWithFlags(&ds->ev_async[slot], cudaEventDisableTiming));
                buf = ds->d_async[slot];
                // the previous copy out of this buffer (two frames ago) must have left it
                CUDA_TRY(cudaStreamWaitEvent(ds->stream, ds->ev_async[slot], 0));
                CUDA_TRY(cudaMemsetAsync(buf, 0, pixels * sizeof(float4), ds->stream));
                int rc2 = render_on_device(scene, ds->device, &cam_c, &cfg_c, buf, ds->stream, true, &j->stats, nullptr);
                if (rc2) return rc2;
                done = ds->ev_async[slot];
                CUDA_TRY(cudaEventRecord(done, ds->stream));
                CUDA_TRY(cudaStreamWaitEvent(ds->stream_copy, done, 0));
                CUDA_TRY(cudaMemcpyAsync(out_rgba_sum, buf, pixels * sizeof(float4), cudaMemcpyDeviceToHost, ds->stream_copy));
                CUDA_TRY(cudaEventRecord(done, ds->stream_copy)); // now marks "copy finished": guards the buffer's reuse
            }
            CUDA_TRY(cudaEventSynchronize(done));
            return TRN_OK;
        };
        try {
            j->rc = body();
            if (j->rc) j->err = g_last_error;
        } catch (const std::exception& e) {
            j->rc = TRN_ERR_LIMIT;
            j->err = e.what();
        }
    });
    *job_out = job.release();
    return TRN_OK;
    TRN_GUARD_END
}

int32_t trn_wait(trn_job* job, trn_stats* stats) {
    if (!job) return fail(TRN_ERR_INVALID, "null argument");
    if (job->worker.joinable()) job->worker.join();
    const int rc = job->rc;
    if (rc) g_last_error = job->err;
    else if (stats) *stats = job->stats;
    delete job;
    return rc;
}

// ---- occlusion parity hook: the any-hit predicate of pathtracer.cpp:49-53 for arbitrary rays, through the production
// shadow kernels (a shadow wave whose "pixel" is the ray index and whose contribution is 1)
int32_t trn_occluded(trn_scene* scene, int32_t device, const float* origins, const float* dirs, const float* tmax, uint64_t n,
                     uint8_t* occluded) {
    if (!scene || !origins || !dirs || !tmax || !occluded) return fail(TRN_ERR_INVALID, "null argument");
    TRN_GUARD_BEGIN
    DeviceScene* ds = nullptr;
    int rc = get_device_scene(scene, device, &ds);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> guard(ds->mu);
    CUDA_TRY(cudaSetDevice(ds->device));
    const uint64_t chunk = 4ull << 20;
    const uint64_t cn = std::min<uint64_t>(chunk, std::max<uint64_t>(n, 1));
    DevBuf b_o, b_d, b_t, b_w, b_acc, b_out;
    CUDA_TRY(b_o.alloc(cn * 12));
    CUDA_TRY(b_d.alloc(cn * 12));
    CUDA_TRY(b_t.alloc(cn * 4));
    CUDA_TRY(b_w.alloc(cn * 3 * sizeof(float4)));
    CUDA_TRY(b_acc.alloc(cn * sizeof(float4)));
    CUDA_TRY(b_out.alloc(cn));
    float4* w = b_w.as<float4>();
    const ShadowWave sw{w, w + cn, w + 2 * cn};
    for (uint64_t off = 0; off < n; off += chunk) {
        const uint32_t c = static_cast<uint32_t>(std::min<uint64_t>(chunk, n - off));
        uint32_t cs;
        rc = alloc_slot(ds, ds->stream, &cs);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(b_o.p, origins + 3 * off, size_t(c) * 12, cudaMemcpyHostToDevice, ds->stream));
        CUDA_TRY(cudaMemcpyAsync(b_d.p, dirs + 3 * off, size_t(c) * 12, cudaMemcpyHostToDevice, ds->stream));
        CUDA_TRY(cudaMemcpyAsync(b_t.p, tmax + off, size_t(c) * 4, cudaMemcpyHostToDevice, ds->stream));
        CUDA_TRY(cudaMemsetAsync(b_acc.p, 0, size_t(c) * sizeof(float4), ds->stream));
        pack_shadow_queries_kernel<<<blocks_for(c, 256), 256, 0, ds->stream>>>(b_o.as<float>(), b_d.as<float>(), b_t.as<float>(), c, sw,
                                                                              ds->d_counters + cs);
        launch_shadow(ds, persistent_mode(ds, true), ds->stream, sw, c, ds->d_counters + cs, b_acc.as<float4>());
        unpack_occlusion_kernel<<<blocks_for(c, 256), 256, 0, ds->stream>>>(b_acc.as<float4>(), c, b_out.as<uint8_t>());
        CUDA_TRY(cudaMemcpyAsync(occluded + off, b_out.p, c, cudaMemcpyDeviceToHost, ds->stream));
        CUDA_TRY(cudaStreamSynchronize(ds->stream));
    }
    CUDA_TRY(cudaGetLastError());
    return TRN_OK;
    TRN_GUARD_END
}

// ---- measured gather peaks of this device (the denominators of bench.py's roofline): random 16-byte gathers, the access
// shape of the traversal kernels, over a working set that lives in L1 / in L2 / in HBM.
int32_t trn_measure_gather_peak(int32_t device, uint64_t set_bytes, int32_t mode, double* gbps) {
    if (!gbps || set_bytes < 4096) return fail(TRN_ERR_INVALID, "bad argument");
    TRN_GUARD_BEGIN
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(TRN_ERR_CUDA, "no CUDA device available");
    if (device < 0) CUDA_TRY(cudaGetDevice(&device));
    if (device >= ndev) return fail(TRN_ERR_INVALID, "device ordinal out of range");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const uint64_t elems = set_bytes / 16;
    DevBuf buf, sink;
    CUDA_TRY(buf.alloc(elems * 16));
    CUDA_TRY(sink.alloc(16));
    CUDA_TRY(cudaMemset(buf.p, 1, elems * 16));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gather_peak_kernel, 256, 0));
    const int grid = std::max(1, per_sm) * prop.multiProcessorCount;
    const int iters = mode == 0 ? 4096 : 256; // loads per thread
    Events<2> ev;
    CUDA_TRY(ev.create());
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) { // first repetition warms the caches
        CUDA_TRY(cudaEventRecord(ev.e[0], nullptr));
        gather_peak_kernel<<<grid, 256>>>(buf.as<uint4>(), static_cast<uint32_t>(elems), iters, mode, sink.as<uint4>());
        CUDA_TRY(cudaEventRecord(ev.e[1], nullptr));
        CUDA_TRY(cudaEventSynchronize(ev.e[1]));
        float ms = 0;
        cudaEventElapsedTime(&ms, ev.e[0], ev.e[1]);
        const double g = 16.0 * iters * 256.0 * grid / (ms * 1e6);
        if (rep > 0) best = std::max(best, g);
    }
    CUDA_TRY(cudaGetLastError());
    *gbps = best;
    return TRN_OK;
    TRN_GUARD_END
}

} // extern "C"

// ------------------------------------------------------------------ kdtree.cache (main.cpp:142-167)
namespace {
struct CacheWriter {
    FILE* f;
    bool ok = true;
    void raw(const void* p, size_t n) { ok = ok && std::fwrite(p, 1, n, f) == n; }
    void f32(const float* p, size_t n) { raw(p, n * sizeof(float)); }
    void u64(uint64_t v) { raw(&v, 8); }
};
struct CacheReader {
    FILE* f;
    bool ok = true;
    void raw(void* p, size_t n) { ok = ok && std::fread(p, 1, n, f) == n; }
    uint64_t u64() {
        uint64_t v = 0;
        raw(&v, 8);
        return v;
    }
};
} // namespace

using trn::fail;

int32_t trn_scene_save_cache(const trn_scene* scene, const char* path) {
    if (!scene || !path) return trn::fail(TRN_ERR_INVALID, "null argument");
    trn::ensure_reference_shape(const_cast<trn_scene*>(scene));
    FILE* f = std::fopen(path, "wb");
    if (!f) return trn::fail(TRN_ERR_IO, std::string("cannot write ") + path);
    CacheWriter w{f};
    const uint8_t little = 1; // cereal::PortableBinaryOutputArchive: endianness flag of the writer
    w.raw(&little, 1);
    const trn::HostTriangles& t = scene->tris;
    w.u64(t.count);
    const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
    for (uint32_t i = 0; i < t.count; ++i) { // Triangle::serialize, lib/triangle.h:89-92
        const float* is = &t.isect[size_t(i) * 16]; // v0 n u v | uv vv uu denom
        const float* sh = &t.shade[size_t(i) * 16]; // n0 n1 n2 reflectivity - - rgba
        w.f32(&t.verts[size_t(i) * 9], 9);          // vertices
        w.f32(sh, 9);                               // normals
        w.f32(zero4, 4);                            // ambient (not carried, never read by a tracer)
        w.f32(sh + 12, 4);                          // diffuse
        w.f32(sh + 12, 4);                          // emissive: main.cpp:43 loads it from the DIFFUSE key
        w.f32(&t.mirror[size_t(i) * 4], 4);         // reflective
        w.f32(sh + 9, 1);                           // reflectivity
        w.f32(is + 6, 3);                           // u
        w.f32(is + 9, 3);                           // v
        w.f32(is + 3, 3);                           // normal
        w.f32(is + 12, 4);                          // uv vv uu denom
    }
    w.f32(scene->tree.box, 6);
    w.u64(scene->tree.nodes.size());
    w.raw(scene->tree.nodes.data(), scene->tree.nodes.size() * 8);
    const bool ok = w.ok && std::fclose(f) == 0;
    return ok ? TRN_OK : trn::fail(TRN_ERR_IO, std::string("short write to ") + path);
}

int32_t trn_scene_load_cache(const char* path, trn_scene** out) {
    if (!path || !out) return trn::fail(TRN_ERR_INVALID, "null argument");
    FILE* f = std::fopen(path, "rb");
    if (!f) return trn::fail(TRN_ERR_IO, std::string("cannot read ") + path);
    TRN_GUARD_BEGIN
    struct Closer {
        FILE*& f;
        ~Closer() {
            if (f) std::fclose(f);
        }
    } closer{f};
    CacheReader r{f};
    uint8_t little = 0;
    r.raw(&little, 1);
    const uint64_t n = r.u64();
    // the header is untrusted: the triangle count must fit the file (1 flag + 8 count + 192 per triangle + 24 box + 8 node
    // count) before anything is sized from it
    uint64_t file_size = 0;
    if (std::fseek(f, 0, SEEK_END) == 0) {
        const long e = std::ftell(f);
        if (e > 0) file_size = static_cast<uint64_t>(e);
    }
    std::fseek(f, 9, SEEK_SET);
    if (!r.ok || little != 1 || n == 0 || n >= TRN_MISS_ID || file_size < 41 || n > (file_size - 41) / 192)
        return trn::fail(TRN_ERR_INVALID, std::string(path) + ": not a little-endian kdtree.cache (or truncated)");
    std::vector<float> verts(n * 9), normals(n * 9), diffuse(n * 4), reflective(n * 4), reflectivity(n);
    float rec[48];
    for (uint64_t i = 0; i < n && r.ok; ++i) {
        r.raw(rec, sizeof rec);
        std::memcpy(&verts[i * 9], rec, 9 * sizeof(float));
        std::memcpy(&normals[i * 9], rec + 9, 9 * sizeof(float));
        std::memcpy(&diffuse[i * 4], rec + 22, 4 * sizeof(float));
        std::memcpy(&reflective[i * 4], rec + 30, 4 * sizeof(float));
        reflectivity[i] = rec[34];
    }
    float box[6];
    r.raw(box, sizeof box);
    const uint64_t num_nodes = r.u64();
    std::vector<uint64_t> nodes;
    if (r.ok && num_nodes < (1ull << 32) && num_nodes <= (file_size - 41 - n * 192) / 8) {
        nodes.resize(num_nodes);
        r.raw(nodes.data(), num_nodes * 8);
    } else {
        r.ok = false;
    }
    std::fclose(f);
    f = nullptr;
    if (!r.ok) return trn::fail(TRN_ERR_INVALID, std::string(path) + ": truncated kdtree.cache");
    trn_scene* sc = nullptr;
    const int32_t rc = trn_scene_create_ex(verts.data(), normals.data(), diffuse.data(), reflective.data(), reflectivity.data(),
                                           static_cast<uint32_t>(n), &sc);
    if (rc != TRN_OK) return rc;
    // the reference would now traverse whatever the file holds (main.cpp:147-152); here the cached tree must be the tree
    // of the cached triangles
    if (sc->tree.nodes != nodes || std::memcmp(sc->tree.box, box, sizeof box) != 0) {
        trn_scene_destroy(sc);
        return trn::fail(TRN_ERR_INVALID, std::string(path) + ": cached kd-tree does not belong to the cached triangles (stale or foreign kdtree.cache)");
    }
    *out = sc;
    return TRN_OK;
    TRN_GUARD_END
}
