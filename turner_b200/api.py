"""ctypes mirror of include/turner_b200.h -- the only way Python reaches the product.

Nothing here computes: every function forwards to libturner_b200.so (CUDA kernels for
sm_100a + C ABI). There is no CPU fallback; if the library is missing or no GPU is
usable the calls raise. PyTorch is optional plumbing (device buffers / streams /
torch.distributed in bench.py), not a dependency of this module.

Names follow the reference: TracerConfig fields (config.h:103-153), Camera
(lib/types.h:89-143), Light (lib/types.h:81-84), Stats (lib/stats.h:5-23).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libturner_b200.so")
MISS_ID = 0x40000000
PATHTRACER, RAYCASTER, RAYTRACER = 0, 1, 2

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


class TurnerError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("turner_b200 error %d: %s" % (code, msg))
        self.code = code


class Camera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("rot", C.c_float * 9), ("delta_x", C.c_float), ("delta_y", C.c_float)]


class Light(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("rgba", C.c_float * 4)]


class RenderConfig(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("max_depth", C.c_int32), ("mc_samples", C.c_int32),
        ("pixel_samples", C.c_int32), ("integrator", C.c_int32),
        ("bg_rgba", C.c_float * 4), ("max_visibility", C.c_float),
        ("num_lights", C.c_int32), ("light", Light),
        ("seed", C.c_uint64),
        ("sample_begin", C.c_int32), ("sample_stride", C.c_int32),
        ("shadow_intensity", C.c_float),
    ]


class Stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("prim_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("launches", C.c_uint64),
                ("ms_render", C.c_double), ("ms_trace", C.c_double), ("ms_shadow", C.c_double),
                ("ms_shade", C.c_double), ("ms_other", C.c_double),
                ("trace_launches", C.c_uint64), ("trace_queries", C.c_uint64), ("shadow_launches", C.c_uint64),
                ("trace_inner", C.c_uint64), ("trace_leaf_nodes", C.c_uint64), ("trace_tri_tests", C.c_uint64),
                ("shadow_inner", C.c_uint64), ("shadow_leaf_nodes", C.c_uint64), ("shadow_tri_tests", C.c_uint64),
                ("trace_actual_inner", C.c_uint64), ("trace_actual_leaf_nodes", C.c_uint64),
                ("trace_actual_tri_tests", C.c_uint64), ("shadow_actual_inner", C.c_uint64),
                ("shadow_actual_leaf_nodes", C.c_uint64), ("shadow_actual_tri_tests", C.c_uint64),
                ("ms_reduce", C.c_double), ("ms_d2h", C.c_double),
                ("trace_pooled", C.c_uint64 * 10), ("shadow_pooled", C.c_uint64 * 10), ("flat_records", C.c_uint64)]


class SceneInfo(C.Structure):
    _fields_ = [("num_triangles", C.c_uint64), ("num_nodes", C.c_uint64), ("kdtree_height", C.c_uint64),
                ("num_leaf_refs", C.c_uint64), ("num_cut_nodes", C.c_uint64), ("box", C.c_float * 6), ("build_ms", C.c_double),
                ("upload_ms", C.c_double)]


class LoadedScene(C.Structure):
    _fields_ = [("num_triangles", C.c_uint32), ("verts", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)),
                ("diffuse", C.POINTER(C.c_float)), ("reflective", C.POINTER(C.c_float)),
                ("reflectivity", C.POINTER(C.c_float)), ("has_camera", C.c_int32), ("cam_trafo4x4", C.c_float * 16),
                ("cam_hfov", C.c_float), ("cam_aspect", C.c_float), ("num_lights", C.c_int32), ("light", Light)]


EXPORTS = [
    "trn_last_error", "trn_device_count", "trn_scene_create", "trn_scene_create_ex", "trn_scene_destroy", "trn_scene_get_info",
    "trn_scene_get_nodes", "trn_intersect", "trn_primary_hits", "trn_render", "trn_render_device", "trn_render_multi",
    "trn_set_profiling", "trn_set_counting", "trn_intersect_counted", "trn_camera_setup", "trn_tonemap", "trn_write_p3", "trn_load_blend", "trn_load_soup", "trn_loaded_scene_free", "trn_scene_save_cache", "trn_scene_load_cache",
    "trn_comm_unique_id", "trn_comm_init_rank", "trn_comm_destroy", "trn_render_rank", "trn_render_async", "trn_wait",
    "trn_occluded", "trn_measure_gather_peak", "trn_scene_create_gpu",
]

_lib = None
_lib_override = None


def use_library(path):
    """development only (tools/ab_bench.py): load another build of the library; must be called before first use"""
    global _lib_override
    if _lib is not None:
        raise RuntimeError("library already loaded")
    _lib_override = path


def build(verbose=False):
    """compile libturner_b200.so + the CLI executables in-tree (nvcc, sm_100a)"""
    out = subprocess.run(["make", "-C", os.path.join(HERE, "csrc"), "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building turner_b200 failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no fallback path)" % LIB_PATH)
        L = C.CDLL(_lib_override or LIB_PATH)
        L.trn_last_error.restype = C.c_char_p
        L.trn_device_count.restype = C.c_int32
        L.trn_scene_create.argtypes = [_f32p, _f32p, _f32p, C.c_uint32, C.POINTER(C.c_void_p)]
        L.trn_scene_create_ex.argtypes = [_f32p, _f32p, _f32p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]
        L.trn_scene_create_ex.restype = C.c_int32
        L.trn_scene_create_gpu.argtypes = [_f32p, _f32p, _f32p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.POINTER(C.c_void_p)]
        L.trn_scene_create_gpu.restype = C.c_int32
        L.trn_scene_destroy.argtypes = [C.c_void_p]
        L.trn_scene_save_cache.argtypes = [C.c_void_p, C.c_char_p]
        L.trn_scene_load_cache.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.trn_scene_get_info.argtypes = [C.c_void_p, C.POINTER(SceneInfo)]
        L.trn_scene_get_nodes.argtypes = [C.c_void_p, _u64p]
        L.trn_intersect.argtypes = [C.c_void_p, C.c_int32, _f32p, _f32p, C.c_uint64, _u32p, _f32p]
        L.trn_primary_hits.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Camera), C.POINTER(RenderConfig), _u32p, _f32p]
        L.trn_render.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Camera), C.POINTER(RenderConfig), C.c_void_p,
                                 C.POINTER(Stats)]
        L.trn_render_device.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Camera), C.POINTER(RenderConfig), C.c_void_p,
                                        C.c_void_p, C.POINTER(Stats)]
        L.trn_render_multi.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.POINTER(Camera),
                                       C.POINTER(RenderConfig), C.c_void_p, C.POINTER(Stats)]
        L.trn_set_profiling.argtypes = [C.c_int32]
        L.trn_set_counting.argtypes = [C.c_int32]
        L.trn_intersect_counted.argtypes = [C.c_void_p, C.c_int32, _f32p, _f32p, C.c_uint64, _u32p, _f32p, _u64p]
        L.trn_camera_setup.argtypes = [_f32p, C.c_float, C.c_float, C.c_int32, C.POINTER(Camera), C.POINTER(C.c_int32)]
        L.trn_tonemap.argtypes = [_f32p, C.c_uint64, C.c_int32, C.c_float, C.c_int32, C.c_float, _f32p]
        L.trn_write_p3.restype = C.c_uint64
        L.trn_write_p3.argtypes = [_f32p, C.c_int32, C.c_int32, C.c_char_p, C.c_uint64]
        L.trn_load_blend.argtypes = [C.c_char_p, C.POINTER(LoadedScene)]
        L.trn_load_soup.argtypes = [C.c_char_p, C.POINTER(LoadedScene)]
        L.trn_load_soup.restype = C.c_int32
        L.trn_loaded_scene_free.argtypes = [C.POINTER(LoadedScene)]
        _u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
        L.trn_comm_unique_id.argtypes = [_u8p]
        L.trn_comm_init_rank.argtypes = [_u8p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        L.trn_comm_destroy.argtypes = [C.c_void_p]
        L.trn_comm_destroy.restype = None
        L.trn_render_rank.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Camera), C.POINTER(RenderConfig), C.c_void_p,
                                      C.POINTER(Stats)]
        L.trn_render_async.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Camera), C.POINTER(RenderConfig), C.c_void_p,
                                       C.POINTER(C.c_void_p)]
        L.trn_wait.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.trn_occluded.argtypes = [C.c_void_p, C.c_int32, _f32p, _f32p, _f32p, C.c_uint64, _u8p]
        L.trn_measure_gather_peak.argtypes = [C.c_int32, C.c_uint64, C.c_int32, C.POINTER(C.c_double)]
        for f in ("trn_comm_unique_id", "trn_comm_init_rank", "trn_render_rank", "trn_render_async", "trn_wait",
                  "trn_occluded", "trn_measure_gather_peak", "trn_scene_save_cache", "trn_scene_load_cache"):
            getattr(L, f).restype = C.c_int32
        for f in ("trn_scene_create", "trn_scene_get_info", "trn_scene_get_nodes", "trn_intersect", "trn_primary_hits",
                  "trn_intersect_counted",
                  "trn_render", "trn_render_device", "trn_render_multi", "trn_camera_setup", "trn_tonemap",
                  "trn_load_blend"):
            getattr(L, f).restype = C.c_int32
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise TurnerError(rc, lib().trn_last_error().decode())


def device_count():
    return lib().trn_device_count()


def set_profiling(on):
    lib().trn_set_profiling(1 if on else 0)


def set_counting(on):
    lib().trn_set_counting(1 if on else 0)


def camera_setup(trafo4x4, hfov, aspect, width):
    """Camera(trafo, aiCamera) + image height (lib/types.h:92-105, main.cpp:178-179)"""
    cam = Camera()
    h = C.c_int32()
    _check(lib().trn_camera_setup(np.ascontiguousarray(trafo4x4, np.float32).reshape(-1), hfov, aspect, width,
                                  C.byref(cam), C.byref(h)))
    return cam, h.value


def make_config(scene, width, max_depth=3, mc_samples=8, pixel_samples=1, integrator=PATHTRACER, bg=(0, 0, 0, 1),
                max_visibility=2.0, aspect=1.0, seed=1, sample_begin=0, sample_stride=1, shadow_intensity=0.5):
    """TracerConfig defaults of the reference's USAGE text (pathtracer.h:3-25); returns (Camera, RenderConfig)"""
    cam, height = camera_setup(scene["camera"]["trafo4x4"], scene["camera"]["hfov"], aspect, width)
    cfg = RenderConfig()
    cfg.width, cfg.height = width, height
    cfg.max_depth, cfg.mc_samples, cfg.pixel_samples, cfg.integrator = max_depth, mc_samples, pixel_samples, integrator
    cfg.bg_rgba = (C.c_float * 4)(*bg)
    cfg.max_visibility = max_visibility
    light = scene.get("light")
    cfg.num_lights = 1 if light else 0
    if light:
        cfg.light.pos = (C.c_float * 3)(*light["pos"])
        cfg.light.rgba = (C.c_float * 4)(*light["color"])
    cfg.seed = seed
    cfg.sample_begin, cfg.sample_stride = sample_begin, sample_stride
    cfg.shadow_intensity = shadow_intensity
    return cam, cfg


class Scene:
    """KDTree(triangles_from_scene(scene)) (main.cpp:25-82,156-157): triangles + kd-tree, host + device copies"""

    def __init__(self, vertices, normals, diffuse, reflective=None, reflectivity=None, builder="host", device=-1):
        """builder: "host" = the reference's tree, node for node (trn_scene_create); "gpu" = device build (trn_scene_create_gpu)"""
        v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 9)
        n = np.ascontiguousarray(normals, np.float32).reshape(-1, 9)
        d = np.ascontiguousarray(diffuse, np.float32).reshape(-1, 4)
        assert v.shape[0] == n.shape[0] == d.shape[0]
        self.h = C.c_void_p()
        if builder == "gpu":
            m = None if reflective is None else np.ascontiguousarray(reflective, np.float32).reshape(-1, 4)
            k = None if reflectivity is None else np.ascontiguousarray(reflectivity, np.float32).reshape(-1)
            _check(lib().trn_scene_create_gpu(v, n, d, m.ctypes.data if m is not None else None,
                                              k.ctypes.data if k is not None else None, v.shape[0], device, C.byref(self.h)))
        elif reflective is None and reflectivity is None:
            _check(lib().trn_scene_create(v, n, d, v.shape[0], C.byref(self.h)))
        else:
            m = np.ascontiguousarray(reflective, np.float32).reshape(-1, 4)
            k = np.ascontiguousarray(reflectivity, np.float32).reshape(-1)
            assert m.shape[0] == k.shape[0] == v.shape[0]
            _check(lib().trn_scene_create_ex(v, n, d, m.ctypes.data, k.ctypes.data, v.shape[0], C.byref(self.h)))
        self.info = SceneInfo()
        _check(lib().trn_scene_get_info(self.h, C.byref(self.info)))

    @classmethod
    def from_dict(cls, scene, builder="host", device=-1):
        return cls(scene["vertices"], scene["normals"], scene["diffuse"], scene.get("reflective"), scene.get("reflectivity"),
                   builder=builder, device=device)

    @classmethod
    def load_cache(cls, path):
        """a kdtree.cache (main.cpp:147-152) written by this library or by the reference; refused when stale"""
        self = cls.__new__(cls)
        self.h = C.c_void_p()
        _check(lib().trn_scene_load_cache(os.fsencode(path), C.byref(self.h)))
        self.info = SceneInfo()
        _check(lib().trn_scene_get_info(self.h, C.byref(self.info)))
        return self

    def save_cache(self, path):
        """write the tree in the reference's kdtree.cache layout (main.cpp:158-165)"""
        _check(lib().trn_scene_save_cache(self.h, os.fsencode(path)))

    def close(self):
        if getattr(self, "h", None):
            lib().trn_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def refresh_info(self):
        """re-read trn_scene_info (upload_ms is known after the first use on a device)"""
        _check(lib().trn_scene_get_info(self.h, C.byref(self.info)))
        return self.info

    @property
    def num_triangles(self):
        return self.info.num_triangles

    @property
    def num_nodes(self):
        return self.info.num_nodes

    @property
    def height(self):
        return self.info.kdtree_height

    def nodes(self):
        out = np.zeros(self.info.num_nodes, np.uint64)
        _check(lib().trn_scene_get_nodes(self.h, out))
        _check(lib().trn_scene_get_info(self.h, C.byref(self.info)))  # a device-built scene has its reference view now
        return out

    def intersect(self, origins, dirs, device=-1):
        """KDTreeIntersection::intersect for a batch of rays (lib/kdtree.cpp:515-578): ids (MISS_ID on miss), rst"""
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        ids = np.zeros(o.shape[0], np.uint32)
        rst = np.zeros((o.shape[0], 3), np.float32)
        _check(lib().trn_intersect(self.h, device, o, d, o.shape[0], ids, rst))
        return ids, rst

    def intersect_counted(self, origins, dirs, device=-1):
        """intersect() + (inner visits, reference leaf nodes, triangle tests) of the batch"""
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        ids = np.zeros(o.shape[0], np.uint32)
        rst = np.zeros((o.shape[0], 3), np.float32)
        cnt = np.zeros(6, np.uint64)
        _check(lib().trn_intersect_counted(self.h, device, o, d, o.shape[0], ids, rst, cnt))
        return ids, rst, cnt

    def primary_hits(self, cam, cfg, device=-1):
        n = cfg.width * cfg.height * cfg.pixel_samples
        ids = np.zeros(n, np.uint32)
        rst = np.zeros((n, 3), np.float32)
        _check(lib().trn_primary_hits(self.h, device, C.byref(cam), C.byref(cfg), ids, rst))
        shape = (cfg.height, cfg.width, cfg.pixel_samples)
        return ids.reshape(shape), rst.reshape(shape + (3,))

    def render(self, cam, cfg, device=-1, out=None):
        """the render loop (main.cpp:187-236) up to the per-pixel SUM over pixel samples; host buffer out"""
        if out is None:
            out = np.zeros((cfg.height, cfg.width, 4), np.float32)
        st = Stats()
        _check(lib().trn_render(self.h, device, C.byref(cam), C.byref(cfg), out.ctypes.data, C.byref(st)))
        return out, st

    def render_device(self, cam, cfg, d_accum_ptr, stream_ptr=0, device=-1, want_stats=True):
        """accumulate into a caller-owned device buffer (e.g. torch tensor .data_ptr()) on a caller stream"""
        st = Stats()
        _check(lib().trn_render_device(self.h, device, C.byref(cam), C.byref(cfg), C.c_void_p(d_accum_ptr),
                                       C.c_void_p(stream_ptr), C.byref(st) if want_stats else None))
        return st

    def occluded(self, origins, dirs, tmax, device=-1):
        """shadow predicate of pathtracer.cpp:49-53 for a batch of rays: True where an accepted hit has 0 <= r <= tmax"""
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(tmax, np.float32).reshape(-1)
        assert o.shape[0] == d.shape[0] == t.shape[0]
        out = np.zeros(o.shape[0], np.uint8)
        _check(lib().trn_occluded(self.h, device, o, d, t, o.shape[0], out))
        return out.astype(bool)

    def render_rank(self, comm, cam, cfg, out=None, out_ptr=None):
        """this rank's share of a process-per-GPU job + the ncclReduce onto rank 0 (+ D2H into `out` on rank 0)"""
        st = Stats()
        ptr = out_ptr if out_ptr is not None else (out.ctypes.data if out is not None else None)
        _check(lib().trn_render_rank(self.h, comm.h, C.byref(cam), C.byref(cfg), C.c_void_p(ptr) if ptr else None, C.byref(st)))
        return st

    def render_async(self, cam, cfg, out, device=-1):
        """start a frame; returns a job whose wait() gives the Stats once `out` (host array) holds the image"""
        job = C.c_void_p()
        _check(lib().trn_render_async(self.h, device, C.byref(cam), C.byref(cfg), C.c_void_p(out.ctypes.data), C.byref(job)))
        return Job(job, out)

    def render_multi(self, cam, cfg, devices):
        out = np.zeros((cfg.height, cfg.width, 4), np.float32)
        st = Stats()
        devs = (C.c_int32 * len(devices))(*devices)
        _check(lib().trn_render_multi(self.h, devs, len(devices), C.byref(cam), C.byref(cfg), out.ctypes.data,
                                      C.byref(st)))
        return out, st


class Job:
    def __init__(self, h, out):
        self.h, self.out = h, out

    def wait(self):
        st = Stats()
        h, self.h = self.h, None
        _check(lib().trn_wait(h, C.byref(st)))
        return self.out, st


class Comm:
    """one rank of a process-per-GPU job: NCCL communicator owned by the library (trn_comm_*)"""

    def __init__(self, unique_id, nranks, rank, device):
        self.h = C.c_void_p()
        self.nranks, self.rank, self.device = nranks, rank, device
        _check(lib().trn_comm_init_rank(np.ascontiguousarray(unique_id, np.uint8), nranks, rank, device, C.byref(self.h)))

    @staticmethod
    def unique_id():
        out = np.zeros(128, np.uint8)
        _check(lib().trn_comm_unique_id(out))
        return out

    def close(self):
        if getattr(self, "h", None):
            lib().trn_comm_destroy(self.h)
            self.h = None


def measure_gather_peak(set_bytes, mode=1, device=-1):
    g = C.c_double()
    _check(lib().trn_measure_gather_peak(device, set_bytes, mode, C.byref(g)))
    return g.value


def copy_config(cfg):
    c = RenderConfig()
    C.memmove(C.byref(c), C.byref(cfg), C.sizeof(RenderConfig))
    return c


def tonemap(rgba_sum, pixel_samples, exposure=1.0, gamma_enabled=True, inverse_gamma=0.454545, out=None):
    a = np.ascontiguousarray(rgba_sum, np.float32)
    if out is None:
        out = np.zeros_like(a)
    _check(lib().trn_tonemap(a.reshape(-1), a.size // 4, pixel_samples, exposure, 1 if gamma_enabled else 0,
                             inverse_gamma, out.reshape(-1)))
    return out


def write_p3(rgba):
    a = np.ascontiguousarray(rgba, np.float32)
    h, w = a.shape[0], a.shape[1]
    n = lib().trn_write_p3(a.reshape(-1), w, h, None, 0)
    buf = C.create_string_buffer(n)
    lib().trn_write_p3(a.reshape(-1), w, h, buf, n)
    return buf.raw[:n].decode()


def save_soup(scene, path):
    """write a scene dict in the neutral triangle-soup text format the CLI reads"""
    f9 = lambda a: " ".join("%.9g" % float(x) for x in a)
    with open(path, "w") as f:
        f.write("# turner_b200 triangle soup: %s\n" % scene.get("name", ""))
        f.write("camera %s %.9g\n" % (f9(scene["camera"]["trafo4x4"]), scene["camera"]["hfov"]))
        if scene.get("light"):
            f.write("light %s %s\n" % (f9(scene["light"]["pos"]), f9(scene["light"]["color"])))
        nt = len(scene["vertices"])
        mir = scene.get("reflective", np.zeros((nt, 4), np.float32))
        ref = scene.get("reflectivity", np.zeros(nt, np.float32))
        for v, n, d, m, k in zip(scene["vertices"], scene["normals"], scene["diffuse"], mir, ref):
            f.write("tri %s %s %s %s %.9g\n" % (f9(v), f9(n), f9(d), f9(m), float(k)))


def load_blend(path):
    """.blend (or, with any other extension, the neutral soup format) -> scene dict (same shape as turner_b200.scenes)"""
    ls = LoadedScene()
    loader = lib().trn_load_blend if path.endswith(".blend") else lib().trn_load_soup
    _check(loader(path.encode(), C.byref(ls)))
    try:
        n = ls.num_triangles
        sc = {
            "name": os.path.basename(path),
            "vertices": np.ctypeslib.as_array(ls.verts, (n, 9)).copy(),
            "normals": np.ctypeslib.as_array(ls.normals, (n, 9)).copy(),
            "diffuse": np.ctypeslib.as_array(ls.diffuse, (n, 4)).copy(),
            "reflective": np.ctypeslib.as_array(ls.reflective, (n, 4)).copy(),
            "reflectivity": np.ctypeslib.as_array(ls.reflectivity, (n,)).copy(),
            "camera": {"trafo4x4": list(ls.cam_trafo4x4), "hfov": float(ls.cam_hfov)} if ls.has_camera else None,
            "light": {"pos": list(ls.light.pos), "color": list(ls.light.rgba)} if ls.num_lights else None,
        }
    finally:
        lib().trn_loaded_scene_free(C.byref(ls))
    return sc
