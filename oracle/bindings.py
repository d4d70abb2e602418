"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU oracle (libturner_oracle.so)
and for oracle/_ref (the reference's own sources compiled here, see build_ref.sh).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module. The product path (turner_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MISS = 1 << 30  # OptionalId miss value, /root/reference/lib/kdtree.h:156-161

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def build(force=False):
    """compile the oracle (and oracle/_ref when /root/reference is present)"""
    so = os.path.join(HERE, "libturner_oracle.so")
    src = os.path.join(HERE, "turner_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "libturner_oracle.so"])
    ref_so = os.path.join(HERE, "_ref", "libturner_ref_pathtracer.so")
    if os.path.isdir(os.environ.get("TURNER_REFERENCE", "/root/reference")) and (force or not os.path.exists(ref_so)):
        subprocess.check_call([os.path.join(HERE, "build_ref.sh")])


class RenderCfg(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("max_depth", C.c_int32), ("mc_samples", C.c_int32), ("pixel_samples", C.c_int32), ("num_threads", C.c_int32),
        ("integrator", C.c_int32), ("rng_mode", C.c_int32),
        ("seed", C.c_uint64),
        ("sample_begin", C.c_int32), ("sample_stride", C.c_int32),
        ("bg", C.c_float * 4),
        ("max_visibility", C.c_float),
        ("num_lights", C.c_int32),
        ("light_pos", C.c_float * 3),
        ("light_color", C.c_float * 4),
        ("cam_pos", C.c_float * 3),
        ("cam_rot", C.c_float * 9),
        ("delta_x", C.c_float), ("delta_y", C.c_float),
        ("shadow_intensity", C.c_float),
    ]


class RenderStats(C.Structure):
    _fields_ = [("num_rays", C.c_uint64), ("num_prim_rays", C.c_uint64), ("num_shadow_rays", C.c_uint64),
                ("runtime_ms", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(HERE, "libturner_oracle.so"))
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_create.argtypes = [_f32p, _f32p, _f32p, C.c_uint32]
        L.orc_scene_create_ex.restype = C.c_void_p
        L.orc_scene_create_ex.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_uint32]
        L.orc_scene_create_prebuilt.restype = C.c_void_p
        L.orc_scene_create_prebuilt.argtypes = [_f32p, _f32p, _f32p, C.c_uint32, _u64p, C.c_uint64, _f32p]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_scene_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                     _f32p, C.POINTER(C.c_double)]
        L.orc_scene_nodes.argtypes = [C.c_void_p, _u64p]
        L.orc_triangle_fields.argtypes = [C.c_void_p, C.c_uint32, _f32p]
        L.orc_intersect.argtypes = [C.c_void_p, _f32p, _f32p, C.c_uint64, C.c_int32, _u32p, _f32p, C.c_void_p]
        L.orc_ray_box.restype = C.c_int
        L.orc_ray_box.argtypes = [_f32p, _f32p, _f32p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_ray_triangle.restype = C.c_int
        L.orc_ray_triangle.argtypes = [_f32p, _f32p, _f32p, _f32p]
        L.orc_clipped_box.argtypes = [_f32p, _f32p, _f32p]
        L.orc_xorshift_float.argtypes = [C.c_uint64, C.c_uint64, _f32p]
        L.orc_xorshift_u64.argtypes = [C.c_uint64, C.c_uint64, _u64p]
        L.orc_node_state.restype = C.c_uint64
        L.orc_node_state.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_hemisphere.argtypes = [C.c_uint64, _f32p]
        L.orc_frame_apply.argtypes = [_f32p, _f32p, _f32p]
        L.orc_camera_setup.argtypes = [_f32p, C.c_float, C.c_float, C.c_int32, _f32p, _f32p, _f32p,
                                       C.POINTER(C.c_int32)]
        L.orc_primary_dirs.argtypes = [C.POINTER(RenderCfg), _f32p]
        L.orc_render.restype = C.c_int
        L.orc_render.argtypes = [C.c_void_p, C.POINTER(RenderCfg), _f32p, C.c_void_p, C.POINTER(RenderStats)]
        L.orc_tonemap.argtypes = [_f32p, C.c_uint64, C.c_int32, C.c_float, C.c_int32, C.c_float, _f32p]
        L.orc_write_p3.restype = C.c_uint64
        L.orc_write_p3.argtypes = [_f32p, C.c_int32, C.c_int32, C.c_char_p, C.c_uint64]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def camera_setup(trafo4x4, hfov, aspect, width):
    pos = np.zeros(3, np.float32)
    rot = np.zeros(9, np.float32)
    dxy = np.zeros(2, np.float32)
    h = C.c_int32(0)
    lib().orc_camera_setup(_f32(trafo4x4).reshape(-1), hfov, aspect, width, pos, rot, dxy, C.byref(h))
    return pos, rot, float(dxy[0]), float(dxy[1]), h.value


def make_cfg(scene, width, max_depth=3, mc_samples=8, pixel_samples=1, num_threads=1, integrator=0, rng_mode=0,
             seed=1, sample_begin=0, sample_stride=1, bg=(0, 0, 0, 1), max_visibility=2.0, aspect=1.0,
             shadow_intensity=0.5):
    """scene: dict with 'camera' {'trafo4x4','hfov'} and 'light' (or None), as in tests/golden/*.json"""
    cam = scene["camera"]
    pos, rot, dx, dy, height = camera_setup(cam["trafo4x4"], cam["hfov"], aspect, width)
    cfg = RenderCfg()
    cfg.width, cfg.height = width, height
    cfg.max_depth, cfg.mc_samples, cfg.pixel_samples, cfg.num_threads = max_depth, mc_samples, pixel_samples, num_threads
    cfg.integrator, cfg.rng_mode, cfg.seed = integrator, rng_mode, seed
    cfg.sample_begin, cfg.sample_stride = sample_begin, sample_stride
    cfg.bg = (C.c_float * 4)(*bg)
    cfg.max_visibility = max_visibility
    light = scene.get("light")
    cfg.num_lights = 1 if light else 0
    if light:
        cfg.light_pos = (C.c_float * 3)(*light["pos"])
        cfg.light_color = (C.c_float * 4)(*light["color"])
    cfg.cam_pos = (C.c_float * 3)(*pos)
    cfg.cam_rot = (C.c_float * 9)(*rot)
    cfg.delta_x, cfg.delta_y = dx, dy
    cfg.shadow_intensity = shadow_intensity
    return cfg


class OracleScene:
    def __init__(self, verts, normals, diffuse, nodes=None, box=None, reflective=None, reflectivity=None):
        self.verts = _f32(verts).reshape(-1, 9)
        self.normals = _f32(normals).reshape(-1, 9)
        self.diffuse = _f32(diffuse).reshape(-1, 4)
        n = self.verts.shape[0]
        if reflective is not None:
            assert nodes is None
            self.h = lib().orc_scene_create_ex(self.verts, self.normals, self.diffuse, _f32(reflective).reshape(-1, 4),
                                               _f32(reflectivity).reshape(-1), n)
        elif nodes is None:
            self.h = lib().orc_scene_create(self.verts, self.normals, self.diffuse, n)
        else:
            nodes = np.ascontiguousarray(nodes, dtype=np.uint64)
            self.h = lib().orc_scene_create_prebuilt(self.verts, self.normals, self.diffuse, n, nodes, nodes.size,
                                                     _f32(box).reshape(-1))
        nn, hh, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
        b = np.zeros(6, np.float32)
        ms = C.c_double()
        lib().orc_scene_info(self.h, C.byref(nn), C.byref(hh), C.byref(nt), b, C.byref(ms))
        self.num_nodes, self.height, self.num_tris, self.box, self.build_ms = nn.value, hh.value, nt.value, b, ms.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_scene_destroy(self.h)
            self.h = None

    def nodes(self):
        out = np.zeros(self.num_nodes, np.uint64)
        lib().orc_scene_nodes(self.h, out)
        return out

    def triangle_fields(self, i):
        out = np.zeros(48, np.float32)
        lib().orc_triangle_fields(self.h, i, out)
        return out

    def intersect(self, o, d, mode=0, counters=False):
        o = _f32(o).reshape(-1, 3)
        d = _f32(d).reshape(-1, 3)
        n = o.shape[0]
        ids = np.zeros(n, np.uint32)
        rst = np.zeros((n, 3), np.float32)
        cnt = np.zeros(8, np.uint64)
        lib().orc_intersect(self.h, o, d, n, mode, ids, rst, cnt.ctypes.data if counters else None)
        return (ids, rst, cnt) if counters else (ids, rst)

    def render(self, cfg, want_sumsq=False):
        n = cfg.width * cfg.height * 4
        out = np.zeros(n, np.float32)
        sq = np.zeros(n, np.float32) if want_sumsq else None
        st = RenderStats()
        rc = lib().orc_render(self.h, C.byref(cfg), out, sq.ctypes.data if want_sumsq else None, C.byref(st))
        assert rc == 0
        shape = (cfg.height, cfg.width, 4)
        return out.reshape(shape), (sq.reshape(shape) if want_sumsq else None), st


def primary_dirs(cfg):
    out = np.zeros((cfg.height, cfg.width, cfg.pixel_samples, 3), np.float32)
    lib().orc_primary_dirs(C.byref(cfg), out.reshape(-1))
    return out


def tonemap(sum_rgba, pps, exposure=1.0, gamma_enabled=True, inverse_gamma=0.454545):
    a = _f32(sum_rgba)
    out = np.zeros_like(a)
    lib().orc_tonemap(a.reshape(-1), a.size // 4, pps, exposure, 1 if gamma_enabled else 0, inverse_gamma,
                      out.reshape(-1))
    return out


def write_p3(rgba):
    a = _f32(rgba)
    h, w = a.shape[0], a.shape[1]
    n = lib().orc_write_p3(a.reshape(-1), w, h, None, 0)
    buf = C.create_string_buffer(n)
    lib().orc_write_p3(a.reshape(-1), w, h, buf, n)
    return buf.raw[:n].decode()


# --------------------------------------------------------------------- oracle/_ref
class RefCamera(C.Structure):
    _fields_ = [("trafo4x4", C.c_float * 16), ("hfov", C.c_float), ("aspect", C.c_float)]


class RefConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("max_depth", C.c_int32), ("mc_samples", C.c_int32),
                ("pixel_samples", C.c_int32), ("num_threads", C.c_int32), ("gamma_enabled", C.c_int32),
                ("bg", C.c_float * 4), ("exposure", C.c_float), ("inverse_gamma", C.c_float),
                ("max_visibility", C.c_float), ("num_lights", C.c_int32), ("light_pos", C.c_float * 3),
                ("light_color", C.c_float * 4), ("shadow_intensity", C.c_float),
                ("row_begin", C.c_int32), ("row_stride", C.c_int32)]


class RefStats(C.Structure):
    _fields_ = [("num_rays", C.c_uint64), ("num_prim_rays", C.c_uint64), ("runtime_ms", C.c_uint64),
                ("width", C.c_uint64), ("height", C.c_uint64)]


_ref = {}


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libturner_ref_pathtracer.so"))


def ref_lib(kind="pathtracer"):
    if kind not in _ref:
        build()
        L = C.CDLL(os.path.join(HERE, "_ref", "libturner_ref_%s.so" % kind))
        L.ref_scene_create.restype = C.c_void_p
        L.ref_scene_create.argtypes = [_f32p, _f32p, _f32p, C.c_uint32]
        L.ref_scene_create_ex.restype = C.c_void_p
        L.ref_scene_create_ex.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_uint32]
        L.ref_scene_create_prebuilt.restype = C.c_void_p
        L.ref_scene_create_prebuilt.argtypes = [_f32p, _f32p, _f32p, C.c_uint32, _u64p, C.c_uint64, _f32p]
        L.ref_scene_destroy.argtypes = [C.c_void_p]
        L.ref_cache_write.restype = C.c_int
        L.ref_cache_write.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_cache_read.restype = C.c_void_p
        L.ref_cache_read.argtypes = [C.c_char_p]
        L.ref_scene_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                     _f32p]
        L.ref_scene_nodes.argtypes = [C.c_void_p, _u64p]
        L.ref_triangle_fields.argtypes = [C.c_void_p, C.c_uint32, _f32p]
        L.ref_intersect.argtypes = [C.c_void_p, _f32p, _f32p, C.c_uint64, _u32p, _f32p]
        L.ref_intersect_ray_box.restype = C.c_int
        L.ref_intersect_ray_box.argtypes = [_f32p, _f32p, _f32p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ref_primary_dirs.argtypes = [C.POINTER(RefCamera), C.c_int32, C.c_int32, _f32p, C.c_void_p,
                                       C.POINTER(C.c_int32)]
        L.ref_render.restype = C.c_int
        L.ref_render.argtypes = [C.c_void_p, C.POINTER(RefCamera), C.POINTER(RefConfig), _f32p, C.c_void_p, C.c_void_p,
                                 C.POINTER(RefStats)]
        L.ref_write_p3.restype = C.c_uint64
        L.ref_write_p3.argtypes = [_f32p, C.c_int32, C.c_int32, C.c_char_p, C.c_uint64]
        L.ref_xorshift_float.argtypes = [C.c_uint64, C.c_uint64, _f32p]
        L.ref_xorshift_u64.argtypes = [C.c_uint64, C.c_uint64, _u64p]
        L.ref_hemisphere.argtypes = [C.c_uint64, _f32p]
        L.ref_tonemap.argtypes = [_f32p, C.c_uint64, C.c_float, C.c_int, C.c_float, _f32p]
        L.ref_usage.restype = C.c_char_p
        _ref[kind] = L
    return _ref[kind]


def ref_camera(scene, aspect=1.0):
    rc = RefCamera()
    rc.trafo4x4 = (C.c_float * 16)(*scene["camera"]["trafo4x4"])
    rc.hfov = scene["camera"]["hfov"]
    rc.aspect = aspect
    return rc


def ref_config(scene, width, max_depth=3, mc_samples=8, pixel_samples=1, num_threads=1, bg=(0, 0, 0, 1),
               exposure=1.0, gamma_enabled=True, inverse_gamma=0.454545, max_visibility=2.0, shadow_intensity=0.5):
    c = RefConfig()
    c.width, c.max_depth, c.mc_samples, c.pixel_samples, c.num_threads = width, max_depth, mc_samples, pixel_samples, num_threads
    c.gamma_enabled = 1 if gamma_enabled else 0
    c.bg = (C.c_float * 4)(*bg)
    c.exposure, c.inverse_gamma, c.max_visibility = exposure, inverse_gamma, max_visibility
    c.shadow_intensity = shadow_intensity
    light = scene.get("light")
    c.num_lights = 1 if light else 0
    if light:
        c.light_pos = (C.c_float * 3)(*light["pos"])
        c.light_color = (C.c_float * 4)(*light["color"])
    return c


class RefScene:
    """the reference's own KDTree (+ trace()) behind oracle/ref_driver.cpp"""

    def __init__(self, verts, normals, diffuse, kind="pathtracer", nodes=None, box=None, reflective=None,
                 reflectivity=None):
        self.L = ref_lib(kind)
        self.verts = _f32(verts).reshape(-1, 9)
        self.normals = _f32(normals).reshape(-1, 9)
        self.diffuse = _f32(diffuse).reshape(-1, 4)
        n = self.verts.shape[0]
        if reflective is not None:
            assert nodes is None
            self.h = self.L.ref_scene_create_ex(self.verts, self.normals, self.diffuse, _f32(reflective).reshape(-1, 4),
                                                _f32(reflectivity).reshape(-1), n)
        elif nodes is None:
            self.h = self.L.ref_scene_create(self.verts, self.normals, self.diffuse, n)
        else:
            nodes = np.ascontiguousarray(nodes, dtype=np.uint64)
            self.h = self.L.ref_scene_create_prebuilt(self.verts, self.normals, self.diffuse, n, nodes, nodes.size,
                                                      _f32(box).reshape(-1))
        nn, hh, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
        b = np.zeros(6, np.float32)
        self.L.ref_scene_info(self.h, C.byref(nn), C.byref(hh), C.byref(nt), b)
        self.num_nodes, self.height, self.num_tris, self.box = nn.value, hh.value, nt.value, b

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_scene_destroy(self.h)
            self.h = None

    @classmethod
    def from_cache(cls, path, kind="pathtracer"):
        """main.cpp:147-152: KDTree loaded from a kdtree.cache through the reference's serialize()"""
        self = cls.__new__(cls)
        self.L = ref_lib(kind)
        self.h = self.L.ref_cache_read(os.fsencode(path))
        if not self.h:
            raise RuntimeError("the reference's KDTree::serialize() could not read %s" % path)
        nn, hh, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
        b = np.zeros(6, np.float32)
        self.L.ref_scene_info(self.h, C.byref(nn), C.byref(hh), C.byref(nt), b)
        self.num_nodes, self.height, self.num_tris, self.box = nn.value, hh.value, nt.value, b
        return self

    def write_cache(self, path):
        """main.cpp:158-165: the reference's KDTree::serialize() into a PortableBinary archive"""
        if self.L.ref_cache_write(self.h, os.fsencode(path)) != 0:
            raise RuntimeError("cannot write %s" % path)

    def nodes(self):
        out = np.zeros(self.num_nodes, np.uint64)
        self.L.ref_scene_nodes(self.h, out)
        return out

    def triangle_fields(self, i):
        out = np.zeros(48, np.float32)
        self.L.ref_triangle_fields(self.h, i, out)
        return out

    def intersect(self, o, d):
        o = _f32(o).reshape(-1, 3)
        d = _f32(d).reshape(-1, 3)
        n = o.shape[0]
        ids = np.zeros(n, np.uint32)
        rst = np.zeros((n, 3), np.float32)
        self.L.ref_intersect(self.h, o, d, n, ids, rst)
        return ids, rst

    def render(self, cam, cfg, want_sumsq=False, want_final=False):
        height = int(np.float32(cfg.width) / np.float32(cam.aspect))
        n = cfg.width * height * 4
        out = np.zeros(n, np.float32)
        sq = np.zeros(n, np.float32) if want_sumsq else None
        fin = np.zeros(n, np.float32) if want_final else None
        st = RefStats()
        rc = self.L.ref_render(self.h, C.byref(cam), C.byref(cfg), out, sq.ctypes.data if want_sumsq else None,
                               fin.ctypes.data if want_final else None, C.byref(st))
        assert rc == 0 and st.height == height
        shape = (height, cfg.width, 4)
        return (out.reshape(shape), sq.reshape(shape) if want_sumsq else None,
                fin.reshape(shape) if want_final else None, st)


def ref_primary_dirs(cam, width, pps, kind="pathtracer"):
    L = ref_lib(kind)
    pos = np.zeros(3, np.float32)
    h = C.c_int32()
    L.ref_primary_dirs(C.byref(cam), width, pps, pos, None, C.byref(h))
    dirs = np.zeros((h.value, width, pps, 3), np.float32)
    L.ref_primary_dirs(C.byref(cam), width, pps, pos, dirs.ctypes.data, C.byref(h))
    return pos, dirs


def ref_write_p3(rgba, kind="pathtracer"):
    L = ref_lib(kind)
    a = _f32(rgba)
    h, w = a.shape[0], a.shape[1]
    n = L.ref_write_p3(a.reshape(-1), w, h, None, 0)
    buf = C.create_string_buffer(n)
    L.ref_write_p3(a.reshape(-1), w, h, buf, n)
    return buf.raw[:n].decode()
