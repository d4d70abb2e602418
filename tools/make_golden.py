#!/usr/bin/env python3
"""Generates tests/golden/reference_outputs.npz from the REFERENCE ITSELF (oracle/_ref = the reference's kdtree.cpp +
pathtracer.cpp / raycaster.cpp / raytracer.cpp compiled in place, see oracle/build_ref.sh). Run in the build container
(needs /root/reference):   python tools/make_golden.py
The fixtures pin the oracle (tests/test_golden_fixtures.py, CPU) and the CUDA path (tests/test_gpu_parity.py) to
outputs of the reference's own code even where oracle/_ref is not available."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import bindings as ob
from turner_b200 import scenes

assert ob.ref_available(), "oracle/_ref is not built"
out = {}
for name, bg in (("cornell_box", (0, 0, 0, 1)), ("colored_cube", (0.1, 0.2, 0.3, 1)), ("furnace_test", (1, 1, 1, 1))):
    sc = scenes.fixture(name)
    kw = dict(reflective=sc["reflective"], reflectivity=sc["reflectivity"])
    cam = ob.ref_camera(sc)
    r = ob.RefScene(sc["vertices"], sc["normals"], sc["diffuse"], kind="pathtracer", **kw)
    out[name + "/nodes"] = r.nodes()
    out[name + "/height"] = np.array([r.height])
    # pathtracer.cpp, reference streams, 1 thread: W 24, D 3, m 2, pps 2
    img, sq, fin, st = r.render(cam, ob.ref_config(sc, 24, 3, 2, 2, bg=bg), want_sumsq=True, want_final=True)
    out[name + "/pt_sum"], out[name + "/pt_final"] = img, fin
    out[name + "/pt_rays"] = np.array([st.num_rays, st.num_prim_rays])
    # primary rays + their closest hits (KDTreeIntersection::intersect), W 48, pps 1
    pos, dirs = ob.ref_primary_dirs(cam, 48, 1)
    ids, rst = r.intersect(np.tile(pos, (dirs.size // 3, 1)), dirs.reshape(-1, 3))
    out[name + "/cam_pos"], out[name + "/prim_dirs"], out[name + "/prim_ids"], out[name + "/prim_rst"] = pos, dirs, ids, rst
    rc = ob.RefScene(sc["vertices"], sc["normals"], sc["diffuse"], kind="raycaster", **kw)
    img, _, fin, st = rc.render(cam, ob.ref_config(sc, 48, bg=bg, max_visibility=2.0), want_final=True)
    out[name + "/rc_sum"], out[name + "/rc_rays"] = img, np.array([st.num_rays, st.num_prim_rays])
    out[name + "/rc_p3"] = np.frombuffer(ob.ref_write_p3(fin, kind="raycaster").encode(), np.uint8)
    if sc["light"]:
        rt = ob.RefScene(sc["vertices"], sc["normals"], sc["diffuse"], kind="raytracer", **kw)
        img, _, _, st = rt.render(cam, ob.ref_config(sc, 48, max_depth=4, bg=bg, shadow_intensity=0.5, pixel_samples=2))
        out[name + "/rt_sum"], out[name + "/rt_rays"] = img, np.array([st.num_rays, st.num_prim_rays])
x = np.zeros(64, np.uint64)
ob.ref_lib().ref_xorshift_u64(42, 64, x)
out["xorshift64star_u64_seed42"] = x
f = np.zeros(64, np.float32)
ob.ref_lib().ref_xorshift_float(4, 64, f)
out["xorshift64star_float_seed4"] = f
h = np.zeros(4 * 64, np.float32)
ob.ref_lib().ref_hemisphere(64, h)
out["hemisphere_first64"] = h.reshape(64, 4)
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_outputs.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes,", len(out), "arrays")
