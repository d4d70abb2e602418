#!/usr/bin/env python3
"""README.md:22-36 replay: `pathtracer cornell_box.blend --width 320 --max-depth 3 --monte-carlo-samples 1 --pixel-samples 8`
(BASELINE config 1; the reference reports 2 632 399 rays, 819 200 primary, 129 419 rays/s on unstated hardware, 1 thread,
as-shipped flags). Times the reference's own code compiled here (as shipped and -O2 -DNDEBUG), and the GPU path if a GPU
is present. Development/report aid; prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import bindings as ob
from turner_b200 import api, scenes

sc = scenes.fixture("cornell_box")
out = {"config": "cornell_box -w 320 -d 3 -m 1 -p 8", "readme": {"rays": 2632399, "prim": 819200, "rays_per_s": 129419}}
cam = ob.ref_camera(sc)
for kind, threads in (("pathtracer_shipped", 1), ("pathtracer", 1), ("pathtracer", os.cpu_count() or 1)):
    if not os.path.exists(os.path.join(ob.HERE, "_ref", "libturner_ref_%s.so" % kind)):
        continue
    r = ob.RefScene(sc["vertices"], sc["normals"], sc["diffuse"], kind=kind)
    cfg = ob.ref_config(sc, 320, 3, 1, 8, num_threads=threads)
    t = time.perf_counter()
    _, _, _, st = r.render(cam, cfg)
    dt = time.perf_counter() - t
    out["%s_t%d" % (kind, threads)] = {"rays": st.num_rays, "prim": st.num_prim_rays, "seconds": round(dt, 3),
                                       "rays_per_s": int(st.num_rays / dt)}
if api.device_count() > 0:
    p = api.Scene.from_dict(sc)
    cam2, cfg2 = api.make_config(sc, 320, max_depth=3, mc_samples=1, pixel_samples=8, seed=1)
    p.render(cam2, cfg2)
    img, st = p.render(cam2, cfg2)
    out["b200"] = {"rays": int(st.rays), "prim": int(st.prim_rays), "ms": st.ms_render, "rays_per_s": int(st.rays / st.ms_render * 1e3)}
print(json.dumps(out))
