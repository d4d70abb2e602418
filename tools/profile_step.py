#!/usr/bin/env python3
"""one or two render steps of a bench workload, for ncu captures (development aid)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from turner_b200 import api

name = sys.argv[1] if len(sys.argv) > 1 else "mesh1m"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = bench.WORKLOADS[name]
sc = bench.load_scene(name)
scene = api.Scene.from_dict(sc, builder="gpu" if name == "mesh1m" else "host")  # the trees bench.py renders on
cam, cfg = api.make_config(sc, w["width"], max_depth=w["max_depth"], mc_samples=w["mc_samples"], pixel_samples=w["pixel_samples"], seed=1)
for i in range(steps):
    cfg.sample_begin, cfg.sample_stride = i, w["pixel_samples"]
    img, st = scene.render(cam, cfg)
    print(name, "step", i, "rays", st.rays, "shadow", st.shadow_rays, "ms %.2f" % st.ms_render, "launches", st.launches, flush=True)
