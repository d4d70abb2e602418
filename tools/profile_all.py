#!/usr/bin/env python3
"""for ncu: the gather micro-benchmark (L1 / L2 sets) followed by one render step of a bench workload (development aid)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from turner_b200 import api

name = sys.argv[1] if len(sys.argv) > 1 else "mesh1m"
print("gather l1 %.0f GB/s" % api.measure_gather_peak(32 << 10, 0, 0), flush=True)
print("gather l2 %.0f GB/s" % api.measure_gather_peak(96 << 20, 1, 0), flush=True)
w = bench.WORKLOADS[name]
sc = bench.load_scene(name)
scene = api.Scene.from_dict(sc)
cam, cfg = api.make_config(sc, w["width"], max_depth=w["max_depth"], mc_samples=w["mc_samples"], pixel_samples=w["pixel_samples"], seed=1)
cfg.sample_begin, cfg.sample_stride = 0, w["pixel_samples"]
img, st = scene.render(cam, cfg)
print(name, "rays", st.rays, "shadow", st.shadow_rays, "ms %.2f" % st.ms_render, "launches", st.launches, flush=True)
