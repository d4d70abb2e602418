#!/usr/bin/env python3
"""bisect a pooled-kernel problem on the 1M mesh: intersect / primary hits / renders of growing size, with progress prints"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from turner_b200 import api, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 288
sc = scenes.cubesphere(n)
t = time.time()
p = api.Scene.from_dict(sc)
print("scene built", n, "%.1fs" % (time.time() - t), flush=True)
for cnt in (1000, 100000, 2000000):
    ro, rd = scenes.random_rays(sc, cnt, seed=3, inside=True)
    t = time.time()
    i1, r1 = p.intersect(ro, rd)
    print("intersect", cnt, "hits", int((i1 != 0x40000000).sum()), "%.3fs" % (time.time() - t), flush=True)
for W in (240, 960, 1920):
    cam, cfg = api.make_config(sc, W, max_depth=3, mc_samples=4, pixel_samples=1, seed=1)
    t = time.time()
    ids, rst = p.primary_hits(cam, cfg)
    print("primary", W, "hits", int((ids != 0x40000000).sum()), "%.3fs" % (time.time() - t), flush=True)
    for depth in (1, 2, 3):
        cam, cfg = api.make_config(sc, W, max_depth=depth, mc_samples=4, pixel_samples=1, seed=1)
        t = time.time()
        img, st = p.render(cam, cfg)
        print("render", W, "depth", depth, "rays", st.rays, "shadow", st.shadow_rays, "ms %.2f" % st.ms_render, "%.3fs" % (time.time() - t), flush=True)
print("REPRO_DONE", flush=True)
