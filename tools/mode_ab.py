#!/usr/bin/env python3
"""development aid: render time of the traversal schedules (TRN_PERSISTENT unset / 0 / 2 / 3) on trees of different leaf sizes"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from turner_b200 import api, scenes

api.set_profiling(True)
cases = [("furnace", scenes.fixture("furnace_test"), 1024, 8, 8, (1, 1, 1, 1)), ("cs48", scenes.cubesphere(48), 1024, 3, 4, (0, 0, 0, 1)),
         ("cs96", scenes.cubesphere(96), 1024, 3, 4, (0, 0, 0, 1)), ("soup5000", scenes.random_soup(5000, 2), 1024, 3, 4, (0, 0, 0, 1)),
         ("colored_cube", scenes.fixture("colored_cube"), 1024, 3, 4, (0, 0, 0, 1))]
for name, sc, W, D, m, bg in cases:
    p = api.Scene.from_dict(sc)
    cam, cfg = api.make_config(sc, W, max_depth=D, mc_samples=m, pixel_samples=2, seed=1, bg=bg)
    line = [name]
    for mode in (None, "0", "2", "3"):
        if mode is None:
            os.environ.pop("TRN_PERSISTENT", None)
        else:
            os.environ["TRN_PERSISTENT"] = mode
        for it in range(2):
            img, st = p.render(cam, cfg)
        line.append("%s: %.2f ms (trace %.2f shadow %.2f)" % (mode or "default", st.ms_render, st.ms_trace, st.ms_shadow))
    print("  ".join(line), "rays", st.rays, flush=True)
