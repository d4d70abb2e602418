#!/usr/bin/env python3
"""Quick on-GPU sanity run (development aid): GPU path vs the CPU oracle on small inputs + rough timings."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import bindings as ob
from turner_b200 import api, scenes

print("devices", api.device_count(), flush=True)
ok = True
for sc in [scenes.four_triangles(), scenes.fixture("cornell_box"), scenes.fixture("furnace_test"),
           scenes.random_soup(2000, 2), scenes.cubesphere(48)]:
    o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
    p = api.Scene.from_dict(sc)
    for inside in (False, True):
        ro, rd = scenes.random_rays(sc, 200000, seed=7, inside=inside)
        i0, r0 = o.intersect(ro, rd, 0)
        t = time.time()
        i1, r1 = p.intersect(ro, rd)
        dt = time.time() - t
        same = np.array_equal(i0, i1) and np.array_equal(r0.view(np.uint32), r1.view(np.uint32))
        ok &= same
        print(sc["name"], "inside", inside, "hits", int((i0 != ob.MISS).sum()), "ids+rst bit-equal", same,
              "id mismatches", int((i0 != i1).sum()), "%.3fs" % dt, flush=True)

sc = scenes.fixture("cornell_box")
o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
p = api.Scene.from_dict(sc)
W = 128
cam, cfg = api.make_config(sc, W, max_depth=3, mc_samples=4, pixel_samples=4, seed=7)
ocfg = ob.make_cfg(sc, W, max_depth=3, mc_samples=4, pixel_samples=4, rng_mode=1, seed=7)
ids, rst = p.primary_hits(cam, cfg)
dirs = ob.primary_dirs(ocfg)
oi, orst = o.intersect(np.tile(np.array(list(ocfg.cam_pos), np.float32), (dirs.size // 3, 1)), dirs.reshape(-1, 3), 0)
print("primary ids equal", np.array_equal(ids.reshape(-1), oi), "rst equal",
      np.array_equal(rst.reshape(-1, 3).view(np.uint32), orst.view(np.uint32)), flush=True)
img, st = p.render(cam, cfg)
ref, _, ost = o.render(ocfg)
d = np.abs(img - ref)
print("render rays gpu/oracle", st.rays, ost.num_rays, "prim", st.prim_rays, ost.num_prim_rays, "shadow", st.shadow_rays,
      ost.num_shadow_rays, "max abs diff", d.max(), "mean abs", d.mean(), "mean img", ref.mean(),
      "pixels off by >1e-3:", int((d.max(-1) > 1e-3).sum()), "of", W * W, "ms %.2f" % st.ms_render, flush=True)

# throughput feel: cornell 1024^2, m4, pps 4
api.set_profiling(True)
for (name, scn, W, pps) in [("cornell", sc, 1024, 4), ("cubesphere96", scenes.cubesphere(96), 1024, 4)]:
    ps = api.Scene.from_dict(scn)
    cam, cfg = api.make_config(scn, W, max_depth=3, mc_samples=4, pixel_samples=pps, seed=1)
    for it in range(2):
        img, st = ps.render(cam, cfg)
    print(name, "W", W, "pps", pps, "rays", st.rays, "shadow", st.shadow_rays, "ms %.1f" % st.ms_render,
          "Mrays/s %.1f" % (st.rays / st.ms_render / 1e3), "trace %.1f shadow %.1f shade %.1f other %.1f launches %d" %
          (st.ms_trace, st.ms_shadow, st.ms_shade, st.ms_other, st.launches), flush=True)
print("ALL_OK" if ok else "MISMATCH")
