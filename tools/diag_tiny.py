"""diagnostic: the 'tiny scale' case of test_device_builder_degenerate_and_awkward_inputs on both builders (run on a GPU box)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bindings as ob
from turner_b200 import api, scenes

ob.build()
for scale in (float(x) for x in (sys.argv[1:] or ["1e-3", "1e-2", "1.0"])):
    rng = np.random.RandomState(9)
    v = ((rng.uniform(-1, 1, (2000, 1, 3)) + 0.05 * rng.uniform(-1, 1, (2000, 3, 3))).astype(np.float32) * np.float32(scale)).reshape(-1, 3, 3)
    nrm = np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True) + 1e-30
    base = scenes.four_triangles()
    sc = {**base, "name": "custom", "vertices": v.reshape(-1, 9), "normals": np.repeat(nrm, 3, axis=0).reshape(-1, 9).astype(np.float32),
          "diffuse": np.tile(np.array([[0.5, 0.5, 0.5, 1]], np.float32), (v.shape[0], 1))}
    o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
    lo, hi = v.reshape(-1, 3).min(0), v.reshape(-1, 3).max(0)
    ctr, ext = 0.5 * (lo + hi), hi - lo
    n = 100000
    org = (ctr + ext * rng.uniform(-1.5, 1.5, (n, 3))).astype(np.float32)
    pick = rng.randint(0, v.shape[0], n)
    bary = rng.dirichlet([1, 1, 1], n)
    tgt = (v[pick] * bary[:, :, None]).sum(1).astype(np.float32)
    d = (tgt - org).astype(np.float32)
    i_o, r_o = o.intersect(org, d, 0)
    for builder in ("host", "gpu"):
        for mode in ("",):
            if mode:
                os.environ["TRN_PERSISTENT"] = mode
            else:
                os.environ.pop("TRN_PERSISTENT", None)
            p = api.Scene.from_dict(sc, builder=builder)
            i_g, r_g = p.intersect(org, d)
            bad = (i_g != i_o)
            print("scale %g builder %s mode '%s': id diff %d, r-bit diff %d, gpu-miss/oracle-hit %d, height %d" % (
                scale, builder, mode, bad.sum(), (r_g[:, 0].view(np.uint32) != r_o[:, 0].view(np.uint32)).sum(),
                ((i_g == api.MISS_ID) & (i_o != ob.MISS)).sum(), p.height), flush=True)
            if bad.any() and mode == "":
                for j in np.nonzero(bad)[0][:3]:
                    print("    ray", j, org[j], d[j], "oracle", i_o[j], r_o[j], "gpu", i_g[j], r_g[j])
