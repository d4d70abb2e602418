import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import bindings as ob
from turner_b200 import api, scenes
for n in (4, 16):
    sc = scenes.tiled_box(n)
    o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
    p = api.Scene.from_dict(sc)
    rng = np.random.RandomState(n)
    m = 200000
    org = rng.uniform(0.05, 0.95, (m, 3)).astype(np.float32)
    d = rng.normal(size=(m, 3)).astype(np.float32)
    k = m // 4
    corner = rng.randint(0, n + 1, (k, 3)).astype(np.float32) / n
    face = rng.randint(0, 3, k)
    corner[np.arange(k), face] = rng.randint(0, 2, k)
    d[:k] = corner - org[:k]
    d[k:2 * k] = np.round(d[k:2 * k] * 4) / 8
    org[k:2 * k] = np.round(org[k:2 * k] * 8) / 8
    i_o, r_o = o.intersect(org, d, 0)
    i_e, r_e = o.intersect(org, d, 1)
    i_b, r_b = o.intersect(org, d, 2)
    i_g, r_g = p.intersect(org, d)
    bad = np.nonzero(i_g != i_o)[0]
    print("n", n, "mismatch gpu/ref", bad.size, "by group", [(int(((bad >= a) & (bad < b)).sum())) for a, b in ((0, k), (k, 2 * k), (2 * k, m))],
          "oracle early-exit vs ref", int((i_e != i_o).sum()), "brute vs ref", int((i_b != i_o).sum()), "gpu vs early", int((i_g != i_e).sum()))
    same_r = (r_g[bad, 0] == r_o[bad, 0]).sum()
    miss_g = (i_g[bad] == api.MISS_ID).sum(); miss_o = (i_o[bad] == ob.MISS).sum()
    print("   of the mismatches: equal r (ties)", int(same_r), "gpu miss", int(miss_g), "ref miss", int(miss_o))
    for j in [x for x in bad if i_g[x] == api.MISS_ID][:3] + list(bad[:3]):
        print("   ray", j, org[j].tolist(), d[j].tolist(), "ref", i_o[j], r_o[j], "gpu", i_g[j], r_g[j], "early", i_e[j], "brute", i_b[j])
