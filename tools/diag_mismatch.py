import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import bindings as ob
from turner_b200 import api, scenes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_parity as T
for sc in T.all_scenes(scenes):
    o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
    p = api.Scene.from_dict(sc)
    for inside in (False, True):
        ro, rd = scenes.random_rays(sc, 100000, seed=21, inside=inside)
        rd[::7, 0] = 0; rd[::11, 1] = 0; rd[::13, 2] = 0
        i_o, r_o = o.intersect(ro, rd, 0)
        i_g, r_g = p.intersect(ro, rd)
        bad = np.nonzero(i_g != i_o)[0]
        print(sc["name"], inside, "mismatches", bad.size, "height", p.height)
        for k in bad[:4]:
            print("   ray", k, "o", ro[k], "d", rd[k], "oracle", i_o[k], r_o[k], "gpu", i_g[k], r_g[k])
