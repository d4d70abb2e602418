#!/usr/bin/env python3
"""development aid: where the oracle's traversal of a device-built tree's reference view differs from its own tree"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turner_b200 import api, scenes
from oracle import bindings as ob

for sc in [scenes.fixture("cornell_box"), scenes.cubesphere(32), scenes.random_soup(3000, 4)]:
    p = api.Scene.from_dict(sc, builder="gpu")
    o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
    nodes = p.nodes()
    o2 = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"], nodes=nodes, box=np.array(p.info.box, np.float32))
    ro, rd = scenes.random_rays(sc, 40000, seed=12, inside=True)
    i1, r1 = o.intersect(ro, rd, 0)
    i2, r2 = o2.intersect(ro, rd, 0)
    ig, rg = p.intersect(ro, rd)
    bad = np.nonzero(i1 != i2)[0]
    print(sc["name"], "height", p.height, "nodes", len(nodes), "diff oracle-on-view vs oracle:", len(bad), " gpu vs oracle:", int((ig != i1).sum()))
    for k in bad[:8]:
        print("  ray", k, "own tree id", i1[k], "r", r1[k], " view id", i2[k], "r", r2[k], " gpu id", ig[k], "o", ro[k], "d", rd[k])
