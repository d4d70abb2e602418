"""diagnostic: where trn_occluded differs from the reference predicate on the tiled box (run on a GPU box)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import bindings as ob
from turner_b200 import api, scenes
from test_gpu_parity import _secondary_shaped_rays

ob.build()
for sc in (scenes.tiled_box(8), scenes.tiled_box(16)):
    p = api.Scene.from_dict(sc)
    o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
    n = 100000
    rng = np.random.RandomState(5)
    sets = [("secondary", _secondary_shaped_rays(sc, n, 11)), ("inside", scenes.random_rays(sc, n, seed=9, inside=True)),
            ("outside", scenes.random_rays(sc, n // 2, seed=10, inside=False))]
    for name, (ro, rd) in sets:
        m = ro.shape[0]
        i_o, r_o = o.intersect(ro, rd, 0)
        i_g, r_g = p.intersect(ro, rd)
        hit = i_o != ob.MISS
        r = r_o[:, 0]
        tmax = rng.uniform(0.0, 3.0, m).astype(np.float32) * np.float32(np.abs(sc["vertices"]).max())
        sel = hit & (rng.rand(m) < 0.5)
        kind = rng.randint(0, 3, m)
        exact = np.where(kind == 0, r, np.where(kind == 1, np.nextafter(r, np.float32(-1)), np.nextafter(r, np.float32(np.inf))))
        tmax = np.where(sel, exact, tmax).astype(np.float32)
        want = hit & (r <= tmax)
        got = p.occluded(ro, rd, tmax)
        bad = got != want
        print(sc["name"], name, "closest id diff", int((i_g != i_o).sum()), "occl diff", int(bad.sum()), "of", m,
              "| got&!want", int((got & ~want).sum()), "want&!got", int((want & ~got).sum()),
              "| by kind (sel)", [int((bad & sel & (kind == k)).sum()) for k in range(3)], "random tmax", int((bad & ~sel).sum()),
              "| zero-dir comps among bad", int((rd[bad] == 0).any(1).sum()))
        for j in np.nonzero(bad)[0][:4]:
            print("   ray", j, "o", ro[j], "d", rd[j], "r", r[j], "tmax", tmax[j], "id", i_o[j], "got", got[j], "want", want[j])
