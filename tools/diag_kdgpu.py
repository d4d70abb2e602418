"""diagnostic: device kd builder on small scenes with per-level statistics (run on a GPU box)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TRN_KD_DEBUG"] = "1"
from turner_b200 import api, scenes

names = sys.argv[1:] or ["four", "cornell", "cs48"]
for nm in names:
    sc = {"four": scenes.four_triangles, "cube": scenes.unit_cube, "cornell": lambda: scenes.fixture("cornell_box"),
          "cs48": lambda: scenes.cubesphere(48), "cs288": lambda: scenes.cubesphere(288), "soup": lambda: scenes.random_soup(5000, 2)}[nm]()
    print("==", sc["name"], len(sc["vertices"]), flush=True)
    try:
        p = api.Scene.from_dict(sc, builder="gpu")
        print("   ok: build %.1f ms height %d refs %d cuts %d nodes %d" % (p.info.build_ms, p.height, p.info.num_leaf_refs, p.info.num_cut_nodes, p.num_nodes), flush=True)
    except Exception as e:
        print("   FAILED", e, flush=True)
