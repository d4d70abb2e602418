import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TRN_KD_DEBUG"] = "1"
import numpy as np, torch
import bench
from turner_b200 import api
sc = bench.load_scene("mesh1m")
scene = api.Scene.from_dict(sc, builder="gpu")
w = bench.WORKLOADS["mesh1m"]
cam, cfg = api.make_config(sc, w["width"], max_depth=3, mc_samples=4, pixel_samples=128, seed=1)
cfg.sample_begin, cfg.sample_stride = 0, 128
acc = torch.zeros(cfg.height, cfg.width, 4, device="cuda")
api.set_counting(True)
st = scene.render_device(cam, cfg, acc.data_ptr(), 0, device=0)
api.set_counting(False)
tp = list(st.trace_pooled); sp = list(st.shadow_pooled)
print("closest: queries", st.trace_queries, "steps/q %.2f cut steps/q %.2f leaves/q %.2f" % (tp[0]/st.trace_queries, tp[8]/st.trace_queries, tp[7]/st.trace_queries))
print("raw", tp, sp)
