#!/bin/bash
# device kd build of the 1M mesh, three times in one process, phase timing (TRN_KD_DEBUG)
cd "$GRAFT_REPO_ROOT"
TRN_KD_DEBUG=1 python - 2>&1 <<'P' | grep -v "^\[kd-gpu\] level\|^   " | tail -40
import sys, time
sys.path.insert(0, '.')
from turner_b200 import api, scenes
sc = scenes.cubesphere(288)
for k in range(3):
    t = time.time()
    s = api.Scene.from_dict(sc, builder="gpu", device=0)
    print("build %d: wall %.1f ms, build_ms %.1f, height %d" % (k, 1e3 * (time.time() - t), s.info.build_ms, s.height), flush=True)
    del s
P
