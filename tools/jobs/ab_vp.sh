#!/bin/bash
# A/B: inline pop at cut-off voids (TRN_PQ_VOIDPOP) on the 1M mesh, pass mode; dump of the device-built tree for the CPU emulator
cd "$GRAFT_REPO_ROOT"
tools/ab_bench.sh "X=base"
TRN_AB_LIB=turner_b200/libturner_b200_vp.so tools/ab_bench.sh "X=vp" "X=vp TRN_PQ_GATE=14" "X=vp TRN_PQ_GATE=6" "X=vp TRN_PQ_WALK=16"
TRN_AB_LIB=turner_b200/libturner_b200_vp3.so tools/ab_bench.sh "X=vp3"
TRN_KD_DUMP=/tmp/kd_mesh1m.bin python - <<'P'
import sys
sys.path.insert(0, '.')
from turner_b200 import api, scenes
sc = scenes.cubesphere(288)
s = api.Scene.from_dict(sc, builder="gpu", device=0)
print("dumped", s.height)
P
gzip -1 -c /tmp/kd_mesh1m.bin > gpurun_out/kd_mesh1m.bin.gz; ls -la gpurun_out/kd_mesh1m.bin.gz
