#!/bin/bash
# device-built tree of cubesphere(32) for the CPU suite's emulator test (tests/golden/kd_gpu_cubesphere32.bin.gz)
cd "$GRAFT_REPO_ROOT"
TRN_KD_DUMP=/tmp/kd_cs32.bin python - <<'P'
import sys
sys.path.insert(0, '.')
from turner_b200 import api, scenes
s = api.Scene.from_dict(scenes.cubesphere(32), builder="gpu", device=0)
print("height", s.height)
P
gzip -9 -c /tmp/kd_cs32.bin > gpurun_out/kd_gpu_cubesphere32.bin.gz; ls -la gpurun_out/kd_gpu_cubesphere32.bin.gz
