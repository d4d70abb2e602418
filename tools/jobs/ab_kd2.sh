#!/bin/bash
# per-reference plane records are in: full GPU suite, then the device builder's constants once more on the cheaper TEST phase
cd "$GRAFT_REPO_ROOT"
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
tools/ab_bench.sh "X=base" "TRN_KD_KT=30" "TRN_KD_KT=50" "TRN_KD_KT=60" "TRN_KD_KI=15" "TRN_KD_KI=25" "TRN_KD_LEAF=4" "TRN_KD_LEAF=6" "TRN_KD_LAMBDA=0.85" "TRN_KD_LAMBDA=0.95"
