#!/bin/bash
# N-GPU check on the per-reference plane layout: the multi-GPU tests (skipped on 1-GPU boxes) and the mesh job line
N=$1
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "multi_gpu or process_per_gpu" > gpurun_out/r2c_pytest_${N}gpu.log 2>&1; tail -2 gpurun_out/r2c_pytest_${N}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-roofline --no-cpu-baseline > gpurun_out/r2c_job_n$N.json 2> gpurun_out/s_m$N.err
python tools/summarize_bench.py < gpurun_out/r2c_job_n$N.json
