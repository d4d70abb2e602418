#!/bin/bash
# round 2, third session: ncu capture of the pooled kernel on the per-reference plane layout + launch list of the job bench
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:trace_pooled -c 8 -f -o gpurun_out/r2c_prof_pooled_mesh1m python tools/profile_step.py mesh1m 1 > gpurun_out/ncu_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2c_launches_bench_mesh1m.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
tail -2 gpurun_out/ncu_a.log; tail -c 400 gpurun_out/ncu_c.log
