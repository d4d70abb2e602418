mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
ncu --set full --clock-control none --import-source on -k regex:trace_pooled -c 8 -f -o gpurun_out/r2b_prof_pooled_mesh1m python tools/profile_step.py mesh1m 1 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_flat -c 6 -f -o gpurun_out/r2b_prof_flat_cornell python tools/profile_step.py cornell 1 > gpurun_out/ncu_b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2b_launches_bench_mesh1m.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
tail -2 gpurun_out/ncu_a.log gpurun_out/ncu_b.log; tail -c 600 gpurun_out/ncu_c.log
