mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:trace_flat -c 4 -f -o gpurun_out/r2_flat python tools/profile_step.py cornell 1 > gpurun_out/ncu_flat.log 2>&1
tail -3 gpurun_out/ncu_flat.log
