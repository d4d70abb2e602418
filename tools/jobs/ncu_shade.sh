mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:shade_bounce -s 2 -c 2 -f -o gpurun_out/r2_shade python tools/profile_step.py cornell 1 > gpurun_out/ncu_shade.log 2>&1
tail -3 gpurun_out/ncu_shade.log
