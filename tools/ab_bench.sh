#!/bin/bash
# development aid: A/B runs of bench.py under different kernel settings, one summary line each
# usage: tools/ab_bench.sh [--job] "ENV1=a ENV2=b" "ENV1=c" ...      (TRN_AB_LIB=<path> selects another build of the library)
mode="--mode pass --steps 4 --warmup 3"
if [ "$1" = "--job" ]; then mode="--mode job --steps 1 --warmup 3 --no-roofline"; shift; fi
for cfg in "$@"; do
  out=gpurun_out/ab_$(echo "$cfg" | tr ' =/' '___').json
  env $cfg timeout -s KILL 300 python tools/ab_run.py $mode --no-cpu-baseline > "$out" 2> "$out.err"
  echo "[$cfg] rc=$? $(python tools/summarize_bench.py < "$out" | head -1)"
done
