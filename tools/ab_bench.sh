#!/bin/bash
# development aid: A/B runs of bench.py under different kernel settings, one summary line each
# usage: tools/ab_bench.sh "ENV1=a ENV2=b" "ENV1=c" ...
for cfg in "$@"; do
  out=gpurun_out/ab_$(echo "$cfg" | tr ' =/' '___').json
  env $cfg timeout -s KILL 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > "$out" 2> "$out.err"
  echo "[$cfg] rc=$? $(python tools/summarize_bench.py < "$out" | head -1)"
done
