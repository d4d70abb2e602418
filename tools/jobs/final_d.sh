#!/bin/bash
# round 2, third session: final N=1 bench lines (per-reference plane layout) + smoke
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python bench.py > gpurun_out/r2c_bench_mesh1m_n1.json 2> gpurun_out/bm.err; python tools/summarize_bench.py < gpurun_out/r2c_bench_mesh1m_n1.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_driver_like_n1.json 2> gpurun_out/bd.err; python tools/summarize_bench.py < gpurun_out/r2c_driver_like_n1.json
python bench.py --workload cornell --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2c_bench_cornell_n1.json 2> gpurun_out/bc.err; python tools/summarize_bench.py < gpurun_out/r2c_bench_cornell_n1.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
