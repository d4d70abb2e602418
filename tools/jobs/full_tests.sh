mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --workload cornell --mode pass --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/b_c.json 2> gpurun_out/b_c.err; python tools/summarize_bench.py < gpurun_out/b_c.json
python bench.py --workload mesh1m --mode pass --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/b_m.json 2> gpurun_out/b_m.err; python tools/summarize_bench.py < gpurun_out/b_m.json
