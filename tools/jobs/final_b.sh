mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "one_leaf" 2>&1 | tail -3
python bench.py > gpurun_out/r2b_bench_mesh1m_n1.json 2> gpurun_out/bm.err; python tools/summarize_bench.py < gpurun_out/r2b_bench_mesh1m_n1.json
python bench.py --workload cornell --steps 2 --warmup 1 > gpurun_out/r2b_bench_cornell_n1.json 2> gpurun_out/bc.err; python tools/summarize_bench.py < gpurun_out/r2b_bench_cornell_n1.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_driver_like_n1.json 2> gpurun_out/bd.err; python tools/summarize_bench.py < gpurun_out/r2b_driver_like_n1.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
