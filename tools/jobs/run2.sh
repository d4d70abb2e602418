mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cornell or one_leaf or scheduling_modes or occluded or golden or config1 or raycaster or radiance" 2>&1 | tail -25 > gpurun_out/t_flat.log
tail -3 gpurun_out/t_flat.log
python bench.py --workload cornell --mode pass --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/b_flat1.json 2> gpurun_out/b_flat1.err
python tools/summarize_bench.py < gpurun_out/b_flat1.json
