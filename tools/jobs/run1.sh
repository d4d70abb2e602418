mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cornell or known_answers or closest_hit_bit_exact or edge_case or raycaster or scheduling_modes or occluded or golden or ragged or config1" 2>&1 | tail -15 > gpurun_out/t_flat.log
for f in 1 0; do
  TRN_FLAT_PAIRS=$f python bench.py --workload cornell --mode pass --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/b_flat$f.json 2> gpurun_out/b_flat$f.err
  python tools/summarize_bench.py < gpurun_out/b_flat$f.json
done
