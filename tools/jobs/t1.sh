mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "one_leaf" 2>&1 | tail -40 > gpurun_out/t1.log
grep -n "Error\|assert" gpurun_out/t1.log | head
