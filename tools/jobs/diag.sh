mkdir -p gpurun_out
python tools/diag_refview.py 2>&1 | tee gpurun_out/diag.log
