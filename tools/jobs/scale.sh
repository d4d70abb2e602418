N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-roofline > gpurun_out/r2b_job_n$N.json 2> gpurun_out/s_m$N.err
python tools/summarize_bench.py < gpurun_out/r2b_job_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload cornell --steps 2 --warmup 1 --no-roofline > gpurun_out/r2b_job_cornell_n$N.json 2> gpurun_out/s_c$N.err
python tools/summarize_bench.py < gpurun_out/r2b_job_cornell_n$N.json
