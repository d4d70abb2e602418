mkdir -p gpurun_out
tools/ab_bench.sh "TRN_KD_LAMBDA=0.8" "TRN_KD_LAMBDA=0.85" "TRN_KD_LAMBDA=0.9" "TRN_KD_LAMBDA=0.95" "TRN_KD_LAMBDA=1.0" "TRN_KD_LAMBDA=0.85 TRN_KD_KT=40" "TRN_KD_LAMBDA=0.9 TRN_KD_KT=40" "TRN_KD_LAMBDA=0.85 TRN_KD_KT=22" 2>&1 | tee gpurun_out/ab_mesh.log
