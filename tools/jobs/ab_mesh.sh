mkdir -p gpurun_out
tools/ab_bench.sh "TRN_X=1" "TRN_AB_LIB=turner_b200/libturner_b200_nofold.so" 2>&1 | tee gpurun_out/ab_mesh.log
