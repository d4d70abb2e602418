#!/bin/bash
cd "$GRAFT_REPO_ROOT"
tools/ab_bench.sh "X=base" "TRN_L2_PERSIST=1 TRN_KD_DEBUG=1"
TRN_AB_LIB=turner_b200/libturner_b200_ide.so tools/ab_bench.sh "X=ide"
tools/ab_bench.sh --job "X=base" "TRN_L2_PERSIST=1"
grep "\[l2\]" gpurun_out/ab_TRN_L2_PERSIST_1_TRN_KD_DEBUG_1.json.err | head -2
