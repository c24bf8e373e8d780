#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02_27
timeout 120 tools/ubench/ubench_softmax_variants 2>&1 | tee gpurun_out/r02_27/ubench_softmax_variants.txt
L=$PWD/flash-attention-v100_b200/lib
FA_B200_LIB=$L/libfa_b200_trace.so timeout 200 python tools/trace_timeline.py 1 1024 > gpurun_out/r02_27/timeline_summary_causal_1024.txt 2>&1
head -4 gpurun_out/r02_27/timeline_summary_causal_1024.txt
cp gpurun_out/timeline_causal_1024.txt gpurun_out/r02_27/
