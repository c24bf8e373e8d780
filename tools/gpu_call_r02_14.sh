#!/usr/bin/env bash
# timeline of CTA 0 (non-causal S=4096) from the -DFA_TRACE build
cd "$(dirname "$0")/.."
out=gpurun_out/r02_14
mkdir -p "$out"
L=$PWD/flash-attention-v100_b200/lib
FA_B200_LIB=$L/libfa_b200_trace.so timeout 200 python tools/trace_timeline.py 0 4096 300 > "$out/timeline_trace.txt" 2>&1
head -3 "$out/timeline_trace.txt"
