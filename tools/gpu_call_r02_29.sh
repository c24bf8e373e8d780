#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02_29
L=$PWD/flash-attention-v100_b200/lib
echo "=== parity (need lib)"; FA_B200_LIB=$L/libfa_b200_need.so timeout 300 python tests/gpu_quick.py parity 2>&1 | grep -E '"ok": false|rror' | cut -c1-200
AB_FILTER='C2_bf16|S1024|_full|C3_|D64|D256|S16384' ROUNDS=2 bash tools/gpu_ab.sh 2>&1 | grep -v "^=== parity" | tee gpurun_out/r02_29/ab.log
FA_B200_LIB=$L/libfa_b200_trace.so timeout 200 python tools/trace_timeline.py 1 1024 > gpurun_out/r02_29/timeline_summary_causal_1024.txt 2>&1
cp gpurun_out/timeline_causal_1024.txt gpurun_out/r02_29/
