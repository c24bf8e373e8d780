#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02_38
AB_FILTER='C2_bf16|S1024|_full|D64|C3_' ROUNDS=3 bash tools/gpu_ab.sh 2>&1 | grep -v "^=== parity" | tee gpurun_out/r02_38/ab.log
