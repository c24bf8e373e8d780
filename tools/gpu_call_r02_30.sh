#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02_30
AB_FILTER='C2_bf16|S1024|_full|C3_|S16384|C2gqa' ROUNDS=4 bash tools/gpu_ab.sh 2>&1 | grep -v "^=== parity" | tee gpurun_out/r02_30/ab.log
