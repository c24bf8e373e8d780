#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02_32
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r02_32/gpu_tests.txt
bash tools/gpu_bench.sh 2>&1 | tee gpurun_out/r02_32/bench_summary.txt
cp gpurun_out/bench/bench_c2.json gpurun_out/r02_32/
