#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02_15
timeout 120 tools/ubench/ubench_pipes 2>&1 | tee gpurun_out/r02_15/ubench_pipes.txt
