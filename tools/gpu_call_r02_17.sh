#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02_17
timeout 120 tools/ubench/ubench_softmax_tile_r192 2>&1 | tee gpurun_out/r02_17/ubench_softmax_tile_ovh.txt
