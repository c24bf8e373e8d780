#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02_26
timeout 120 tools/ubench/ubench_softmax_variants 2>&1 | tee gpurun_out/r02_26/ubench_softmax_variants.txt
