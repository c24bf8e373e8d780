#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02_17
for r in 255 192 168; do echo "== maxrregcount $r"; timeout 120 tools/ubench/ubench_softmax_tile_r$r; done 2>&1 | tee gpurun_out/r02_17/ubench_softmax_tile.txt
