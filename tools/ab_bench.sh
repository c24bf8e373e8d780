#!/usr/bin/env bash
# Runs tools/gpu_quick.py once per library variant found in flash-attention-v100_b200/lib (A/B tuning).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in flash-attention-v100_b200/lib/libfa_b200*.so; do
  tag=$(basename "$lib" .so); tag=${tag#libfa_b200}; tag=${tag#_}; tag=${tag:-default}
  echo "=== $tag"
  FA_B200_LIB="$PWD/$lib" timeout -s KILL 300 python tools/gpu_quick.py "$tag" 2>&1 | grep -E '"ms"|error|"ok": false|Traceback' | cut -c1-200
done
