#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02_43
L=$PWD/flash-attention-v100_b200/lib
for round in 1 2; do
for tag in default narrow; do
  lib=$L/libfa_b200.so; [ $tag = narrow ] && lib=$L/libfa_b200_narrow.so
  echo "=== $tag (round $round)"
  [ $round = 2 ] && export QUICK_BENCH_ONLY=1
  FA_B200_LIB=$lib timeout -s KILL 300 python tests/gpu_quick_d256.py $tag 2>&1 | grep -E '"name"|rror|Traceback' | cut -c1-160
done
done 2>&1 | tee gpurun_out/r02_43/d256_ab.log
