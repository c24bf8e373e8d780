#!/usr/bin/env bash
# Runs tests/gpu_quick.py (timings only) for every library variant in flash-attention-v100_b200/lib,
# interleaved over ROUNDS rounds on the same box, so run-to-run clock differences cancel out (A/B tuning).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for round in $(seq 1 ${ROUNDS:-2}); do
for lib in flash-attention-v100_b200/lib/libfa_b200*.so; do
  tag=$(basename "$lib" .so); tag=${tag#libfa_b200}; tag=${tag#_}; tag=${tag:-default}
  echo "=== $tag (round $round)"
  QUICK_BENCH_ONLY=1 FA_B200_LIB="$PWD/$lib" timeout -s KILL ${AB_TIMEOUT:-90} python tests/gpu_quick.py "$tag" 2>&1 | grep -E '"ms"|rror|Traceback' | cut -c1-120
done
done
