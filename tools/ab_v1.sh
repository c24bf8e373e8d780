#!/usr/bin/env bash
# interleaved A/B of library variants on the round-1 forward kernel path (FA_B200_FWD_KERNEL=1)
cd "$(dirname "$0")/.."
for round in $(seq 1 ${ROUNDS:-2}); do
for lib in flash-attention-v100_b200/lib/libfa_b200*.so; do
  tag=$(basename "$lib" .so); tag=${tag#libfa_b200}; tag=${tag#_}; tag=${tag:-default}
  echo "=== $tag (round $round)"
  FA_B200_FWD_KERNEL=${FWD_KERNEL:-1} QUICK_BENCH_ONLY=1 FA_B200_LIB="$PWD/$lib" timeout -s KILL ${AB_TIMEOUT:-90} python tests/gpu_quick.py "$tag" 2>&1 | grep -E '"ms"|rror|Traceback' | grep -E "${AB_FILTER:-.}" | cut -c1-120
done
done
