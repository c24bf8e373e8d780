#!/usr/bin/env bash
# One gpurun call of the forward tuning loop: parity of the default library (tests/gpu_quick.py), interleaved A/B of
# every flash-attention-v100_b200/lib/libfa_b200*.so (except the trace build), the CTA-0 timeline of the -DFA_TRACE
# build if there is one, and optionally the GPU test suite (TESTS=1).  Output: gpurun_out/ab/.
cd "$(dirname "$0")/.."
out=gpurun_out/ab
mkdir -p "$out"
L=$PWD/flash-attention-v100_b200/lib
echo "=== parity (default lib)"; timeout 300 python tests/gpu_quick.py parity 2>&1 | grep -E '"ok": false|rror' | cut -c1-200
for round in $(seq 1 ${ROUNDS:-2}); do
for lib in $L/libfa_b200*.so; do
  case "$lib" in *trace*|*jitter*) continue;; esac
  tag=$(basename "$lib" .so); tag=${tag#libfa_b200}; tag=${tag#_}; tag=${tag:-default}
  echo "=== $tag (round $round)"
  QUICK_BENCH_ONLY=1 FA_B200_LIB="$lib" timeout -s KILL ${AB_TIMEOUT:-90} python tests/gpu_quick.py "$tag" 2>&1 | grep -E '"ms"|rror|Traceback' | grep -E "${AB_FILTER:-C2_bf16|S1024|_full|C2gqa|S16384|D64|C3_}" | sed -E 's/"ms": ([0-9.]{6})[0-9]*, "tflops": ([0-9.]{6})[0-9]*/\1 ms \2 TF/' | cut -c1-100
done
done 2>&1 | tee "$out/ab.log"
if [ -f $L/libfa_b200_trace.so ]; then
  FA_B200_LIB=$L/libfa_b200_trace.so timeout 200 python tools/trace_timeline.py ${TRACE_ARGS:-0 4096 300} > "$out/timeline.txt" 2>&1; head -3 "$out/timeline.txt"
fi
if [ -n "${TESTS:-}" ]; then timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee "$out/tests.log"; fi
