#!/usr/bin/env bash
# forward v2 bring-up: quick parity cases first (v2 is the default for head_dim 128), then timings v1 / v2 / round-1 library
cd "$(dirname "$0")/.."
out=gpurun_out/r02_4
mkdir -p "$out"
timeout 300 python tests/gpu_quick.py v2 > "$out/quick_v2.log" 2>&1; echo "quick v2 exit $?"; grep -E '"ok"|error|rror' "$out/quick_v2.log" | cut -c1-200 | head -20
grep -E '"ms"' "$out/quick_v2.log" | cut -c1-120
echo "=== v1 (this build)"; FA_B200_FWD_KERNEL=1 QUICK_BENCH_ONLY=1 timeout 200 python tests/gpu_quick.py v1 2>&1 | grep -E '"ms"|rror' | cut -c1-120
echo "=== r01 library"; QUICK_BENCH_ONLY=1 FA_B200_LIB=$PWD/flash-attention-v100_b200/lib/libfa_b200_r01.so timeout 200 python tests/gpu_quick.py r01 2>&1 | grep -E '"ms"|rror' | cut -c1-120
echo "=== v2 again"; QUICK_BENCH_ONLY=1 timeout 200 python tests/gpu_quick.py v2b 2>&1 | grep -E '"ms"|rror' | cut -c1-120
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > "$out/tests.log"; tail -5 "$out/tests.log"
