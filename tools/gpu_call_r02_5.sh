#!/usr/bin/env bash
cd "$(dirname "$0")/.."
out=gpurun_out/r02_5
mkdir -p "$out"
timeout 300 python tests/gpu_quick.py v2p > "$out/quick_v2p.log" 2>&1; echo "quick v2 pair exit $?"; grep -E '"ok"|error|rror' "$out/quick_v2p.log" | cut -c1-160 | head -20
grep -E '"ms"' "$out/quick_v2p.log" | cut -c1-120
echo "=== v1 (this build)"; FA_B200_FWD_KERNEL=1 QUICK_BENCH_ONLY=1 timeout 200 python tests/gpu_quick.py v1 2>&1 | grep -E '"ms"|rror' | cut -c1-120
echo "=== v2 single"; FA_B200_FWD_KERNEL=2s QUICK_BENCH_ONLY=1 timeout 200 python tests/gpu_quick.py v2s 2>&1 | grep -E '"ms"|rror' | cut -c1-120
echo "=== v2 pair again"; QUICK_BENCH_ONLY=1 timeout 200 python tests/gpu_quick.py v2pb 2>&1 | grep -E '"ms"|rror' | cut -c1-120
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > "$out/tests.log"; tail -5 "$out/tests.log"
