#!/usr/bin/env bash
cd "$(dirname "$0")/.."
out=gpurun_out/r02_10
mkdir -p "$out"
AB_FILTER="C2_bf16|S1024|_full|D64|C3_|C1_" ROUNDS=2 FWD_KERNEL=1 bash tools/ab_v1.sh 2>&1 | tee "$out/ab_default_vs_r01.log" | cut -c1-110
for mode in 1 2s 2p; do
  FA_B200_FWD_KERNEL=$mode timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > "$out/tests_mode_$mode.log"; echo "mode $mode: $(tail -1 $out/tests_mode_$mode.log)"
done
