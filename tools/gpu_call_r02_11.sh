#!/usr/bin/env bash
cd "$(dirname "$0")/.."
out=gpurun_out/r02_11
mkdir -p "$out"
AB_FILTER="C2_bf16|S1024|_full|D64|C3_|C1_" ROUNDS=2 FWD_KERNEL=1 bash tools/ab_v1.sh 2>&1 | tee "$out/ab_default_vs_r01.log" | cut -c1-110
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > "$out/tests.log"; tail -12 "$out/tests.log"
