#!/usr/bin/env bash
cd "$(dirname "$0")/.."
out=gpurun_out/r02_3
mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > "$out/tests.log"
tail -4 "$out/tests.log"
ROUNDS=2 bash tools/ab_bench.sh > "$out/ab.log" 2>&1; grep -E "===|C2_bf16|S1024|D64|C3_varlen|_full" "$out/ab.log" | cut -c1-110
timeout 600 python bench.py > "$out/bench_c2.json" 2> "$out/bench.err"; cat "$out/bench_c2.json"; tail -3 "$out/bench.err"
