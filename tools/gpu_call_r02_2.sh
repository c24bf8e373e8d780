#!/usr/bin/env bash
# Round-2 GPU call 2: rest of the GPU suite, A/B of the CLC scheduler against the round-1 library, bench with configs.
cd "$(dirname "$0")/.."
out=gpurun_out/r02_2
mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > "$out/tests.log"
tail -8 "$out/tests.log"
ROUNDS=3 bash tools/ab_bench.sh > "$out/ab.log" 2>&1; cat "$out/ab.log"
timeout 400 python bench.py > "$out/bench_c2.json" 2> "$out/bench.err"; cut -c1-300 "$out/bench_c2.json"; tail -3 "$out/bench.err"
