#!/usr/bin/env bash
# Short confirmation pass (one gpurun call): bench line, reference arm, config 3, ncu launch list of the bench.
cd "$(dirname "$0")/.."
out=gpurun_out/reentry2
mkdir -p "$out"
timeout 200 python bench.py --impl reference > "$out/bench_reference_arm.json" 2> "$out/bench.err"
timeout 200 python bench.py > "$out/bench_c2.json" 2>> "$out/bench.err"
timeout 100 python tools/varlen_bench.py > "$out/varlen_c3.log" 2>&1; cp gpurun_out/varlen_bench.json "$out/varlen_c3.json" 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_c2.csv" python bench.py --steps 2 --warmup 1 > "$out/ncu_launch.log" 2>&1
cut -c1-160 "$out/bench_c2.json"; grep -o '"e2e": {[^}]*}' "$out/bench_c2.json"; cut -c1-200 "$out/bench_reference_arm.json"
grep -c fa_fwd "$out/launches_c2.csv"; tail -2 "$out/varlen_c3.log" | cut -c1-300
