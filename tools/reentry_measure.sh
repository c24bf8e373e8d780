#!/usr/bin/env bash
# Short confirmation pass (one gpurun call): e2e pipeline probe, the bench line, then the GPU test suite.
cd "$(dirname "$0")/.."
out=gpurun_out/reentry
mkdir -p "$out"
timeout 150 python tools/e2e_probe.py > "$out/e2e_probe.log" 2>&1; cp gpurun_out/e2e_probe.json "$out/" 2>/dev/null
timeout 200 python bench.py > "$out/bench_c2.json" 2> "$out/bench.err"
timeout 120 env FA_E2E_PIPE=2s python bench.py --steps 10 > "$out/bench_c2_pipe2s.json" 2>> "$out/bench.err"
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -5 > "$out/tests.log"
cat "$out/e2e_probe.json" | tr -d '\n' | cut -c1-1500; echo
cut -c1-200 "$out/bench_c2.json"; grep -o '"e2e": {[^}]*}' "$out/bench_c2.json" "$out/bench_c2_pipe2s.json"
cat "$out/tests.log"
