#!/usr/bin/env bash
# Round-2 GPU call 1: full GPU suite on the CLC-scheduler build, quick bench, yardstick baseline, sanitizers.
cd "$(dirname "$0")/.."
out=gpurun_out/r02_1
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > "$out/gpu.txt" 2>&1
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > "$out/tests.log"
tail -5 "$out/tests.log"
timeout 200 python bench.py > "$out/bench_c2.json" 2> "$out/bench.err"; cut -c1-400 "$out/bench_c2.json"
timeout 400 python tools/yardstick.py --shapes c2,full,s1k --iters 10 --only ours,cudnn --out "$out/yardstick.json" > "$out/yardstick.log" 2> "$out/yardstick.err"
tail -8 "$out/yardstick.log"
timeout 100 python tools/varlen_bench.py > "$out/varlen_c3.log" 2>&1; tail -3 "$out/varlen_c3.log"
for tool in synccheck racecheck; do
  timeout 500 compute-sanitizer --tool "$tool" --print-limit 20 python tools/sanitize_target.py > "$out/sanitizer_${tool}.txt" 2>&1
  echo "== $tool: exit $?"; tail -n 4 "$out/sanitizer_${tool}.txt"
done
