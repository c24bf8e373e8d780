#!/usr/bin/env bash
# One pass over everything profiles/ is built from (run on the GPU box through gpurun). Output: gpurun_out/final/.
cd "$(dirname "$0")/.."
out=gpurun_out/final
mkdir -p "$out"
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3 > "$out/tests.log"
timeout 200 python bench.py > "$out/bench_c2.json" 2> "$out/bench.err"
timeout 200 python bench.py --impl reference > "$out/bench_reference_arm.json" 2>> "$out/bench.err"
timeout 200 python bench.py --workload c2gqa > "$out/bench_c2gqa.json" 2>> "$out/bench.err"
timeout 200 python bench.py --workload c5 --steps 10 > "$out/bench_c5_1gpu.json" 2>> "$out/bench.err"
timeout 100 python tools/varlen_bench.py > "$out/varlen_c3.log" 2>&1; cp gpurun_out/varlen_bench.json "$out/varlen_c3.json" 2>/dev/null
timeout 200 python tools/decode_bench.py > "$out/decode_c4.log" 2>&1; cp gpurun_out/decode_bench.json "$out/decode_c4.json" 2>/dev/null
timeout 100 python tools/bwd_quick.py > "$out/bwd_quick.log" 2>&1
timeout 400 python tools/yardstick.py --shapes c2,full,s1k,d64,d256 --iters 10 --out "$out/yardstick.json" > "$out/yardstick.log" 2> "$out/yardstick.err"
# profiler passes (never a bench number)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_c2.csv" python bench.py --steps 2 --warmup 1 > "$out/ncu_launch.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 3 -c 1 -o "$out/fwd_c2" -f python tools/profile_target.py c2 5 > "$out/ncu_fwd.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fa_bwd -s 3 -c 3 -o "$out/bwd_c2" -f python tools/profile_target.py c2 2 bwd > "$out/ncu_bwd.log" 2>&1
tail -2 "$out/tests.log"; cat "$out/bench_c2.json" | cut -c1-300
