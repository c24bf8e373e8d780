#!/usr/bin/env bash
# Round-2 evidence pass (one gpurun call): GPU tests, both bench arms, every side bench, ncu launch list, ncu full
# captures of the config-2 and config-3 forwards and of the backward. Output: gpurun_out/r02_12/.
cd "$(dirname "$0")/.."
out=gpurun_out/r02_12
mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > "$out/tests.log"; tail -3 "$out/tests.log"
timeout 300 python bench.py --impl reference > "$out/bench_reference_arm.json" 2> "$out/bench.err"
timeout 400 python bench.py > "$out/bench_c2.json" 2>> "$out/bench.err"
cut -c1-400 "$out/bench_c2.json"
timeout 100 python tools/varlen_bench.py > "$out/varlen_c3.log" 2>&1; cp gpurun_out/varlen_bench.json "$out/varlen_c3.json" 2>/dev/null
timeout 200 python tools/decode_bench.py > "$out/decode_c4.log" 2>&1; cp gpurun_out/decode_bench.json "$out/decode_c4.json" 2>/dev/null
timeout 100 python tools/bwd_quick.py > "$out/bwd_quick.log" 2>&1
timeout 200 python tools/feature_bench.py > "$out/features.log" 2>&1; cp gpurun_out/feature_bench.json "$out/features.json" 2>/dev/null
timeout 100 python tools/host_overhead.py > "$out/host_overhead.log" 2>&1
timeout 400 python tools/yardstick.py --shapes c2,full,s1k,d64,d256 --iters 10 --out "$out/yardstick.json" > "$out/yardstick.log" 2> "$out/yardstick.err"
# profiler passes (never a bench number)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_c2.csv" python bench.py --no-configs --steps 2 --warmup 1 > "$out/ncu_launch.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 3 -c 1 -o "$out/fwd_c2" -f python tools/profile_target.py c2 5 > "$out/ncu_fwd.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 3 -c 1 -o "$out/fwd_c3" -f python tools/varlen_bench.py > "$out/ncu_fwd_c3.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fa_bwd -s 3 -c 3 -o "$out/bwd_c2" -f python tools/profile_target.py c2 2 bwd > "$out/ncu_bwd.log" 2>&1
tail -3 "$out/varlen_c3.log" | cut -c1-200; tail -8 "$out/features.log" | cut -c1-200; tail -12 "$out/decode_c4.log" | cut -c1-220
tail -6 "$out/host_overhead.log" | cut -c1-200; tail -8 "$out/bwd_quick.log" | cut -c1-200
