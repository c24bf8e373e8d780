#!/usr/bin/env bash
# Round-2 evidence pass (one gpurun call): GPU suite, both bench arms, side benches, ncu launch list, and ncu --set full
# captures of the config-2 / config-3 forward and the backward exported as text (the .ncu-rep files are deleted on the
# box: together they exceed what gpurun copies back). Output: gpurun_out/evidence/.
cd "$(dirname "$0")/.."
out=gpurun_out/evidence
mkdir -p "$out"
if [ -z "${SKIP_TESTS:-}" ]; then timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > "$out/tests.log"; tail -2 "$out/tests.log"; fi
timeout 300 python bench.py --impl reference > "$out/bench_reference_arm.json" 2> "$out/bench.err"
timeout 400 python bench.py > "$out/bench_c2.json" 2>> "$out/bench.err"
cut -c1-300 "$out/bench_c2.json"
timeout 100 python tools/varlen_bench.py > "$out/varlen_c3.log" 2>&1; cp gpurun_out/varlen_bench.json "$out/varlen_c3.json" 2>/dev/null
timeout 200 python tools/decode_bench.py > "$out/decode_c4.log" 2>&1; cp gpurun_out/decode_bench.json "$out/decode_c4.json" 2>/dev/null
timeout 100 python tools/bwd_quick.py > "$out/bwd_quick.log" 2>&1
timeout 200 python tools/feature_bench.py > "$out/features.log" 2>&1; cp gpurun_out/feature_bench.json "$out/features.json" 2>/dev/null
timeout 100 python tools/host_overhead.py > "$out/host_overhead.log" 2>&1
timeout 400 python tools/yardstick.py --shapes c2,full,s1k,d64,d256 --iters 10 --out "$out/yardstick.json" > "$out/yardstick.log" 2> "$out/yardstick.err"
# profiler passes (never a bench number)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_c2.csv" python bench.py --no-configs --steps 2 --warmup 1 > "$out/ncu_launch.log" 2>&1
bash tools/ncu_capture.sh "$out/ncu_fwd_c2" fa_fwd_sm100 3 python tools/profile_target.py c2 5 > /dev/null
bash tools/ncu_capture.sh "$out/ncu_fwd_c3" fa_fwd_sm100 3 python tools/varlen_bench.py > /dev/null
bash tools/ncu_capture.sh "$out/ncu_bwd_dkdv_c2" "fa_bwd_sm100.*" 4 python tools/profile_target.py c2 3 bwd > /dev/null
bash tools/ncu_capture.sh "$out/ncu_bwd_dq_c2" "fa_bwd_sm100.*" 5 python tools/profile_target.py c2 3 bwd > /dev/null
rm -f "$out"/*.source.csv.tmp
# SASS evidence: tcgen05 / TMA mnemonic counts of the shipped library
cuobjdump -sass flash-attention-v100_b200/lib/libfa_b200.so | grep -oE "UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|LDTM|STTM|UTCBAR|SYNCS|UTMAPF|MUFU\.EX2|FFMA2|HMMA|QGMMA" | sort | uniq -c > "$out/sass_mnemonics.txt"
cat "$out/sass_mnemonics.txt"
tail -2 "$out/varlen_c3.log" | cut -c1-200; tail -6 "$out/features.log" | cut -c1-200; tail -6 "$out/decode_c4.log" | cut -c1-220
tail -6 "$out/host_overhead.log" | cut -c1-200; tail -7 "$out/bwd_quick.log" | cut -c1-200; head -12 "$out/ncu_fwd_c2.summary.txt"
du -sh "$out"
# the opt-in forward v2 (three S buffers, 128-row CTAs / CTA pairs) must still be parity-clean
for m in 2s 2p; do echo "== FA_B200_FWD_KERNEL=$m"; FA_B200_FWD_KERNEL=$m QUICK_PARITY_ONLY=1 timeout 200 python tests/gpu_quick.py v$m 2>&1 | grep -E '"ok": false|rror|elapsed' | cut -c1-160; done | tee "$out/fwd_v2_parity.log"
