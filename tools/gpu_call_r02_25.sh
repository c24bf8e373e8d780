#!/usr/bin/env bash
cd "$(dirname "$0")/.."
L=$PWD/flash-attention-v100_b200/lib
for lib in libfa_b200.so libfa_b200_h100k.so ${EXTRA_LIBS:-}; do
echo "== $lib"
FA_B200_LIB=$L/$lib timeout 200 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:fa_fwd_sm100 -s 3 -c 1 python tools/profile_target.py ${PT_ARGS:-c2 5} 2>&1 | grep -E "inst_executed|time_duration|issue_active|tensor_cycles"
done
