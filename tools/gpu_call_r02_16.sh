#!/usr/bin/env bash
cd "$(dirname "$0")/.."
out=gpurun_out/r02_16; mkdir -p $out
for i in 1 2 3; do
bash tools/ncu_capture.sh $out/fwd_c2_shared fa_fwd_sm100 3 python tools/profile_target.py c2 5 | head -6 | tail -2
grep -E "ERROR|passes" $out/fwd_c2_shared.log | head -3
done
