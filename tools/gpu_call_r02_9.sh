#!/usr/bin/env bash
cd "$(dirname "$0")/.."
L=$PWD/flash-attention-v100_b200/lib
F="C2_bf16|_full|S16384|S1024|C3_"
echo "=== parity v2p"; timeout 200 python tests/gpu_quick.py p 2>&1 | grep -E '"ok": false|rror' | cut -c1-200
for r in 1 2; do
echo "=== v1";            FA_B200_FWD_KERNEL=1  QUICK_BENCH_ONLY=1 timeout 200 python tests/gpu_quick.py a 2>&1 | grep -E '"ms"|rror' | grep -E "$F" | cut -c1-110
echo "=== v2p q2kv4";     FA_B200_FWD_KERNEL=2p QUICK_BENCH_ONLY=1 timeout 200 python tests/gpu_quick.py b 2>&1 | grep -E '"ms"|rror' | grep -E "$F" | cut -c1-110
echo "=== v2p q1kv5";     FA_B200_LIB=$L/libfa_b200_q1kv5.so FA_B200_FWD_KERNEL=2p QUICK_BENCH_ONLY=1 timeout 200 python tests/gpu_quick.py d 2>&1 | grep -E '"ms"|rror' | grep -E "$F" | cut -c1-110
done
