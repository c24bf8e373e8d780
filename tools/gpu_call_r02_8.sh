#!/usr/bin/env bash
cd "$(dirname "$0")/.."
T=$PWD/flash-attention-v100_b200/lib/libfa_b200_trace.so
echo "=== v2 single q1kv5, non-causal 4096"; FA_B200_FWD_KERNEL=2s FA_B200_LIB=$T timeout 120 python tools/trace_timeline.py 0 4096 400 2>&1 | tail -75
