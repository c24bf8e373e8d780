#!/usr/bin/env bash
# A/B: speculative-max softmax (default library) against -DFA_SPEC_MAX=0; parity first.
cd "$(dirname "$0")/.."
out=gpurun_out/r02_13
mkdir -p "$out"
echo "=== parity (default lib)"; timeout 300 python tests/gpu_quick.py spec 2>&1 | grep -E '"ok": false|rror|"ms"' | cut -c1-160
AB_FILTER="C2_bf16|S1024|_full|C2gqa|C5shard|S16384|D64|C3_" ROUNDS=2 FWD_KERNEL=1 bash tools/ab_v1.sh 2>&1 | tee "$out/ab_spec.log" | cut -c1-120
timeout 600 python -m pytest tests/test_gpu_numerics.py tests/test_gpu_dense.py tests/test_gpu_dropout.py tests/test_gpu_random.py -m gpu -q -x 2>&1 | tail -4 | tee "$out/tests.log"
