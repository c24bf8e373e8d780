#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02_41
L=$PWD/flash-attention-v100_b200/lib
echo "=== parity (seps lib)"; FA_B200_LIB=$L/libfa_b200_seps.so timeout 300 python tests/gpu_quick.py parity 2>&1 | grep -E '"ok": false|rror' | cut -c1-200
echo "=== dense + random tests on the seps lib"
FA_B200_LIB=$L/libfa_b200_seps.so timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_random.py tests/test_gpu_varlen_kvcache.py -m gpu -q -x 2>&1 | tail -3
AB_FILTER='C2_bf16|D64' ROUNDS=3 bash tools/gpu_ab.sh 2>&1 | grep -v "^=== parity" | tee gpurun_out/r02_41/ab.log
