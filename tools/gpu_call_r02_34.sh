#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02_34
L=$PWD/flash-attention-v100_b200/lib
echo "=== varlen / random tests on the gskip lib"
FA_B200_LIB=$L/libfa_b200_gskip.so timeout 600 python -m pytest tests/test_gpu_varlen_kvcache.py tests/test_gpu_random.py -m gpu -q -x 2>&1 | tail -3
AB_FILTER='C2_bf16|S1024|C3_' ROUNDS=3 bash tools/gpu_ab.sh 2>&1 | grep -v "^=== parity" | tee gpurun_out/r02_34/ab.log
for r in 1 2; do for t in default gskip; do
  lib=$L/libfa_b200.so; [ $t = gskip ] && lib=$L/libfa_b200_gskip.so
  echo "=== varlen_bench $t"; FA_B200_LIB=$lib timeout 120 python tools/varlen_bench.py 2>&1 | tail -4 | cut -c1-300
done; done 2>&1 | tee gpurun_out/r02_34/varlen.log
