#!/usr/bin/env bash
cd "$(dirname "$0")/.."
out=gpurun_out/r02_21; mkdir -p $out
L=$PWD/flash-attention-v100_b200/lib
for r in 1 2; do
for lib in libfa_b200.so libfa_b200_prev.so; do
echo "=== $lib"; FA_B200_LIB=$L/$lib timeout 200 python tools/decode_bench.py 2>&1 | grep '"B"' | sed -E 's/"caches_rotated.*graph_GBps": ([0-9.]+).*/"graph_GBps": \1/' | cut -c1-150
done; done | tee $out/decode_ab.log
timeout 600 python -m pytest tests/test_gpu_varlen_kvcache.py tests/test_gpu_graphs.py tests/test_gpu_crosscheck.py -m gpu -q -x 2>&1 | tail -3
