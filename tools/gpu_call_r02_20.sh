#!/usr/bin/env bash
cd "$(dirname "$0")/.."
out=gpurun_out/r02_20; mkdir -p $out
timeout 200 python tools/feature_bench.py 2>&1 | grep variant | tee $out/features.log
timeout 300 python -m pytest tests/test_gpu_dense.py tests/test_gpu_varlen_kvcache.py tests/test_gpu_crosscheck.py -m gpu -q -x -k "softcap or alibi or cross" 2>&1 | tail -3
bash tools/sanitize_r02.sh 2>&1 | tail -30
cp gpurun_out/sanitizer_r02_*.txt $out/ 2>/dev/null
