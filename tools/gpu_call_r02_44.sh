#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02_44
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r02_44/gpu_tests.txt
timeout 300 python tests/gpu_quick_d256.py final 2>&1 | grep -E '"name"|rror' | cut -c1-160 | tee gpurun_out/r02_44/d256_quick.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash tools/gpu_bench.sh 2>&1 | tee gpurun_out/r02_44/bench_summary.txt
cp gpurun_out/bench/bench_c2.json gpurun_out/r02_44/
