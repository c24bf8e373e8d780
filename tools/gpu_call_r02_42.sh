#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02_42
timeout 600 python -m pytest tests -m gpu -q -x -k "multi or device or two_gpu or distributed or peer" 2>&1 | tail -3 | tee gpurun_out/r02_42/gpu_tests_2gpu.txt
out=gpurun_out/r02_42
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $out/bench_2gpu.json 2> $out/bench_2gpu.err
echo "rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_42/bench_2gpu.json"))
for k in ("value", "n_gpus", "ms_per_step", "scaling", "e2e", "gpu_launches", "clocks"):
    print(k, d.get(k))
PY
