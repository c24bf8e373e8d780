#!/usr/bin/env bash
# bench.py under torchrun on 2 GPUs, launched the way the driver does it (both arms).
cd "$(dirname "$0")/.."
out=gpurun_out/bench_n2; mkdir -p $out
for impl in reference ours; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --impl $impl > $out/bench_${impl}.json 2> $out/bench_${impl}.err
echo "== $impl rc=$?"; tail -c 600 $out/bench_${impl}.json | head -c 600; echo
done
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n2/bench_ours.json"))
for k in ("value", "n_gpus", "ms_per_step", "scaling", "e2e", "gpu_launches", "clocks"):
    print(k, d.get(k))
for k, v in (d.get("configs") or {}).items():
    print("cfg", k, json.dumps(v)[:260])
print("sustained", json.dumps(d.get("sustained"))[:300])
PY
tail -5 $out/bench_ours.err
