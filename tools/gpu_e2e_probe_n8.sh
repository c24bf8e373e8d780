#!/usr/bin/env bash
cd "$(dirname "$0")/.."
out=gpurun_out/e2e_multi; mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1; nproc >> $out/topo.txt; ls -d /sys/devices/system/node/node* >> $out/topo.txt 2>&1
for n in 1 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n tools/e2e_probe_multi.py 2> $out/err_$n.log | grep n_ranks | tee -a $out/e2e_probe_multi.jsonl
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > $out/bench_8gpu.json 2> $out/bench_8gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/e2e_multi/bench_8gpu.json"))
print("N=8 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "C5", (d.get("configs") or {}).get("C5", {}).get("value"), "sustained", (d.get("sustained") or {}).get("value"))
PY
