#!/usr/bin/env bash
# bench.py (our arm) once; the JSON line goes to gpurun_out/bench/bench_c2.json
cd "$(dirname "$0")/.."
out=gpurun_out/bench; mkdir -p $out
timeout 600 python bench.py ${BENCH_ARGS:-} > $out/bench_c2.json 2> $out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench/bench_c2.json"))
for k in ("value", "ms_per_step", "roofline", "clocks", "e2e", "gpu_launches"):
    print(k, d.get(k))
for k, v in (d.get("configs") or {}).items():
    print("cfg", k, json.dumps(v)[:300])
print("sustained", json.dumps(d.get("sustained"))[:400])
PY
