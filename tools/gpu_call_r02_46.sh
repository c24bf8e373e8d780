#!/usr/bin/env bash
cd "$(dirname "$0")/.."
out=gpurun_out/r02_46; mkdir -p $out
timeout 400 python tools/yardstick.py --shapes c2,c2gqa --iters 10 --only ours,cudnn,fa4 --out "$out/yardstick_gqa.json" > "$out/yardstick_gqa.log" 2> "$out/yardstick_gqa.err"
python - <<'PY'
import json
for l in open("gpurun_out/r02_46/yardstick_gqa.log"):
    if l.startswith("{"):
        r = json.loads(l); print(r["shape"], r["provider"], round(r.get("fwd_tflops", 0)), round(r.get("bwd_tflops", 0) or 0), r.get("error", "")[:60])
PY
