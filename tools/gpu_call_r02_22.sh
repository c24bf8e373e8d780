#!/usr/bin/env bash
cd "$(dirname "$0")/.."
out=gpurun_out/r02_22; mkdir -p $out
cat > /tmp/dec8.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import decode_bench
for B in (8, 1):
    print(decode_bench.run(B, 8, iters=3, n_caches=2))
PY
timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size,dram__bytes_read.sum --clock-control none --csv --log-file $out/launches_decode.csv python /tmp/dec8.py > $out/ncu.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r02_22/launches_decode.csv")))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
idx = {h: i for i, h in enumerate(rows[hdr])}
seen = {}
for r in rows[hdr + 1:]:
    if len(r) < len(idx): continue
    k = (r[idx["Kernel Name"]][:70], r[idx["Metric Name"]])
    seen.setdefault(k, []).append(r[idx["Metric Value"]])
for k, v in seen.items():
    if "fa" in k[0] or "kv_" in k[0] or "rotary" in k[0] or "combine" in k[0]:
        print(k, v[-6:])
PY
