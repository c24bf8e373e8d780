#!/usr/bin/env bash
# decode: KV-split sweep at B = 1 and 8 (FA_B200_DECODE_SPLITS), graph-replay time per step
cd "$(dirname "$0")/.."
out=gpurun_out/r02_23; mkdir -p $out
cat > /tmp/dec_sweep.py <<'PY'
import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import decode_bench
for B in (int(x) for x in os.environ.get("BATCHES", "8,1").split(",")):
    r = decode_bench.run(B, 8, iters=20, n_caches=4)
    print(json.dumps({"splits": os.environ.get("FA_B200_DECODE_SPLITS", "auto"), "B": B, "us": round(r["us"], 1), "graph_us": round(r["graph_us"], 1), "graph_GBps": round(r["graph_GBps"])}))
PY
for s in auto 2 3 4 6 9 13 16 18 32; do
  if [ $s = auto ]; then unset FA_B200_DECODE_SPLITS; else export FA_B200_DECODE_SPLITS=$s; fi
  timeout 120 python /tmp/dec_sweep.py 2>&1 | grep splits
done | tee $out/decode_split_sweep.log
