#!/usr/bin/env bash
cd "$(dirname "$0")/.."
out=gpurun_out/r02_39; mkdir -p $out
timeout 500 python tools/yardstick.py --shapes c2,full,s1k,d64,d256 --iters 10 --out "$out/yardstick.json" > "$out/yardstick.log" 2> "$out/yardstick.err"
tail -30 $out/yardstick.log | cut -c1-200
