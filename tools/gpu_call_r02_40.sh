#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02_40
for tool in memcheck synccheck; do
  timeout 600 compute-sanitizer --tool "$tool" --print-limit 20 python tools/sanitize_target.py > "gpurun_out/r02_40/sanitizer_${tool}.txt" 2>&1
  echo "== $tool: exit $?"; tail -n 4 "gpurun_out/r02_40/sanitizer_${tool}.txt"
done
