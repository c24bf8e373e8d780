#!/usr/bin/env bash
# compute-sanitizer passes over tools/sanitize_target.py (every entry point, ragged shapes): memcheck, synccheck
# and racecheck. Output goes to gpurun_out/sanitizer_r02_<tool>.txt; summaries are copied to profiles/ by hand.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool "$tool" --print-limit 20 python tools/sanitize_target.py \
    > "gpurun_out/sanitizer_r02_${tool}.txt" 2>&1
  echo "== $tool: exit $?"; tail -n 6 "gpurun_out/sanitizer_r02_${tool}.txt"
done
