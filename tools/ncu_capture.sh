#!/usr/bin/env bash
# ncu --set full capture of one launch, exported as text on the GPU box (the .ncu-rep itself is too large to bring
# back): raw metrics page, per-SASS-instruction source page, and tools/ncu_summary.py's digest.
#   ncu_capture.sh <out prefix> <kernel regex> <skip launches> <command ...>
out=$1; re=$2; skip=$3; shift 3
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$re" -s "$skip" -c 1 -o "$out" -f "$@" > "$out.log" 2>&1
ncu -i "$out.ncu-rep" --page raw --csv > "$out.raw.csv" 2>/dev/null
ncu -i "$out.ncu-rep" --page source --csv --print-source sass > "$out.source.csv" 2>/dev/null
python tools/ncu_summary.py "$out.ncu-rep" 40 > "$out.summary.txt" 2>&1
rm -f "$out.ncu-rep"
head -25 "$out.summary.txt"
