#!/usr/bin/env python3
"""Stall-reason digest of an ncu source-page CSV (tools/ncu_capture.sh): per code segment (runs of SASS instructions with the
same execution count) the sample share and the top stall reasons, then the N instructions with the most samples.
    ncu_stalls.py <x.source.csv> [N] [min segment length]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 15
hi = [i for i, r in enumerate(rows) if "# Samples" in r][0]
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


tot = sum(num(r[idx["# Samples"]]) for r in data)
ex = [num(r[idx["Instructions Executed"]]) for r in data]
start = 0
print(f"{len(data)} SASS instructions, {tot:.0f} samples")
for i in range(1, len(ex) + 1):
    if i == len(ex) or ex[i] != ex[start]:
        if i - start >= minlen:
            c = collections.Counter()
            s = 0
            for r in data[start:i]:
                s += num(r[idx["# Samples"]])
                for st in stalls:
                    c[st] += num(r[idx[st]])
            if s / tot > 0.003:
                print(f"instr {start:5d}-{i:5d} exec/instr {ex[start]:9.0f} samples {s:6.0f} ({100 * s / tot:4.1f}%) | " +
                      ", ".join(f"{k[6:]}={v:.0f}" for k, v in c.most_common(7) if v > 0))
        start = i
print("--- top instructions")
order = sorted(range(len(data)), key=lambda i: -num(data[i][idx["# Samples"]]))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((num(r[idx[s]]), s[6:]) for s in stalls), reverse=True)[:3]
    print(f"{i:5d} {num(r[idx['# Samples']]):6.0f} {100 * num(r[idx['# Samples']]) / tot:4.1f}% exec {ex[i]:8.0f}  {r[idx['Source']][:60]:60s} " +
          " ".join(f"{n}={v:.0f}" for v, n in st if v > 0))
