#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page + source page) into a short text: key metrics of every captured launch and,
with a second argument N, the N instructions with the most stall samples per kernel.
    ncu_summary.py <report.ncu-rep> [N]"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        for k in KEYS:
            if k in d:
                print(f"{k} [{units[hdr.index(k)]}] = {d[k]}")
        print("-" * 60)


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def source(rep, top):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    seen = set()
    for a, b in zip(starts, starts[1:]):
        if rows[a][1] in seen:  # one table per view (SASS, then PTX/source-correlated); the first is the SASS one
            continue
        seen.add(rows[a][1])
        hdr = rows[a + 1]
        if "# Samples" not in hdr:  # the SASS table comes first; skip the PTX / CUDA-C views of the same kernel
            continue
        idx = {h: i for i, h in enumerate(hdr)}
        data = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(num(r[idx["# Samples"]]) for r in data) or 1.0
        first = data[0][idx["Source"]] if data else ""
        if not first or first.lstrip().startswith(("//", "#", ".")) or "Address" not in idx:
            continue
        print(f"== {rows[a][1][:100]}  ({len(data)} SASS instructions, {tot:.0f} samples)")
        for r in sorted(data, key=lambda r: -num(r[idx["# Samples"]]))[:top]:
            st = sorted(((num(r[idx[s]]), s) for s in stalls), reverse=True)[:2]
            print(f"{r[idx['Address']][-5:]} {num(r[idx['# Samples']]):7.0f} {100 * num(r[idx['# Samples']]) / tot:5.1f}%  "
                  f"{r[idx['Source']][:64]:64s} {st[0][1]}={st[0][0]:.0f} {st[1][1]}={st[1][0]:.0f}")


if __name__ == "__main__":
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        source(sys.argv[1], int(sys.argv[2]))
