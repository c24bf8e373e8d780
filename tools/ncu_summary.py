#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page + source page) into a short text: key metrics, per-role stall map."""
import csv
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def source(rep, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]

    def num(x):
        try:
            return float(x)
        except ValueError:
            return 0.0
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(num(r[idx["# Samples"]]) for r in data)
    lines = [f"total samples {tot:.0f}"]
    for r in sorted(data, key=lambda r: -num(r[idx["# Samples"]]))[:top]:
        st = sorted(((num(r[idx[s]]), s) for s in stalls), reverse=True)[:2]
        lines.append(f"{r[idx['Address']][-5:]} {num(r[idx['# Samples']]):7.0f} {100 * num(r[idx['# Samples']]) / tot:5.1f}%  "
                     f"{r[idx['Source']][:64]:64s} {st[0][1]}={st[0][0]:.0f} {st[1][1]}={st[1][0]:.0f}")
    return "\n".join(lines)


if __name__ == "__main__":
    rep = sys.argv[1]
    vals, units = raw(rep)
    for k in KEYS:
        if k in vals:
            print(f"{k} [{units[k]}] = {vals[k]}")
    if len(sys.argv) > 2:
        print(source(rep, int(sys.argv[2])))
