#!/usr/bin/env python3
"""Prints the event timeline of CTA 0 from a -DFA_TRACE build (libfa_b200_trace.so). Bring-up tool.
usage: FA_B200_LIB=.../libfa_b200_trace.so python tools/trace_timeline.py [causal 0/1] [S]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v100_b200"))
import flash_attn_v100_cuda as op  # noqa: E402
from flash_attn_v100 import flash_attn_func  # noqa: E402

causal = bool(int(sys.argv[1])) if len(sys.argv) > 1 else False
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
lib = op.load_library()
buf = torch.zeros(16 * 4096, dtype=torch.int64, device="cuda")
q = torch.randn(8, S, 32, 128, device="cuda", dtype=torch.bfloat16)
k, v = torch.randn_like(q), torch.randn_like(q)
for _ in range(3):
    flash_attn_func(q, k, v, causal=causal)
torch.cuda.synchronize()
assert lib.fa_b200_debug_set_trace(ctypes.c_void_p(buf.data_ptr())) == 0
flash_attn_func(q, k, v, causal=causal)
torch.cuda.synchronize()
t = buf.cpu().view(16, 2048, 2)
NAMES = {6: "max done", 7: "prev PV done", 1: "S seen", 2: "S in regs", 3: "max+stats", 4: "P 3/4", 5: "P all", 100: "mma P0 seen", 101: "mma P1 seen",
         110: "mma P0 last", 111: "mma P1 last", 120: "mma QK0 issued", 121: "mma QK1 issued", 200: "corr stats0",
         201: "corr stats1", 210: "corr O0 final", 211: "corr O1 final", 220: "corr epi0 done", 221: "corr epi1 done", 230: "corr got id", 231: "corr geom", 400: "ask work", 401: "work slot full", 240: "corr last chunk0", 241: "corr last chunk1", 242: "corr tma read done0", 243: "corr tma read done1", 130: "mma new item", 131: "mma Q0 landed", 132: "mma K0,Q1 landed", 300: "load wait qempty", 301: "load q free", 302: "load first issued",
         102: "mma P2 seen", 122: "mma QK2 issued", 140: "mma K landed", 141: "mma V landed", 142: "mma PV issued",
         310: "load slot free", 311: "load issued", 303: "load id published"}
events = []
for w in (0, 4, 8, 12, 13):
    for i in range(2048):
        ev, clk = int(t[w, i, 0]), int(t[w, i, 1])
        if clk == 0:
            break
        events.append((clk, w, ev))
events.sort()
t0 = events[0][0]
out = []
for clk, w, ev in events:
    out.append(f"{clk - t0:9d}  warp {w:2d}  {NAMES.get(ev, ev)}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", f"timeline_{'causal' if causal else 'full'}_{S}.txt"), "w").write("\n".join(out))
# per-warp deltas for softmax warps: average phase durations over steady-state steps
import statistics
for w in (0, 4):
    ev = [(int(t[w, i, 0]), int(t[w, i, 1])) for i in range(2048) if int(t[w, i, 1]) != 0]
    seq = {}
    for (e, c) in ev:
        seq.setdefault(e, []).append(c)
    n = min(len(seq.get(e, [])) for e in (1, 2, 3, 4, 5))
    if n <= 12 and all(len(seq.get(e, [])) > 12 for e in (1, 2, 3, 5)):  # forward v2: no 3/4-P event
        n2 = min(len(seq[e]) for e in (1, 2, 3, 5))
        sl = slice(4, min(n2, 30) - 1)
        d = lambda a, b: statistics.mean([y - x for x, y in zip(seq[a][sl], seq[b][sl])])
        period = statistics.mean([y - x for x, y in zip(seq[1][4:min(n2, 30) - 1], seq[1][5:min(n2, 30)])])
        print(f"warp {w}: S seen->regs {d(1, 2):.0f}  regs->max+handoff {d(2, 3):.0f}  max->P all {d(3, 5):.0f}  "
              f"active {d(1, 5):.0f}  period {period:.0f}  wait-for-S {period - d(1, 5):.0f}")
    if n > 12:
        sl = slice(4, min(n, 30) - 1)
        d = lambda a, b: statistics.mean([y - x for x, y in zip(seq[a][sl], seq[b][sl])])
        period = statistics.mean([y - x for x, y in zip(seq[1][4:min(n, 30) - 1], seq[1][5:min(n, 30)])])
        print(f"warp {w}: S seen->regs {d(1, 2):.0f}  regs->max {d(2, 3):.0f}  max->P3/4 {d(3, 4):.0f}  P3/4->Pall {d(4, 5):.0f}  "
              f"active {d(1, 5):.0f}  period {period:.0f}  wait-for-S {period - d(1, 5):.0f}")
mm = [(int(t[12, i, 0]), int(t[12, i, 1])) for i in range(2048) if int(t[12, i, 1]) != 0]
print("mma events", len(mm))
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 200
print("\n".join(out[lo:lo + 90]))
