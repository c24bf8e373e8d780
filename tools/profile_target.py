#!/usr/bin/env python3
"""Minimal launch target for ncu: runs one workload of bench.py a few times (no timing, no CPU arm).
    profile_target.py <workload> <n> [bwd]     -- with `bwd` every iteration also runs the backward."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v100_b200"))
from bench import WORKLOADS  # noqa: E402
from flash_attn_v100 import flash_attn_func  # noqa: E402

w = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
torch.manual_seed(421)
q = torch.randn(w["batch"], w["seqlen"], w["heads"], w["head_dim"], device="cuda", dtype=torch.bfloat16)
k = torch.randn(w["batch"], w["seqlen"], w["heads_k"], w["head_dim"], device="cuda", dtype=torch.bfloat16)
v = torch.randn_like(k)
bwd = len(sys.argv) > 3 and sys.argv[3] == "bwd"
if bwd:
    for t in (q, k, v):
        t.requires_grad_(True)
    do = torch.randn_like(q)
for _ in range(n):
    o = flash_attn_func(q, k, v, causal=w["causal"], window_size=w["window"])
    if bwd:
        torch.autograd.grad(o, (q, k, v), do)
torch.cuda.synchronize()
print("done", float(o.float().abs().mean()))
