#!/usr/bin/env python3
"""Forward throughput against the memory layout of q/k/v (bring-up tool): the API takes (batch, seq, heads, dim) tensors with
arbitrary batch / seq / head strides; `bshd` is the contiguous layout, `bhsd` a (batch, heads, seq, dim)-contiguous buffer
passed as a transposed view (rows of one head 2*D bytes apart instead of 2*H*D)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v100_b200"))
from flash_attn_v100 import flash_attn_func  # noqa: E402


def bench(B, S, H, Hk, D, layout, causal=True, iters=20):
    torch.manual_seed(421)
    def mk(h):
        if layout == "bshd":
            return torch.randn(B, S, h, D, device="cuda", dtype=torch.bfloat16)
        return torch.randn(B, h, S, D, device="cuda", dtype=torch.bfloat16).transpose(1, 2)
    q, k, v = mk(H), mk(Hk), mk(Hk)
    for _ in range(3):
        flash_attn_func(q, k, v, causal=causal)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        flash_attn_func(q, k, v, causal=causal)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return round(4 * B * H * S * S * D * (0.5 if causal else 1.0) / ms / 1e9, 1)


for rnd in range(2):
    for (B, S, H, Hk, D) in ((16, 4096, 16, 16, 128), (8, 4096, 32, 32, 128), (4, 4096, 64, 64, 128), (8, 4096, 32, 8, 128),
                             (8, 4096, 32, 32, 64), (8, 4096, 64, 64, 64), (8, 4096, 16, 16, 256)):
        r = {"B": B, "S": S, "H": H, "Hk": Hk, "D": D}
        for layout in ("bshd", "bhsd"):
            r[layout] = bench(B, S, H, Hk, D, layout)
        print(json.dumps(r), flush=True)
