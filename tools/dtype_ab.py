#!/usr/bin/env python3
"""fp16 against bf16 forward at the same shapes (bring-up tool). usage: python tools/dtype_ab.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v100_b200"))
from flash_attn_v100 import flash_attn_func  # noqa: E402


def bench(B, S, H, D, dt, causal, iters=20):
    torch.manual_seed(421)
    q, k, v = (torch.randn(B, S, H, D, device="cuda", dtype=dt) for _ in range(3))
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
    for (B, S, H, D, causal) in ((8, 4096, 32, 128, True), (8, 4096, 32, 128, False), (8, 4096, 32, 64, True), (8, 4096, 64, 64, True), (8, 4096, 16, 256, True)):
        r = {"B": B, "S": S, "H": H, "D": D, "causal": causal}
        for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
            r[name] = bench(B, S, H, D, dt, causal)
        print(json.dumps(r), flush=True)
