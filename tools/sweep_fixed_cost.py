#!/usr/bin/env python3
"""Fit time-per-CTA = a + b * n_tiles from a non-causal sweep at constant total tokens (bring-up tool)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "flash-attention-v100_b200"))
from flash_attn_v100 import flash_attn_func
H, D = 32, 128
rows = []
for causal in (False, True):
    for S in (256, 512, 1024, 2048, 4096, 8192, 16384):
        B = max(1, 32768 // S)
        q = torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16); k = torch.randn_like(q); v = torch.randn_like(q)
        for _ in range(3): flash_attn_func(q, k, v, causal=causal)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 10
        e0.record()
        for _ in range(it): flash_attn_func(q, k, v, causal=causal)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        ctas = B * H * ((S + 255) // 256)
        tiles = S // 128
        fl = 4 * B * H * S * S * D * (0.5 if causal else 1)
        r = dict(causal=causal, S=S, B=B, ms=ms, tflops=fl / ms / 1e9, ctas=ctas, us_per_cta=ms * 1e3 * 148 / ctas, tiles_noncausal=tiles)
        print(json.dumps(r), flush=True); rows.append(r)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "sweep_fixed_cost.json"), "w"), indent=1)
