#!/usr/bin/env python3
"""BASELINE config 3 timing: 64 packed sequences, randint(1,2049) seed 0, H=32, D=128, bf16 causal."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v100_b200"))
from flash_attn_v100 import flash_attn_varlen_func  # noqa: E402

g = torch.Generator().manual_seed(0)
lens = torch.randint(1, 2049, (64,), generator=g)
H, D = 32, 128
T = int(lens.sum())
torch.manual_seed(421)
q = torch.randn(T, H, D, device="cuda", dtype=torch.bfloat16)
k, v = torch.randn_like(q), torch.randn_like(q)
cu = torch.nn.functional.pad(lens.cumsum(0), (1, 0)).int().cuda()
mx = int(lens.max())
res = []
for Hk in (32, 8):
    kk, vv = k[:, :Hk].contiguous(), v[:, :Hk].contiguous()
    for _ in range(3):
        flash_attn_varlen_func(q, kk, vv, cu, cu, mx, mx, causal=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 30
    e0.record()
    for _ in range(it):
        flash_attn_varlen_func(q, kk, vv, cu, cu, mx, mx, causal=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    flops = 4 * D * H * float((lens.double() * (lens.double() + 1) / 2).sum())
    r = {"config": "C3 varlen 64 seqs", "tokens": T, "Hk": Hk, "ms": ms, "flops": flops, "tflops": flops / ms / 1e9,
         "frac_of_1687.1": flops / ms / 1e9 / 1687.1}
    print(json.dumps(r), flush=True)
    res.append(r)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "varlen_bench.json"), "w"), indent=1)
