#!/usr/bin/env python3
"""A handful of small, ragged calls over every entry point -- the target for
`compute-sanitizer --tool memcheck python tools/sanitize_target.py` (out-of-bounds / misaligned global accesses)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v100_b200"))
from flash_attn_v100 import flash_attn_func, flash_attn_varlen_func, flash_attn_with_kvcache  # noqa: E402

torch.manual_seed(0)
dt = torch.bfloat16
for D in (40, 128, 192, 256):
    q = torch.randn(1, 77, 2, D, device="cuda", dtype=dt, requires_grad=True)
    k = torch.randn(1, 203, 1, D, device="cuda", dtype=dt, requires_grad=True)
    v = torch.randn(1, 203, 1, D, device="cuda", dtype=dt, requires_grad=True)
    o = flash_attn_func(q, k, v, causal=True)
    torch.autograd.grad(o, (q, k, v), torch.randn_like(o))
    o = flash_attn_func(q, k, v, dropout_p=0.1, window_size=(50, 10))
    torch.autograd.grad(o, (q, k, v), torch.randn_like(o))
lens = [3, 130, 1, 257]
cu = torch.tensor([0, 3, 133, 134, 391], dtype=torch.int32, device="cuda")
q = torch.randn(sum(lens), 4, 64, device="cuda", dtype=dt, requires_grad=True)
k = torch.randn(sum(lens), 2, 64, device="cuda", dtype=dt, requires_grad=True)
v = torch.randn(sum(lens), 2, 64, device="cuda", dtype=dt, requires_grad=True)
o = flash_attn_varlen_func(q, k, v, cu, cu, max(lens), max(lens), causal=True)
torch.autograd.grad(o, (q, k, v), torch.randn_like(o))
kc = torch.randn(6, 256, 2, 128, device="cuda", dtype=dt)
vc = torch.randn(6, 256, 2, 128, device="cuda", dtype=dt)
bt = torch.tensor([[4, 1, 0], [2, 5, 3]], dtype=torch.int32, device="cuda")
ls = torch.tensor([700, 5], dtype=torch.int32, device="cuda")
for Sq in (1, 3):
    qd = torch.randn(2, Sq, 8, 128, device="cuda", dtype=dt)
    kn = torch.randn(2, Sq, 2, 128, device="cuda", dtype=dt)
    flash_attn_with_kvcache(qd, kc, vc, kn, torch.randn_like(kn), cache_seqlens=ls, block_table=bt, causal=True)
torch.cuda.synchronize()
print("sanitize target done")
