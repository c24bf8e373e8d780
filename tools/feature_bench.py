#!/usr/bin/env python3
"""Forward / backward timings of the score-modifier and dropout variants at the config-2 shape (how much each
feature costs next to the plain causal kernel)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v100_b200"))
from flash_attn_v100 import flash_attn_func  # noqa: E402

B, S, H, D = 8, 4096, 32, 128
torch.manual_seed(421)
q, k, v = (torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16, requires_grad=True) for _ in range(3))
do = torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16)
slopes = (torch.rand(H, device="cuda") * 0.2).float()
flops = 4.0 * D * B * H * S * S * 0.5


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


res = []
for name, kw in [("plain causal", {}), ("alibi", dict(alibi_slopes=slopes)), ("softcap 30", dict(softcap=30.0)),
                 ("dropout 0.1", dict(dropout_p=0.1)), ("window (1024, 0)", dict(window_size=(1024, 0)))]:
    with torch.no_grad():
        f_ms = timed(lambda: flash_attn_func(q, k, v, causal=True, **kw))
    out = flash_attn_func(q, k, v, causal=True, **kw)
    b_ms = timed(lambda: torch.autograd.grad(out, (q, k, v), do, retain_graph=True))
    scale = 1.0
    if "window_size" in kw:
        w = kw["window_size"][0]
        scale = (w * (w + 1) / 2 + (S - w) * (w + 1)) / (S * S / 2)
    rec = {"variant": name, "fwd_ms": round(f_ms, 3), "fwd_tflops": round(flops * scale / f_ms / 1e9, 1),
           "bwd_ms": round(b_ms, 3), "bwd_tflops": round(2.5 * flops * scale / b_ms / 1e9, 1)}
    print(json.dumps(rec), flush=True)
    res.append(rec)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "feature_bench.json"), "w"), indent=1)
