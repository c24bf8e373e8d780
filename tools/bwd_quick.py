#!/usr/bin/env python3
"""Backward timings (CUDA events, back-to-back launches) for A/B runs: FA_B200_LIB selects the library variant.
FLOP convention: backward = 2.5 x forward (5 GEMMs against 2), causal counted as S*S/2."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v100_b200"))
import flash_attn_v100_cuda as op  # noqa: E402

CASES = {
    "c2": (8, 32, 32, 4096, 128, True), "full": (8, 32, 32, 4096, 128, False), "c2gqa": (8, 32, 8, 4096, 128, True),
    "s1k": (32, 32, 32, 1024, 128, True), "s16k": (2, 32, 32, 16384, 128, True), "d64": (8, 32, 32, 4096, 64, True),
    "d256": (8, 16, 16, 4096, 256, True),
}


def run(name, iters=10):
    B, H, Hk, S, D, causal = CASES[name]
    torch.manual_seed(421)
    dt = torch.bfloat16
    q = torch.randn(B, S, H, D, device="cuda", dtype=dt).permute(0, 2, 1, 3)  # raw operator layout [B,H,S,D] by strides
    k = torch.randn(B, S, Hk, D, device="cuda", dtype=dt).permute(0, 2, 1, 3)
    v = torch.randn(B, S, Hk, D, device="cuda", dtype=dt).permute(0, 2, 1, 3)
    do = torch.randn(B, S, H, D, device="cuda", dtype=dt).permute(0, 2, 1, 3)
    scale = D ** -0.5
    out, lse, _, rng = op.fwd(q, k, v, None, None, 0.0, scale, causal, -1, -1, 0.0, False, None)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)

    def step():
        op.bwd(do, q, k, v, out, lse, dq, dk, dv, None, 0.0, scale, causal, -1, -1, 0.0, False, None, rng)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.5 * 4 * B * H * S * S * D * (0.5 if causal else 1.0)
    rec = {"name": name, "ms": round(ms, 4), "bwd_tflops": round(flops / ms / 1e9, 1),
           "checksum": float(dq.float().abs().mean() + dk.float().abs().mean() + dv.float().abs().mean())}
    print(json.dumps(rec), flush=True)
    return rec


if __name__ == "__main__":
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2", "full", "c2gqa", "s1k", "d64", "d256"]
    res = [run(n) for n in names]
    tag = os.path.basename(os.environ.get("FA_B200_LIB", "default")).replace(".so", "")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"bwd_quick_{tag}.json"), "w"), indent=1)
