#!/usr/bin/env python3
"""Forward / backward attention TFLOP/s of this repo beside the libraries that happen to be in the image
(cuDNN through torch SDPA, flash_attn 2.8, the FA4 CuTe-DSL kernels bundled with vllm, flashinfer).
A yardstick only: none of these libraries is on the product path.

    python tools/yardstick.py [--shapes c2,full,d64,d256] [--iters 20] [--only ours,cudnn,...]

FLOP convention: forward 4*D*pairs, backward 2.5x the forward (5 GEMMs against 2), causal = S*S/2.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHAPES = {
    # name: (B, H, Hk, S, D, causal)
    "c2": (8, 32, 32, 4096, 128, True),
    "c2gqa": (8, 32, 8, 4096, 128, True),
    "full": (8, 32, 32, 4096, 128, False),
    "s8k": (4, 32, 32, 8192, 128, True),
    "s1k": (32, 32, 32, 1024, 128, True),
    "d64": (8, 32, 32, 4096, 64, True),
    "d256": (8, 16, 16, 4096, 256, True),
    "d32": (8, 32, 32, 4096, 32, True),
}


def time_fn(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def providers():
    """name -> callable(q, k, v, causal) returning out (layout (B,S,H,D)); import errors are reported, not fatal."""
    out = {}

    def ours():
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "flash-attention-v100_b200"))
        from flash_attn_v100 import flash_attn_func

        return lambda q, k, v, causal: flash_attn_func(q, k, v, causal=causal)

    def cudnn():
        from torch.nn.attention import SDPBackend, sdpa_kernel
        import torch.nn.functional as F

        def f(q, k, v, causal):
            with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
                return F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                                      is_causal=causal, enable_gqa=q.shape[2] != k.shape[2]).transpose(1, 2)

        return f

    def fa2():
        import importlib

        mod = importlib.import_module("flash_attn")  # the site-packages wheel (the repo's shim is not on sys.path here)
        assert "site-packages" in (mod.__file__ or ""), mod.__file__
        return lambda q, k, v, causal: mod.flash_attn_func(q, k, v, causal=causal)

    def fa4():
        from vllm.vllm_flash_attn.cute.interface import flash_attn_func as f4

        def f(q, k, v, causal):
            r = f4(q, k, v, causal=causal)
            return r[0] if isinstance(r, tuple) else r

        return f

    for name, mk in (("ours", ours), ("cudnn", cudnn), ("fa2", fa2), ("fa4", fa4)):
        try:
            out[name] = mk()
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"provider": name, "error": f"{type(e).__name__}: {e}"[:300]}), flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="c2,full")
    ap.add_argument("--iters", type=int, default=15)
    ap.add_argument("--only", default="")
    ap.add_argument("--no-bwd", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "yardstick.json"))
    args = ap.parse_args()
    provs = providers()
    if args.only:
        provs = {k: v for k, v in provs.items() if k in args.only.split(",")}
    results = []
    for sname in args.shapes.split(","):
        B, H, Hk, S, D, causal = SHAPES[sname]
        torch.manual_seed(421)
        q = torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        k = torch.randn(B, S, Hk, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        v = torch.randn(B, S, Hk, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        do = torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16)
        fwd_flops = 4.0 * D * B * H * S * S * (0.5 if causal else 1.0)
        ref_out = None
        for name, fn in provs.items():
            rec = {"shape": sname, "provider": name, "B": B, "H": H, "Hk": Hk, "S": S, "D": D, "causal": causal}
            try:
                with torch.no_grad():
                    o = fn(q, k, v, causal)
                    if ref_out is None:
                        ref_out = o.float()
                    else:
                        rec["max_abs_diff_vs_first"] = (o.float() - ref_out).abs().max().item()
                    med, best = time_fn(lambda: fn(q, k, v, causal), args.iters)
                rec.update(fwd_ms=med, fwd_tflops=fwd_flops / med / 1e9, fwd_tflops_best=fwd_flops / best / 1e9)
                if not args.no_bwd:
                    o = fn(q, k, v, causal)

                    def bwd():
                        q.grad = k.grad = v.grad = None
                        o.backward(do, retain_graph=True)

                    med, best = time_fn(bwd, args.iters)
                    rec.update(bwd_ms=med, bwd_tflops=2.5 * fwd_flops / med / 1e9, bwd_tflops_best=2.5 * fwd_flops / best / 1e9)
                    q.grad = k.grad = v.grad = None
                    del o
            except Exception as e:  # noqa: BLE001
                rec["error"] = f"{type(e).__name__}: {e}"[:400]
                traceback.print_exc(file=sys.stderr)
            print(json.dumps(rec), flush=True)
            results.append(rec)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
