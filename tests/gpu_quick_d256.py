#!/usr/bin/env python3
"""Quick on-GPU check of the head_dim 129..256 forward (the 256-wide tile): parity against the CPU oracle on a few shapes,
then a coarse timing. Bring-up checker like tests/gpu_quick.py (not collected by pytest; lives under tests/ because it
imports oracle/). usage: FA_B200_LIB=... python tests/gpu_quick_d256.py [tag]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v100_b200"))
from flash_attn_v100 import flash_attn_func, flash_attn_with_kvcache  # noqa: E402
from oracle.attention_oracle import flash_attn_func_ref  # noqa: E402

bf16, f16 = torch.bfloat16, torch.float16


def parity(name, B, Sq, Sk, H, Hk, D, dt, causal, window=(-1, -1), softcap=0.0, alibi=False):
    torch.manual_seed(11)
    q = torch.randn(B, Sq, H, D, device="cuda", dtype=dt)
    k = torch.randn(B, Sk, Hk, D, device="cuda", dtype=dt)
    v = torch.randn(B, Sk, Hk, D, device="cuda", dtype=dt)
    slopes = (torch.rand(H, device="cuda", dtype=torch.float32) * 0.3) if alibi else None
    rec = {"name": name}
    try:
        out = flash_attn_func(q, k, v, causal=causal, window_size=window, softcap=softcap, alibi_slopes=slopes)
        torch.cuda.synchronize()
        ref, _ = flash_attn_func_ref(q, k, v, causal=causal, window_size=window, softcap=softcap, alibi_slopes=slopes)
        err = (out.double().cpu() - ref).abs().max().item()
        rec.update(max_err=err, ok=bool(err < 2e-2 and torch.isfinite(out.float()).all().item()))
    except Exception as ex:  # noqa: BLE001
        rec.update(ok=False, error=f"{type(ex).__name__}: {ex}")
    print(json.dumps(rec), flush=True)
    return rec


def parity_decode(name, B, Sk, H, Hk, D):
    torch.manual_seed(12)
    q = torch.randn(B, 1, H, D, device="cuda", dtype=bf16)
    kc = torch.randn(B, Sk, Hk, D, device="cuda", dtype=bf16)
    vc = torch.randn(B, Sk, Hk, D, device="cuda", dtype=bf16)
    lens = torch.randint(Sk // 2, Sk + 1, (B,), device="cuda", dtype=torch.int32)
    rec = {"name": name}
    try:
        out = flash_attn_with_kvcache(q, kc, vc, cache_seqlens=lens, causal=True)
        torch.cuda.synchronize()
        err = 0.0
        for b in range(B):
            n = int(lens[b])
            ref, _ = flash_attn_func_ref(q[b:b + 1], kc[b:b + 1, :n], vc[b:b + 1, :n], causal=True)
            err = max(err, (out[b:b + 1].double().cpu() - ref).abs().max().item())
        rec.update(max_err=err, ok=bool(err < 2e-2))
    except Exception as ex:  # noqa: BLE001
        rec.update(ok=False, error=f"{type(ex).__name__}: {ex}")
    print(json.dumps(rec), flush=True)
    return rec


def bench(name, B, S, H, Hk, D, dt, causal, iters=20):
    torch.manual_seed(421)
    q = torch.randn(B, S, H, D, device="cuda", dtype=dt)
    k = torch.randn(B, S, Hk, D, device="cuda", dtype=dt)
    v = torch.randn(B, S, Hk, D, device="cuda", dtype=dt)
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
    rec = {"name": name, "ms": round(ms, 4), "tflops": round(4 * B * H * S * S * D * (0.5 if causal else 1.0) / ms / 1e9, 1)}
    print(json.dumps(rec), flush=True)
    return rec


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "default"
    res = []
    if not os.environ.get("QUICK_BENCH_ONLY"):
        res += [parity("d256_bf16_causal_1024", 2, 1024, 1024, 4, 4, 256, bf16, True),
                parity("d256_f16_full_640", 1, 640, 640, 2, 1, 256, f16, False),
                parity("d256_ragged_333x777", 2, 333, 777, 4, 2, 256, bf16, True),
                parity("d192_causal_500", 1, 500, 500, 3, 3, 192, bf16, True),
                parity("d160_full_129", 1, 129, 129, 2, 2, 160, f16, False),
                parity("d256_window", 1, 1536, 1536, 2, 2, 256, bf16, True, window=(300, 0)),
                parity("d256_alibi_softcap", 1, 777, 777, 4, 4, 256, bf16, True, softcap=30.0, alibi=True),
                parity("d256_sq1_sk4000", 2, 1, 4000, 8, 2, 256, bf16, True),
                parity("d256_many_items", 2, 2048, 2048, 40, 8, 256, bf16, True),
                parity("d136_causal_700", 1, 700, 700, 2, 2, 136, bf16, True),
                parity("d192_ragged_333x777_gqa", 2, 333, 777, 4, 2, 192, f16, True),
                parity("d184_window_softcap", 1, 900, 900, 2, 1, 184, bf16, True, window=(200, 0), softcap=20.0),
                parity_decode("d192_decode_B3_3000", 3, 3000, 8, 2, 192),
                parity_decode("d256_decode_B4_8192", 4, 8192, 16, 4, 256),
                parity_decode("d256_decode_B1_1000", 1, 1000, 8, 8, 256)]
    res += [bench("D256_bf16_B8_H16_S4096_causal", 8, 4096, 16, 16, 256, bf16, True),
            bench("D256_bf16_B8_H16_S4096_full", 8, 4096, 16, 16, 256, bf16, False),
            bench("D192_bf16_B8_H16_S4096_causal", 8, 4096, 16, 16, 192, bf16, True),
            bench("D192_bf16_B8_H16_S4096_full", 8, 4096, 16, 16, 192, bf16, False),
            bench("D160_bf16_B8_H16_S4096_causal", 8, 4096, 16, 16, 160, bf16, True),
            bench("D256_bf16_B32_H16_S1024_causal", 32, 1024, 16, 16, 256, bf16, True, iters=50)]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"quick_d256_{tag}.json"), "w"), indent=1)
