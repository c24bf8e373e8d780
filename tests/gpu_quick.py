#!/usr/bin/env python3
"""Quick on-GPU diagnostic: a handful of shapes against the CPU oracle plus a coarse timing.
Writes gpurun_out/quick.json. Not collected by pytest and not the bench -- a bring-up checker; it lives under tests/ because only tests/,
smoke() and the bench's cpu_baseline leg may import oracle/."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v100_b200"))
from flash_attn_v100 import flash_attn_func  # noqa: E402
from oracle.attention_oracle import (fa_tolerance_ok, flash_attn_func_ref, naive_lowp_attention,  # noqa: E402
                                     normalize_mask_args)

results = []


def run_case(name, B, Sq, Sk, H, Hk, D, dtype, causal, window=(-1, -1), softcap=0.0, alibi=False):
    torch.manual_seed(421)
    dev = "cuda"
    q = torch.randn(B, Sq, H, D, device=dev, dtype=dtype)
    k = torch.randn(B, Sk, Hk, D, device=dev, dtype=dtype)
    v = torch.randn(B, Sk, Hk, D, device=dev, dtype=dtype)
    slopes = (torch.rand(H, device=dev, dtype=torch.float32) * 0.3) if alibi else None
    rec = {"name": name}
    try:
        out, lse, _ = None, None, None
        res = flash_attn_func(q, k, v, causal=causal, window_size=window, softcap=softcap, alibi_slopes=slopes)
        torch.cuda.synchronize()
        out = res
        ref, lse_ref = flash_attn_func_ref(q, k, v, causal=causal, window_size=window, softcap=softcap,
                                           alibi_slopes=slopes)
        wl, wr = normalize_mask_args(Sq, Sk, causal, window, alibi)
        naive = naive_lowp_attention(q, k, v, D ** -0.5, wl, wr) if (softcap == 0 and not alibi) else None
        err = (out.double().cpu() - ref).abs().max().item()
        rec["max_err"] = err
        rec["finite"] = bool(torch.isfinite(out.float()).all().item())
        if naive is not None:
            ok, e, en = fa_tolerance_ok(out, ref, naive)
            rec.update(ok=ok, err_naive=en)
        else:
            rec["ok"] = err < 2e-2
        # per-row-block error map to localise layout bugs
        e_rows = (out.double().cpu() - ref).abs().amax(dim=(0, 2, 3))
        blk = min(64, Sq)
        rec["err_by_rowblock"] = [round(x, 4) for x in e_rows[: Sq // blk * blk].view(-1, blk).amax(dim=1).tolist()[:16]]
    except Exception as ex:  # noqa: BLE001
        rec["error"] = f"{type(ex).__name__}: {ex}"
    print(json.dumps(rec), flush=True)
    results.append(rec)


def bench_case(name, B, S, H, Hk, D, dtype, causal, iters=10, window=(-1, -1)):
    dev = "cuda"
    torch.manual_seed(421)
    q = torch.randn(B, S, H, D, device=dev, dtype=dtype)
    k = torch.randn(B, S, Hk, D, device=dev, dtype=dtype)
    v = torch.randn(B, S, Hk, D, device=dev, dtype=dtype)
    for _ in range(3):
        flash_attn_func(q, k, v, causal=causal, window_size=window)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        flash_attn_func(q, k, v, causal=causal, window_size=window)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 4 * B * H * S * S * D * (0.5 if causal else 1.0)
    if window[0] >= 0:
        w = window[0]
        flops = 4 * B * H * D * (w * (w + 1) // 2 + (S - w) * (w + 1))
    rec = {"name": name, "ms": ms, "tflops": flops / ms / 1e9}
    print(json.dumps(rec), flush=True)
    results.append(rec)


def bench_varlen(name, iters=100):
    """BASELINE config 3: 64 packed sequences, randint(1, 2049) seed 0, H=32, D=128, bf16 causal."""
    from flash_attn_v100 import flash_attn_varlen_func

    g = torch.Generator().manual_seed(0)
    lens = torch.randint(1, 2049, (64,), generator=g)
    H, D, T = 32, 128, int(lens.sum())
    torch.manual_seed(421)
    q = torch.randn(T, H, D, device="cuda", dtype=torch.bfloat16)
    k, v = torch.randn_like(q), torch.randn_like(q)
    cu = torch.nn.functional.pad(lens.cumsum(0), (1, 0)).int().cuda()
    mx = int(lens.max())
    for _ in range(3):
        flash_attn_varlen_func(q, k, v, cu, cu, mx, mx, causal=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        flash_attn_varlen_func(q, k, v, cu, cu, mx, mx, causal=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 4 * D * H * float((lens.double() * (lens.double() + 1) / 2).sum())
    rec = {"name": name, "ms": ms, "tflops": flops / ms / 1e9}
    print(json.dumps(rec), flush=True)
    results.append(rec)


if __name__ == "__main__":
    f16, bf16 = torch.float16, torch.bfloat16
    t0 = time.time()
    if os.environ.get("QUICK_BENCH_ONLY"):
        run_case = lambda *a, **k: None  # noqa: E731
    run_case("d128_bf16_full_256", 1, 256, 256, 1, 1, 128, bf16, False)
    run_case("d128_bf16_full_128x384", 1, 128, 384, 2, 2, 128, bf16, False)
    run_case("d128_bf16_causal_512", 2, 512, 512, 4, 4, 128, bf16, True)
    run_case("d64_f16_full_512_C1", 2, 512, 512, 8, 8, 64, f16, False)
    run_case("d64_bf16_causal_1024", 1, 1024, 1024, 4, 2, 64, bf16, True)
    run_case("d128_f16_causal_ragged", 2, 333, 777, 4, 2, 128, f16, True)
    run_case("d128_bf16_window", 1, 1024, 1024, 2, 2, 128, bf16, True, window=(256, 0))
    run_case("d128_bf16_softcap", 1, 512, 512, 2, 2, 128, bf16, False, softcap=30.0)
    run_case("d128_bf16_alibi", 1, 512, 512, 2, 2, 128, bf16, True, alibi=True)
    run_case("d128_bf16_causal_2048_gqa", 1, 2048, 2048, 8, 2, 128, bf16, True)
    tag = sys.argv[1] if len(sys.argv) > 1 else "default"
    if all(r.get("ok") for r in results) and not os.environ.get("QUICK_PARITY_ONLY"):
        bench_case("C2_bf16_B8_H32_S4096_D128_causal", 8, 4096, 32, 32, 128, bf16, True, iters=30)
        bench_case("bf16_B32_H32_S1024_D128_causal", 32, 1024, 32, 32, 128, bf16, True, iters=200)
        bench_case("bf16_B8_H32_S4096_D128_full", 8, 4096, 32, 32, 128, bf16, False)
        bench_case("C2gqa_bf16_B8_H32_Hk8_S4096_causal", 8, 4096, 32, 8, 128, bf16, True, iters=20)
        bench_case("C5shard_bf16_B8_H32_S8192_win4096", 8, 8192, 32, 32, 128, bf16, True, window=(4096, 0))
        bench_case("bf16_B4_H32_S16384_D128_causal", 4, 16384, 32, 32, 128, bf16, True, iters=5)
        bench_case("f16_B8_H32_S4096_D64_causal", 8, 4096, 32, 32, 64, f16, True)
        bench_case("C1_f16_B2_H8_S512_D64_full", 2, 512, 8, 8, 64, f16, False)
        bench_varlen("C3_varlen_64seqs_H32_D128_causal")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", f"quick_{tag}.json"), "w"), indent=1)
    print("elapsed", time.time() - t0)
