"""Helpers shared by the GPU parity tests."""
import math

import torch

from oracle import attention_oracle as ao


def rand_qkv(B, Sq, Sk, H, Hk, D, dtype, seed=421, device="cuda"):
    torch.manual_seed(seed)  # the reference's seed, test.py:151
    q = torch.randn(B, Sq, H, D, device=device, dtype=dtype)
    k = torch.randn(B, Sk, Hk, D, device=device, dtype=dtype)
    v = torch.randn(B, Sk, Hk, D, device=device, dtype=dtype)
    return q, k, v


def check_dense(out, lse, q, k, v, causal=False, window=(-1, -1), softcap=0.0, alibi=None, scale=None,
                lse_tol=1e-3):
    """Reference pass rule (test.py:273-277) against the CPU oracle + LSE check (fp32 quantity)."""
    B, Sq, H, D = q.shape
    ref, lse_ref = ao.flash_attn_func_ref(q, k, v, softmax_scale=scale, causal=causal, window_size=window,
                                          softcap=softcap, alibi_slopes=alibi)
    wl, wr = ao.normalize_mask_args(Sq, k.shape[1], causal, window, alibi is not None)
    sc = D ** -0.5 if scale is None else scale
    if softcap == 0.0 and alibi is None:
        naive = ao.naive_lowp_attention(q, k, v, sc, wl, wr)
        ok, err, err_naive = ao.fa_tolerance_ok(out, ref, naive)
        assert ok, f"max err {err:.3e} > 2 * naive {err_naive:.3e} + 1e-5"
    else:
        # no same-precision torch yardstick for these modifiers: absolute bound for unit-variance inputs
        err = (out.double().cpu() - ref).abs().max().item()
        assert math.isfinite(err) and err <= (2e-2 if q.dtype == torch.bfloat16 else 3e-3), err
    if lse is not None:
        l = lse.double().cpu()
        bad = (l - lse_ref).abs() > lse_tol * torch.clamp(lse_ref.abs(), min=1.0)
        sentinel = lse_ref <= -1e29
        assert not bool((bad & ~sentinel).any()), (l - lse_ref).abs().max().item()
        assert bool((l[sentinel] <= -1e29).all())


def sampled_row_check(out, lse, q, k, v, rows, causal, window=(-1, -1), scale=None):
    """Full-size check: recompute `rows` = [(b, h, i)] exactly on the CPU (float64) and compare.
    Tolerance: 16-bit output rounding + P rounding, 2e-2 abs for bf16 / 3e-3 for fp16 on N(0,1) inputs."""
    B, Sq, H, D = q.shape
    Sk, Hk = k.shape[1], k.shape[2]
    sc = D ** -0.5 if scale is None else scale
    wl, wr = ao.normalize_mask_args(Sq, Sk, causal, window, False)
    tol = 2e-2 if q.dtype == torch.bfloat16 else 3e-3
    worst = 0.0
    for (b, h, i) in rows:
        hk = h // (H // Hk)
        qi = q[b, i, h].double().cpu()
        kk = k[b, :, hk].double().cpu()
        vv = v[b, :, hk].double().cpu()
        s = (kk @ qi) * sc
        j = torch.arange(Sk)
        off = Sk - Sq
        masked = torch.zeros(Sk, dtype=torch.bool)
        if wr >= 0:
            masked |= j > i + off + wr
        if wl >= 0:
            masked |= j < i + off - wl
        s = s.masked_fill(masked, float("-inf"))
        m = s.max()
        p = torch.exp(s - m)
        o = (p @ vv) / p.sum()
        err = (out[b, i, h].double().cpu() - o).abs().max().item()
        worst = max(worst, err)
        assert err <= tol, (b, h, i, err)
        if lse is not None:
            ref_l = (m + torch.log(p.sum())).item()
            assert abs(lse[b, h, i].item() - ref_l) <= 1e-3 * max(1.0, abs(ref_l)), (b, h, i)
    return worst
