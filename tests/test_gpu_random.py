"""Randomised GPU parity sweep: many small shape / feature combinations against the CPU oracle.

Seeds are fixed, so failures reproduce; sizes are small enough for the float64 oracle to finish in milliseconds.
Exercises the persistent scheduler with many work items of different lengths in one launch (ragged varlen
batches, skipped blocks, empty key ranges) -- the situations the hand-picked cases may miss.
"""
import os
import random

import pytest
import torch

from oracle import attention_oracle as ao

pytestmark = pytest.mark.gpu

# FA_FUZZ_MULT=k runs k times as many seeds per sweep (a longer fuzzing session on the GPU box; default 1)
_MULT = int(os.environ.get("FA_FUZZ_MULT", "1"))


@pytest.fixture(scope="module")
def api(fa_lib):
    import flash_attn_v100 as m

    return m


def _tol(dtype):
    return 2e-2 if dtype == torch.bfloat16 else 4e-3


@pytest.mark.parametrize("seed", range(40 * _MULT))
def test_random_dense(api, seed):
    rng = random.Random(seed)
    dtype = rng.choice([torch.float16, torch.bfloat16])
    D = rng.choice([64, 128, 32, 96, 256, 16, 192])
    Hk = rng.choice([1, 2, 3])
    H = Hk * rng.choice([1, 2, 4])
    B = rng.randint(1, 3)
    Sq = rng.choice([1, 7, 64, 128, 129, 255, 256, 257, 300, 513])
    Sk = rng.choice([1, 5, 127, 128, 129, 256, 384, 511, 700])
    causal = rng.random() < 0.5
    window = rng.choice([(-1, -1), (-1, -1), (rng.randint(0, 300), rng.randint(0, 100)), (rng.randint(0, 200), -1)])
    softcap = rng.choice([0.0, 0.0, 0.0, 25.0])
    use_alibi = rng.random() < 0.25
    torch.manual_seed(seed)
    q = torch.randn(B, Sq, H, D, device="cuda", dtype=dtype)
    k = torch.randn(B, Sk, Hk, D, device="cuda", dtype=dtype)
    v = torch.randn(B, Sk, Hk, D, device="cuda", dtype=dtype)
    slopes = (torch.rand(H, device="cuda") * 0.2).float() if use_alibi else None
    out = api.flash_attn_func(q, k, v, causal=causal, window_size=window, softcap=softcap, alibi_slopes=slopes)
    ref, _ = ao.flash_attn_func_ref(q, k, v, causal=causal, window_size=window, softcap=softcap, alibi_slopes=slopes)
    err = (out.double().cpu() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all() and err <= _tol(dtype), (seed, err)


@pytest.mark.parametrize("seed", range(16 * _MULT))
def test_random_varlen(api, seed):
    rng = random.Random(1000 + seed)
    dtype = rng.choice([torch.float16, torch.bfloat16])
    D = rng.choice([64, 128, 256, 40])
    Hk = rng.choice([1, 2])
    H = Hk * rng.choice([1, 4])
    nseq = rng.randint(1, 9)
    lq = [rng.choice([1, 3, 64, 128, 200, 257, 600]) for _ in range(nseq)]
    same = rng.random() < 0.5
    lk = lq if same else [rng.choice([1, 17, 128, 129, 400, 900]) for _ in range(nseq)]
    causal = rng.random() < 0.6
    torch.manual_seed(seed)
    q = torch.randn(sum(lq), H, D, device="cuda", dtype=dtype)
    k = torch.randn(sum(lk), Hk, D, device="cuda", dtype=dtype)
    v = torch.randn(sum(lk), Hk, D, device="cuda", dtype=dtype)
    cq = torch.tensor([0] + list(torch.tensor(lq).cumsum(0)), dtype=torch.int32, device="cuda")
    ck = torch.tensor([0] + list(torch.tensor(lk).cumsum(0)), dtype=torch.int32, device="cuda")
    out = api.flash_attn_varlen_func(q, k, v, cq, ck, max(lq), max(lk), causal=causal)
    ref, _ = ao.flash_attn_varlen_func_ref(q, k, v, cq, ck, max(lq), max(lk), causal=causal)
    err = (out.double().cpu() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all() and err <= _tol(dtype), (seed, err)


@pytest.mark.parametrize("seed", range(14 * _MULT))
def test_random_kvcache(api, seed):
    rng = random.Random(2000 + seed)
    dtype = rng.choice([torch.float16, torch.bfloat16])
    D = rng.choice([64, 128, 256, 32])
    Hk = rng.choice([1, 2, 4])
    H = Hk * rng.choice([1, 2, 8])
    B = rng.randint(1, 4)
    Sq = rng.choice([1, 1, 2, 4, 16, 140])
    cap = 768
    paged = rng.random() < 0.5
    causal = rng.random() < 0.7
    lens = torch.tensor([rng.randint(0, cap - Sq) for _ in range(B)], dtype=torch.int32, device="cuda")
    torch.manual_seed(seed)
    q = torch.randn(B, Sq, H, D, device="cuda", dtype=dtype)
    kn = torch.randn(B, Sq, Hk, D, device="cuda", dtype=dtype)
    vn = torch.randn(B, Sq, Hk, D, device="cuda", dtype=dtype)
    if paged:
        npg = B * (cap // 256)
        kc = torch.randn(npg, 256, Hk, D, device="cuda", dtype=dtype)
        vc = torch.randn(npg, 256, Hk, D, device="cuda", dtype=dtype)
        bt = torch.randperm(npg, generator=torch.Generator().manual_seed(seed)).view(B, -1).int().cuda()
    else:
        kc = torch.randn(B, cap, Hk, D, device="cuda", dtype=dtype)
        vc = torch.randn(B, cap, Hk, D, device="cuda", dtype=dtype)
        bt = None
    ref, lse_ref, kc_ref, vc_ref = ao.flash_attn_with_kvcache_ref(q, kc, vc, kn, vn, cache_seqlens=lens, block_table=bt, causal=causal)
    out, lse = api.flash_attn_with_kvcache(q, kc, vc, kn, vn, cache_seqlens=lens, block_table=bt, causal=causal,
                                           return_softmax_lse=True)
    assert torch.equal(kc.cpu(), kc_ref) and torch.equal(vc.cpu(), vc_ref)  # no rotary: the append is bit-exact
    err = (out.double().cpu() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all() and err <= _tol(dtype), (seed, err)
    assert (lse.double().cpu() - lse_ref).abs().max().item() < 3e-3


@pytest.mark.parametrize("seed", range(24 * _MULT))
def test_random_backward(api, seed):
    """dQ, dK, dV through autograd against the float64 oracle: head dims of all three tile widths, GQA, ragged
    lengths, masks, softcap, dropout (the oracle replays the forward's Philox stream from the generator state)."""
    rng = random.Random(3000 + seed)
    dtype = rng.choice([torch.float16, torch.bfloat16])
    D = rng.choice([64, 128, 256, 32, 96, 192])
    Hk = rng.choice([1, 2])
    H = Hk * rng.choice([1, 2, 4])
    B = rng.randint(1, 2)
    Sq = rng.choice([1, 64, 128, 129, 200, 256, 300, 513])
    Sk = rng.choice([8, 128, 136, 256, 384, 520])
    kw = {}
    if rng.random() < 0.5:
        kw["causal"] = True
    elif rng.random() < 0.4:
        kw["window_size"] = (rng.randint(0, 200), rng.randint(0, 100))
    p_drop = rng.choice([0.0, 0.0, 0.2])
    if p_drop == 0.0 and rng.random() < 0.25:
        kw["softcap"] = 20.0
    torch.manual_seed(seed)
    q = torch.randn(B, Sq, H, D, device="cuda", dtype=dtype, requires_grad=True)
    k = torch.randn(B, Sk, Hk, D, device="cuda", dtype=dtype, requires_grad=True)
    v = torch.randn(B, Sk, Hk, D, device="cuda", dtype=dtype, requires_grad=True)
    do = torch.randn(B, Sq, H, D, device="cuda", dtype=dtype)
    okw = dict(kw)
    if p_drop > 0.0:
        gen = torch.cuda.default_generators[0]
        gen.manual_seed(500 + seed)
        okw["rng_state"] = (gen.initial_seed(), gen.get_offset())
        kw["dropout_p"] = okw["dropout_p"] = p_drop
    out = api.flash_attn_func(q, k, v, **kw)
    dq, dk, dv = torch.autograd.grad(out, (q, k, v), do)
    rq, rk, rv, _ = ao.flash_attn_bwd_ref(do, q, k, v, **okw)
    tol = (4e-2 if dtype == torch.bfloat16 else 6e-3) / (1.0 - p_drop)
    for name, got, ref in (("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv)):
        err = (got.double().cpu() - ref).abs().max().item()
        assert torch.isfinite(got.float()).all() and err <= tol * max(1.0, ref.abs().max().item()), (seed, name, err)


@pytest.mark.parametrize("seed", range(12 * _MULT))
def test_random_varlen_backward(api, seed):
    rng = random.Random(4000 + seed)
    dtype = rng.choice([torch.float16, torch.bfloat16])
    D = rng.choice([64, 128, 256, 48])
    Hk = rng.choice([1, 2])
    H = Hk * rng.choice([1, 2])
    nseq = rng.randint(1, 6)
    lq = [rng.choice([1, 5, 64, 128, 130, 257, 400]) for _ in range(nseq)]
    lk = lq if rng.random() < 0.6 else [rng.choice([8, 24, 128, 136, 392]) for _ in range(nseq)]
    kw = {}
    if rng.random() < 0.6:
        kw["causal"] = True
    elif rng.random() < 0.5:
        kw["window_size"] = (rng.randint(0, 150), rng.randint(0, 60))
    use_alibi = rng.random() < 0.25
    torch.manual_seed(seed)
    q = torch.randn(sum(lq), H, D, device="cuda", dtype=dtype, requires_grad=True)
    k = torch.randn(sum(lk), Hk, D, device="cuda", dtype=dtype, requires_grad=True)
    v = torch.randn(sum(lk), Hk, D, device="cuda", dtype=dtype, requires_grad=True)
    do = torch.randn(sum(lq), H, D, device="cuda", dtype=dtype)
    if use_alibi:
        kw["alibi_slopes"] = (torch.rand(H, device="cuda") * 0.2).float()
    cq = torch.tensor([0] + list(torch.tensor(lq).cumsum(0)), dtype=torch.int32, device="cuda")
    ck = torch.tensor([0] + list(torch.tensor(lk).cumsum(0)), dtype=torch.int32, device="cuda")
    out = api.flash_attn_varlen_func(q, k, v, cq, ck, max(lq), max(lk), **kw)
    dq, dk, dv = torch.autograd.grad(out, (q, k, v), do)
    rq, rk, rv, _ = ao.flash_attn_varlen_bwd_ref(do, q, k, v, cq, ck, max(lq), max(lk), **kw)
    tol = 4e-2 if dtype == torch.bfloat16 else 6e-3
    for name, got, ref in (("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv)):
        err = (got.double().cpu() - ref).abs().max().item()
        assert torch.isfinite(got.float()).all() and err <= tol * max(1.0, ref.abs().max().item()), (seed, name, err)


@pytest.mark.parametrize("seed", range(6 * _MULT))
def test_random_dense_many_items_per_cta(api, seed):
    """More work items than SMs, of very different lengths: every persistent CTA walks several items, so the
    cross-item pipeline (next item's loads under this item's epilogue, barrier phases carried across items) is
    exercised the way the full-size configs do, but at a size the float64 oracle still checks element by element."""
    rng = random.Random(5000 + seed)
    dtype = rng.choice([torch.float16, torch.bfloat16])
    D = rng.choice([64, 128, 128, 256])
    B, H, Hk = 2, 12, rng.choice([2, 12])
    Sq = rng.randint(1500, 2600)
    Sk = Sq if rng.random() < 0.6 else rng.randint(1500, 2600)
    causal = rng.random() < 0.7
    window = (-1, -1) if rng.random() < 0.6 else (rng.randint(100, 900), rng.randint(0, 300) if not causal else 0)
    torch.manual_seed(seed)
    q = torch.randn(B, Sq, H, D, device="cuda", dtype=dtype)
    k = torch.randn(B, Sk, Hk, D, device="cuda", dtype=dtype)
    v = torch.randn(B, Sk, Hk, D, device="cuda", dtype=dtype)
    out = api.flash_attn_func(q, k, v, causal=causal, window_size=window)
    ref, _ = ao.flash_attn_func_ref(q, k, v, causal=causal, window_size=window)
    err = (out.double().cpu() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all() and err <= _tol(dtype), (seed, err)
