"""GPU parity: dropout in the dense and var-len forward vs the CPU oracle.

The reference's scheme (include/softmax.h:96-125, include/philox.h, kernel/fused_mha_forward.cu:373-406):
Philox4x32-10 keyed by the generator seed, counter = offset + (flat index >> 2), flat index = row * N + col
without a batch/head term; kept P scaled by 1/(1-p); normaliser and LSE from P before dropout; dmask holds
+1 (kept) / -1 (dropped); rng_state = [seed, offset]; the generator offset advances by B*H*32. The keep
mask is an integer function, so it is compared bit-exactly; `out` within the bf16/fp16 bound.
"""
import pytest
import torch

from oracle import attention_oracle as ao

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(fa_lib):
    import flash_attn_v100 as m

    return m


def _visible(Sq, Sk, causal, window):
    i = torch.arange(Sq).view(Sq, 1)
    j = torch.arange(Sk).view(1, Sk)
    off = Sk - Sq
    vis = torch.ones(Sq, Sk, dtype=torch.bool)
    wl, wr = window
    if causal:
        wr = 0
    if wr >= 0:
        vis &= j <= i + off + wr
    if wl >= 0:
        vis &= j >= i + off - wl
    return vis


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,Sq,Sk,H,Hk,D,causal,window,p", [
    (2, 256, 256, 4, 4, 128, False, (-1, -1), 0.1),
    (1, 384, 384, 4, 2, 128, True, (-1, -1), 0.25),
    (2, 200, 328, 2, 2, 64, True, (-1, -1), 0.5),      # Sq != Sk, ragged tiles, Sk % 8 == 0
    (1, 130, 203, 2, 1, 64, False, (-1, -1), 0.3),     # Sk % 4 != 0: unaligned Philox words and dmask rows
    (1, 512, 512, 2, 2, 128, False, (100, 30), 0.2),   # sliding window
    (1, 96, 96, 2, 2, 32, True, (-1, -1), 0.15),       # head dim below the tile width
    (1, 300, 300, 2, 2, 256, True, (-1, -1), 0.2),     # head dim 256: both D-halves must see the same mask
])
def test_dense_dropout_matches_oracle(api, dtype, B, Sq, Sk, H, Hk, D, causal, window, p):
    torch.manual_seed(421)
    q = torch.randn(B, Sq, H, D, device="cuda", dtype=dtype)
    k = torch.randn(B, Sk, Hk, D, device="cuda", dtype=dtype)
    v = torch.randn(B, Sk, Hk, D, device="cuda", dtype=dtype)
    gen = torch.cuda.default_generators[0]
    gen.manual_seed(1234)
    gen.set_offset(4 * 77)
    seed, offset = gen.initial_seed(), gen.get_offset()
    out, lse, dmask = api.flash_attn_func(q, k, v, dropout_p=p, causal=causal, window_size=window,
                                          return_attn_probs=True)
    torch.cuda.synchronize()
    assert gen.get_offset() == offset + B * H * 32          # reference fused_mha_forward.cu:382-383
    assert dmask.shape == (B, H, Sq, Sk) and dmask.dtype == dtype

    keep = ao.dropout_keep_mask(p, seed, offset, 0, Sq, Sk, Sk)
    vis = _visible(Sq, Sk, causal, window)
    dm = dmask.float().cpu()
    assert bool(((dm == 1.0) | (dm == -1.0) | (dm == 0.0)).all())
    expect = torch.where(keep, 1.0, -1.0)
    for b in range(B):
        for h in range(H):
            assert bool((dm[b, h][vis] == expect[vis]).all()), (b, h)
    frac = keep[vis].float().mean().item()
    assert abs(frac - (1 - p)) < 0.02

    ref, lse_ref = ao.flash_attn_func_ref(q, k, v, causal=causal, window_size=window, dropout_p=p,
                                          rng_state=(seed, offset))
    err = (out.double().cpu() - ref).abs().max().item()
    assert err <= (2e-2 if dtype == torch.bfloat16 else 4e-3) / (1 - p), err
    # LSE is the no-dropout LSE
    out0, lse0, _ = api.flash_attn_func(q, k, v, causal=causal, window_size=window), None, None
    lerr = (lse.double().cpu() - lse_ref).abs()
    ok = (lerr <= 1e-3 * lse_ref.abs().clamp(min=1.0)) | (lse_ref <= -1e29)
    assert bool(ok.all())
    # and the output really differs from the no-dropout output
    assert (out.float() - out0.float()).abs().max().item() > 1e-2


def test_dense_dropout_rng_state_and_reproducibility(fa_lib):
    import flash_attn_v100_cuda as op

    torch.manual_seed(0)
    q = torch.randn(1, 2, 256, 64, device="cuda", dtype=torch.bfloat16)  # raw layout [B,H,M,D]
    gen = torch.Generator(device="cuda")
    gen.manual_seed(99)
    o1, l1, m1, r1 = op.fwd(q, q, q, None, None, 0.2, 0.125, True, -1, -1, 0.0, True, gen)
    assert r1.tolist() == [99, 0] and gen.get_offset() == 1 * 2 * 32
    o2, l2, m2, r2 = op.fwd(q, q, q, None, None, 0.2, 0.125, True, -1, -1, 0.0, True, gen)
    assert r2.tolist() == [99, 64]
    assert not torch.equal(m1, m2)
    gen.manual_seed(99)
    o3, _, m3, _ = op.fwd(q, q, q, None, None, 0.2, 0.125, True, -1, -1, 0.0, True, gen)
    assert torch.equal(o1, o3) and torch.equal(m1, m3)
    # without return_softmax no mask is materialised
    gen.manual_seed(99)
    o4, _, m4, _ = op.fwd(q, q, q, None, None, 0.2, 0.125, True, -1, -1, 0.0, False, gen)
    assert m4.numel() == 0 and torch.equal(o1, o4)
    with pytest.raises(RuntimeError, match="Softcapping does not support dropout"):
        op.fwd(q, q, q, None, None, 0.2, 0.125, True, -1, -1, 30.0, False, gen)
    with pytest.raises(RuntimeError, match=r"p_dropout must be in \[0, 1\)"):
        op.fwd(q, q, q, None, None, 1.0, 0.125, True, -1, -1, 0.0, False, gen)


def test_dense_dropout_with_alibi(api):
    torch.manual_seed(421)
    q = torch.randn(1, 256, 4, 64, device="cuda", dtype=torch.float16)
    k = torch.randn(1, 256, 4, 64, device="cuda", dtype=torch.float16)
    v = torch.randn(1, 256, 4, 64, device="cuda", dtype=torch.float16)
    slopes = torch.tensor([0.5, 0.25, 0.125, 0.0625], device="cuda", dtype=torch.float32)
    gen = torch.cuda.default_generators[0]
    gen.manual_seed(5)
    seed, offset = gen.initial_seed(), gen.get_offset()
    out = api.flash_attn_func(q, k, v, dropout_p=0.1, causal=True, alibi_slopes=slopes)
    ref, _ = ao.flash_attn_func_ref(q, k, v, causal=True, alibi_slopes=slopes, dropout_p=0.1, rng_state=(seed, offset))
    assert (out.double().cpu() - ref).abs().max().item() <= 5e-3


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("max_k_pad", [0, 3])
def test_varlen_dropout_matches_oracle(api, dtype, max_k_pad):
    torch.manual_seed(421)
    lens = [37, 256, 1, 300, 129]
    H, Hk, D, p = 4, 2, 128, 0.2
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    T = sum(lens)
    q = torch.randn(T, H, D, device="cuda", dtype=dtype)
    k = torch.randn(T, Hk, D, device="cuda", dtype=dtype)
    v = torch.randn(T, Hk, D, device="cuda", dtype=dtype)
    max_k = max(lens) + max_k_pad  # callers may pass a loose bound; it is the dropout row length
    gen = torch.cuda.default_generators[0]
    gen.manual_seed(4321)
    seed, offset = gen.initial_seed(), gen.get_offset()
    out, lse, dmask = api.flash_attn_varlen_func(q, k, v, cu, cu, max(lens), max_k, dropout_p=p, causal=True,
                                                 return_attn_probs=True)
    torch.cuda.synchronize()
    assert gen.get_offset() == offset + len(lens) * H * 32
    assert dmask.shape == (T, H, max_k) and lse.shape == (H, T)
    ref, lse_ref = ao.flash_attn_varlen_func_ref(q, k, v, cu, cu, max(lens), max_k, causal=True, dropout_p=p,
                                                 rng_state=(seed, offset))
    err = (out.double().cpu() - ref).abs().max().item()
    assert err <= (2e-2 if dtype == torch.bfloat16 else 4e-3) / (1 - p), err
    assert bool(((lse.double().cpu() - lse_ref).abs() <= 1e-3 * lse_ref.abs().clamp(min=1.0)).all())
    dm = dmask.float().cpu()
    s0 = 0
    for L in lens:
        keep = ao.dropout_keep_mask(p, seed, offset, s0, L, L, max_k)
        vis = _visible(L, L, True, (-1, -1))
        expect = torch.where(keep, 1.0, -1.0)
        for h in range(H):
            assert bool((dm[s0:s0 + L, h, :L][vis] == expect[vis]).all()), (s0, h)
        assert bool((dm[s0:s0 + L, :, L:] == 0).all())  # columns past the sequence are never written
        s0 += L


def test_full_size_dropout_statistics(api):
    """BASELINE config-2 geometry (one batch element): the kept fraction over the visible triangle is 1-p and
    E[out] is preserved (mean over many rows of out_dropout - out is ~0)."""
    torch.manual_seed(421)
    q = torch.randn(1, 4096, 4, 128, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(1, 4096, 4, 128, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(1, 4096, 4, 128, device="cuda", dtype=torch.bfloat16)
    out, lse, dmask = api.flash_attn_func(q, k, v, dropout_p=0.1, causal=True, return_attn_probs=True)
    out0 = api.flash_attn_func(q, k, v, causal=True)
    tri = torch.ones(4096, 4096, dtype=torch.bool, device="cuda").tril()
    kept = (dmask[0, 0][tri] == 1).float().mean().item()
    assert abs(kept - 0.9) < 2e-3
    assert torch.equal(dmask[0, 0], dmask[0, 3])  # the reference's index has no head term
    bias = (out.float() - out0.float())[:, 2048:].mean().item()
    assert abs(bias) < 2e-3
