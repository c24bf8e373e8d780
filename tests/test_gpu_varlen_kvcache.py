"""GPU parity: packed variable-length and KV-cache forwards vs the CPU oracle.

The reference never tests these paths ("parity unpinned"); the oracle restates the
semantics of kernel/fused_mha_forward_varlen.cu and kernel/fused_mha_forward_kvcache.cu + include/rotary.h.
"""
import math

import pytest
import torch

from oracle import attention_oracle as ao

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(fa_lib):
    import flash_attn_v100 as m

    return m


@pytest.fixture(scope="module")
def op(fa_lib):
    import flash_attn_v100_cuda as m

    return m


def _tol(dtype):
    return 2e-2 if dtype == torch.bfloat16 else 3e-3


def _cu(lens, device="cuda"):
    return torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=device)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("lens", [[5, 33, 1, 64, 300, 257, 128], [1], [700, 3]])
def test_varlen_vs_oracle(api, op, dtype, causal, lens):
    _varlen_case(api, op, dtype, causal, lens, 128)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("D", [32, 256])
def test_varlen_head_dims(api, op, causal, D):
    _varlen_case(api, op, torch.bfloat16, causal, [5, 33, 1, 64, 300, 257, 128], D)


def _varlen_case(api, op, dtype, causal, lens, D):
    torch.manual_seed(421)
    H, Hk = 4, 2
    T = sum(lens)
    q = torch.randn(T, H, D, device="cuda", dtype=dtype)
    k = torch.randn(T, Hk, D, device="cuda", dtype=dtype)
    v = torch.randn(T, Hk, D, device="cuda", dtype=dtype)
    cu = _cu(lens)
    out = api.flash_attn_varlen_func(q, k, v, cu, cu, max(lens), max(lens), causal=causal)
    assert out.shape == q.shape
    ref, lse_ref = ao.flash_attn_varlen_func_ref(q, k, v, cu, cu, max(lens), max(lens), causal=causal)
    assert (out.double().cpu() - ref).abs().max().item() <= _tol(dtype)
    o2, lse, _, _ = op.varlen_fwd(q, k, v, None, cu, cu, None, None, None, None, max(lens), max(lens), 0.0,
                                  D ** -0.5, False, causal, -1, -1, 0.0, False, None, 0)
    assert lse.shape == (H, T)  # reference kernel/fused_mha_forward_varlen.cu:519
    assert (lse.double().cpu() - lse_ref).abs().max().item() < 2e-3
    assert torch.equal(o2, out)


def test_varlen_different_q_and_k_lengths_window_seqused(op):
    torch.manual_seed(1)
    H, Hk, D = 4, 4, 64
    lq, lk = [10, 200, 77], [300, 200, 5]
    q = torch.randn(sum(lq), H, D, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(sum(lk), Hk, D, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(sum(lk), Hk, D, device="cuda", dtype=torch.bfloat16)
    cq, ck = _cu(lq), _cu(lk)
    used = torch.tensor([250, 0, 5], dtype=torch.int32, device="cuda")
    out, lse, _, _ = op.varlen_fwd(q, k, v, None, cq, ck, used, None, None, None, max(lq), max(lk), 0.0, D ** -0.5,
                                   False, True, 50, -1, 0.0, False, None, 0)
    ref, lse_ref = ao.flash_attn_varlen_func_ref(q, k, v, cq, ck, max(lq), max(lk), causal=True, window_size=(50, -1),
                                                 seqused_k=used.cpu())
    assert (out.double().cpu() - ref).abs().max().item() <= 2e-2
    # sequence 1 has seqused_k == 0: no keys => out 0, lse sentinel (reference ..._varlen.cu:100-111)
    assert out[10:210].abs().max().item() == 0
    assert (lse[:, 10:210] <= -1e29).all()


def test_varlen_paged_kv(op):
    torch.manual_seed(2)
    H, Hk, D, page = 4, 2, 128, 256
    lq, lk = [3, 130, 64], [600, 130, 1000]
    n_pages = 12
    kc = torch.randn(n_pages, page, Hk, D, device="cuda", dtype=torch.bfloat16)
    vc = torch.randn(n_pages, page, Hk, D, device="cuda", dtype=torch.bfloat16)
    bt = torch.tensor([[3, 1, 9, 0], [7, 0, 0, 0], [2, 5, 11, 4]], dtype=torch.int32, device="cuda")
    q = torch.randn(sum(lq), H, D, device="cuda", dtype=torch.bfloat16)
    cq, ck = _cu(lq), _cu(lk)
    out, lse, _, _ = op.varlen_fwd(q, kc, vc, None, cq, ck, None, None, bt, None, max(lq), max(lk), 0.0, D ** -0.5,
                                   False, True, -1, -1, 0.0, False, None, 0)
    ref, _ = ao.flash_attn_varlen_func_ref(q, kc, vc, cq, ck, max(lq), max(lk), causal=True, block_table=bt)
    assert (out.double().cpu() - ref).abs().max().item() <= 2e-2


def test_config3_scaled_and_full_sampled(api):
    """BASELINE config 3: 64 packed sequences, randint(1,2049) seed 0, H=32, D=128, causal.
    Full size on the GPU; the CPU oracle recomputes a sample of whole sequences."""
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(1, 2049, (64,), generator=g).tolist()
    H, D = 32, 128
    T = sum(lens)
    torch.manual_seed(421)
    q = torch.randn(T, H, D, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(T, H, D, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(T, H, D, device="cuda", dtype=torch.bfloat16)
    cu = _cu(lens)
    out = api.flash_attn_varlen_func(q, k, v, cu, cu, max(lens), max(lens), causal=True)
    assert torch.isfinite(out.float()).all()
    order = sorted(range(64), key=lambda b: lens[b])
    for b in (order[0], order[1], order[20], order[-1]):  # shortest ... longest
        s, e = int(cu[b]), int(cu[b + 1])
        hs = slice(0, 32, 11)
        ref, _ = ao.flash_attn_func_ref(q[s:e, hs][None], k[s:e, hs][None], v[s:e, hs][None], causal=True)
        assert (out[s:e, hs].double().cpu() - ref[0]).abs().max().item() <= 2e-2, b
    # packing invariance: a sequence computed alone is bit-identical to its slice of the packed call
    b = order[30]
    s, e = int(cu[b]), int(cu[b + 1])
    alone = api.flash_attn_func(q[s:e][None], k[s:e][None], v[s:e][None], causal=True)
    assert torch.equal(alone[0], out[s:e])


# ------------------------------------------------------------------------------------------ kv-cache
def _rotary_tables(seqlen, rot, dtype, device="cuda"):
    inv = 1.0 / (10000 ** (torch.arange(0, rot, 2, dtype=torch.float32) / rot))
    ang = torch.outer(torch.arange(seqlen, dtype=torch.float32), inv)
    return ang.cos().to(dtype).to(device), ang.sin().to(dtype).to(device)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("paged", [False, True])
@pytest.mark.parametrize("Sq,interleaved,causal", [(1, True, True), (1, False, False), (5, False, True), (130, True, True)])
def test_kvcache_append_rotary_vs_oracle(api, dtype, paged, Sq, interleaved, causal):
    _kvcache_append_rotary_case(api, dtype, paged, Sq, interleaved, causal, 128)


@pytest.mark.parametrize("D", [32, 96, 256])
@pytest.mark.parametrize("Sq,paged", [(1, True), (5, False), (130, True)])
def test_kvcache_head_dims_without_padding(api, D, Sq, paged):
    """Head dims below the kernel's tile width go through unpadded (TMA zero-fills the tile's extra columns),
    so the cache is appended to and read in place."""
    _kvcache_append_rotary_case(api, torch.bfloat16, paged, Sq, False, True, D)


def _kvcache_append_rotary_case(api, dtype, paged, Sq, interleaved, causal, D):
    torch.manual_seed(3)
    B, H, Hk = 3, 8, 2
    cap, page = 1024, 256
    lens = torch.tensor([500, 37, 1024 - Sq], dtype=torch.int32, device="cuda")
    if paged:
        n_pages = B * (cap // page) + 2
        kc = torch.randn(n_pages, page, Hk, D, device="cuda", dtype=dtype)
        vc = torch.randn(n_pages, page, Hk, D, device="cuda", dtype=dtype)
        bt = torch.randperm(n_pages, generator=torch.Generator().manual_seed(0))[: B * (cap // page)].view(B, -1).int().cuda()
    else:
        kc = torch.randn(B, cap, Hk, D, device="cuda", dtype=dtype)
        vc = torch.randn(B, cap, Hk, D, device="cuda", dtype=dtype)
        bt = None
    q = torch.randn(B, Sq, H, D, device="cuda", dtype=dtype)
    kn = torch.randn(B, Sq, Hk, D, device="cuda", dtype=dtype)
    vn = torch.randn(B, Sq, Hk, D, device="cuda", dtype=dtype)
    cos, sin = _rotary_tables(cap, D, dtype)
    ref, lse_ref, kc_ref, vc_ref = ao.flash_attn_with_kvcache_ref(
        q, kc, vc, kn, vn, cos, sin, lens, block_table=bt, causal=causal, rotary_interleaved=interleaved)
    out, lse = api.flash_attn_with_kvcache(q, kc, vc, kn, vn, rotary_cos=cos, rotary_sin=sin, cache_seqlens=lens,
                                           block_table=bt, causal=causal, rotary_interleaved=interleaved,
                                           return_softmax_lse=True)
    # the cache is mutated in place (reference README.md:151); V bit-exact, K within 1 ulp on rare fma ties
    assert torch.equal(vc.cpu(), vc_ref)
    kd = (kc.float().cpu() - kc_ref.float()).abs()
    assert (kd > 0).float().mean().item() < 1e-4 and kd.max().item() <= 0.04
    assert (out.double().cpu() - ref).abs().max().item() <= _tol(dtype)
    assert lse.shape == (B, H, Sq)
    assert (lse.double().cpu() - lse_ref).abs().max().item() < 2e-3


def test_kvcache_batch_idx_leftpad_no_append(api):
    torch.manual_seed(4)
    B, H, Hk, D, cap = 2, 4, 4, 64, 512
    kc = torch.randn(5, cap, Hk, D, device="cuda", dtype=torch.float16)
    vc = torch.randn(5, cap, Hk, D, device="cuda", dtype=torch.float16)
    q = torch.randn(B, 3, H, D, device="cuda", dtype=torch.float16)
    lens = torch.tensor([100, 257], dtype=torch.int32, device="cuda")
    idx = torch.tensor([4, 1], dtype=torch.int32, device="cuda")
    pad = torch.tensor([8, 40], dtype=torch.int32, device="cuda")
    out = api.flash_attn_with_kvcache(q, kc, vc, cache_seqlens=lens, cache_batch_idx=idx, cache_leftpad=pad, causal=True)
    ref, _, _, _ = ao.flash_attn_with_kvcache_ref(q, kc, vc, cache_seqlens=lens, cache_batch_idx=idx, cache_leftpad=pad, causal=True)
    assert (out.double().cpu() - ref).abs().max().item() <= 3e-3
    # int cache_seqlens and window
    out = api.flash_attn_with_kvcache(q[:, :1], kc[:2], vc[:2], cache_seqlens=300, window_size=(64, 0))
    ref, _, _, _ = ao.flash_attn_with_kvcache_ref(q[:, :1], kc[:2], vc[:2], cache_seqlens=300, window_size=(64, 0))
    assert (out.double().cpu() - ref).abs().max().item() <= 3e-3


def test_config4_decode_paged_rotary(api):
    """BASELINE config 4: decode Sq=1, Sk=8192 (8191 cached + 1 appended), H=32, Hk=8, D=128, rotary,
    paged page=256, block_table = random permutation. B=8 here so the CPU oracle is quick."""
    torch.manual_seed(421)
    B, H, Hk, D, page, Sk = 8, 32, 8, 128, 256, 8192
    n_pages = B * (Sk // page)
    kc = torch.randn(n_pages, page, Hk, D, device="cuda", dtype=torch.bfloat16)
    vc = torch.randn(n_pages, page, Hk, D, device="cuda", dtype=torch.bfloat16)
    bt = torch.randperm(n_pages, generator=torch.Generator().manual_seed(0)).view(B, -1).int().cuda()
    lens = torch.full((B,), Sk - 1, dtype=torch.int32, device="cuda")
    q = torch.randn(B, 1, H, D, device="cuda", dtype=torch.bfloat16)
    kn = torch.randn(B, 1, Hk, D, device="cuda", dtype=torch.bfloat16)
    vn = torch.randn(B, 1, Hk, D, device="cuda", dtype=torch.bfloat16)
    cos, sin = _rotary_tables(Sk, D, torch.bfloat16)
    for interleaved in (False, True):
        kc2, vc2 = kc.clone(), vc.clone()
        ref, lse_ref, _, _ = ao.flash_attn_with_kvcache_ref(q, kc2, vc2, kn, vn, cos, sin, lens, block_table=bt,
                                                            causal=True, rotary_interleaved=interleaved)
        out, lse = api.flash_attn_with_kvcache(q, kc2, vc2, kn, vn, rotary_cos=cos, rotary_sin=sin, cache_seqlens=lens,
                                               block_table=bt, causal=True, rotary_interleaved=interleaved,
                                               return_softmax_lse=True)
        assert (out.double().cpu() - ref).abs().max().item() <= 2e-2
        assert (lse.double().cpu() - lse_ref).abs().max().item() < 2e-3


def test_kvcache_argument_errors(op):
    q = torch.zeros(2, 1, 4, 64, device="cuda", dtype=torch.float16)
    kc = torch.zeros(2, 256, 4, 64, device="cuda", dtype=torch.float16)
    lens = torch.zeros(2, dtype=torch.int32, device="cuda")
    with pytest.raises(RuntimeError, match="num_splits > 1"):
        op.fwd_kvcache(q, kc, kc, None, None, lens, None, None, None, None, None, None, None, 1.0, False, -1, -1, 0.0, True, 2)
    with pytest.raises(RuntimeError, match="seqlens_k is required"):
        op.fwd_kvcache(q, kc, kc, q, q, None, None, None, None, None, None, None, None, 1.0, False, -1, -1, 0.0, True, 0)
    bt = torch.zeros(2, 1, dtype=torch.int32, device="cuda")
    with pytest.raises(RuntimeError, match="cache_batch_idx"):
        op.fwd_kvcache(q, kc, kc, None, None, lens, None, None, lens, None, bt, None, None, 1.0, False, -1, -1, 0.0, True, 0)
    with pytest.raises(RuntimeError, match="softcap does not support window"):
        op.fwd_kvcache(q, kc, kc, None, None, lens, None, None, None, None, None, None, None, 1.0, False, 3, -1, 5.0, True, 0)


@pytest.mark.parametrize("Sq", [1, 200])
@pytest.mark.parametrize("paged", [False, True])
def test_kvcache_rows_past_the_valid_length_may_hold_nan(api, Sq, paged):
    """A cache is allowed to contain uninitialised memory beyond cache_seqlens: P is 0 there, but 0 * NaN
    would poison P V unless the ragged V tile is sanitised (csrc/fwd_sm100.cuh, warp 14)."""
    torch.manual_seed(5)
    B, H, Hk, D, cap, page = 2, 8, 2, 128, 1024, 256
    lens = torch.tensor([300, 777], dtype=torch.int32, device="cuda")
    q = torch.randn(B, Sq, H, D, device="cuda", dtype=torch.bfloat16)
    if paged:
        kc = torch.randn(B * cap // page, page, Hk, D, device="cuda", dtype=torch.bfloat16)
        vc = torch.randn_like(kc)
        bt = torch.arange(B * cap // page, dtype=torch.int32, device="cuda").view(B, -1).flip(1).contiguous()
        flat_k, flat_v = kc.view(-1, Hk, D), vc.view(-1, Hk, D)
        for b in range(B):
            for pg in range(cap // page):
                lo = max(int(lens[b]) - pg * page, 0)
                base = int(bt[b, pg]) * page
                flat_k[base + lo: base + page] = float("nan")
                flat_v[base + lo: base + page] = float("nan")
    else:
        kc = torch.randn(B, cap, Hk, D, device="cuda", dtype=torch.bfloat16)
        vc = torch.randn_like(kc)
        bt = None
        for b in range(B):
            kc[b, int(lens[b]):] = float("nan")
            vc[b, int(lens[b]):] = float("nan")
    out = api.flash_attn_with_kvcache(q, kc, vc, cache_seqlens=lens, block_table=bt, causal=True)
    assert torch.isfinite(out.float()).all()
    kz, vz = torch.nan_to_num(kc, nan=0.0), torch.nan_to_num(vc, nan=0.0)
    ref, _, _, _ = ao.flash_attn_with_kvcache_ref(q, kz, vz, cache_seqlens=lens, block_table=bt, causal=True)
    assert (out.double().cpu() - ref).abs().max().item() <= 2e-2
