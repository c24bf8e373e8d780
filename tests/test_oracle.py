"""Pins the CPU oracle (oracle/attention_oracle.py, oracle/attn_oracle.c) to the reference.

  * golden vectors produced by the reference's own CPU code (tests/golden/make_golden.py):
      - `cpu_attention` known-answer case, utils/sass/mma_swizzle/forward_kernel.cu:346-370, 394-407, 439
      - `ref_mha_forward`, test.py:18-34, on the seed-421 inputs of test.py:151-157
  * when oracle/_ref is present (dev container), the compiled reference function itself
  * torch SDPA on CPU for the subset SDPA expresses
"""
import ctypes
import ctypes.util
import glob
import os

import numpy as np
import pytest
import torch

from oracle import native
from oracle import attention_oracle as ao

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _bhmd_to_bshd(x):
    return torch.from_numpy(x).permute(0, 2, 1, 3).contiguous()


def test_kat_inputs_are_the_srand42_stream():
    """The fixture's inputs are exactly what the reference harness draws (srand(42), rand())."""
    g = np.load(os.path.join(GOLD, "ref_cpu_attention_kat.npz"))
    libc = ctypes.CDLL(ctypes.util.find_library("c"))
    libc.srand(42)
    q = ao.c_rand_uniform_pm1(128 * 128).reshape(128, 128)
    np.testing.assert_array_equal(q, g["q"])
    assert np.all(np.abs(g["q"]) <= 1.0)


def test_oracle_matches_reference_cpu_attention_golden():
    g = np.load(os.path.join(GOLD, "ref_cpu_attention_kat.npz"))
    q, k, v = (torch.from_numpy(g[n]).view(1, 128, 1, 128) for n in ("q", "k", "v"))
    out, lse = ao.flash_attn_func_ref(q, k, v, softmax_scale=float(g["scale"]), causal=bool(g["causal"]))
    err = np.abs(out.view(128, 128).numpy() - g["out"]).max()
    # reference harness tolerance is 5e-2 (forward_kernel.cu:433); fp32-vs-fp64 accumulation gives ~1e-6
    assert err < 2e-5, err
    assert lse.shape == (1, 1, 128)


def test_c_oracle_matches_reference_cpu_attention_golden():
    g = np.load(os.path.join(GOLD, "ref_cpu_attention_kat.npz"))
    q, k, v = (g[n].reshape(128, 1, 128) for n in ("q", "k", "v"))
    out, _ = native.c_oracle_attention(q, k, v, float(g["scale"]), wl=-1, wr=0)
    assert np.abs(out.reshape(128, 128) - g["out"]).max() < 2e-5


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "ref_mha_forward_*.npz"))))
def test_oracle_matches_reference_ref_mha_forward_golden(path):
    g = np.load(path)
    q, k, v = (_bhmd_to_bshd(g[n]) for n in ("q", "k", "v"))  # test.py layout is [B,H,M,D]
    out, _ = ao.flash_attn_func_ref(q, k, v, softmax_scale=float(g["scale"]), causal=bool(g["causal"]))
    ref = torch.from_numpy(g["out"]).permute(0, 2, 1, 3)
    assert (out.float() - ref).abs().max().item() < 5e-6


@pytest.mark.skipif(not native.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("causal", [False, True])
def test_oracle_matches_compiled_reference_live(causal):
    rng = np.random.default_rng(7)
    q = rng.standard_normal((3, 96, 64), dtype=np.float32)
    k = rng.standard_normal((3, 96, 64), dtype=np.float32)
    v = rng.standard_normal((3, 96, 64), dtype=np.float32)
    ref = native.ref_cpu_attention(q, k, v, 0.125, causal, threads=2)
    tq, tk, tv = (torch.from_numpy(t).permute(1, 0, 2).unsqueeze(0) for t in (q, k, v))  # (1,S,H,D)
    out, _ = ao.flash_attn_func_ref(tq, tk, tv, softmax_scale=0.125, causal=causal)
    assert np.abs(out[0].permute(1, 0, 2).numpy() - ref).max() < 2e-5


@pytest.mark.parametrize("causal,hk", [(False, 4), (True, 4), (True, 1), (False, 2)])
def test_oracle_matches_sdpa_cpu(causal, hk):
    torch.manual_seed(0)
    B, S, H, D = 2, 80, 4, 32
    q = torch.randn(B, S, H, D)
    k = torch.randn(B, S, hk, D)
    v = torch.randn(B, S, hk, D)
    out, _ = ao.flash_attn_func_ref(q, k, v, causal=causal)
    sd = torch.nn.functional.scaled_dot_product_attention(
        q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=causal, enable_gqa=hk != H).transpose(1, 2)
    assert (out.float() - sd).abs().max().item() < 2e-5


def test_oracle_window_and_bottom_right_alignment_vs_sdpa_mask():
    torch.manual_seed(1)
    B, Sq, Sk, H, D = 1, 40, 100, 2, 16
    q, k, v = torch.randn(B, Sq, H, D), torch.randn(B, Sk, H, D), torch.randn(B, Sk, H, D)
    wl, wr = 17, 3
    out, _ = ao.flash_attn_func_ref(q, k, v, window_size=(wl, wr))
    i = torch.arange(Sq).view(-1, 1) + (Sk - Sq)
    j = torch.arange(Sk).view(1, -1)
    keep = (j <= i + wr) & (j >= i - wl)
    sd = torch.nn.functional.scaled_dot_product_attention(
        q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), attn_mask=keep).transpose(1, 2)
    assert (out.float() - sd).abs().max().item() < 2e-5


def test_oracle_rows_without_keys_are_zero_with_sentinel_lse():
    torch.manual_seed(2)
    q, k, v = torch.randn(1, 8, 1, 16), torch.randn(1, 3, 1, 16), torch.randn(1, 3, 1, 16)
    out, lse = ao.flash_attn_func_ref(q, k, v, causal=True)  # Sq > Sk: first 5 rows see nothing
    assert out[0, :5].abs().max().item() == 0.0
    assert torch.all(lse[0, 0, :5] == ao.NEG_SENTINEL)
    assert torch.isfinite(lse[0, 0, 5:]).all()


def test_oracle_empty_kv_matches_reference_wrapper():
    out, lse = ao.flash_attn_func_ref(torch.randn(1, 4, 2, 8), torch.zeros(1, 0, 2, 8), torch.zeros(1, 0, 2, 8))
    assert out.abs().max().item() == 0.0 and torch.isinf(lse).all() and (lse < 0).all()


@pytest.mark.parametrize("softcap,alibi", [(0.0, True), (15.0, False), (15.0, True)])
def test_c_oracle_matches_python_oracle_features(softcap, alibi):
    torch.manual_seed(3)
    Sq, Sk, H, Hk, D = 37, 91, 4, 2, 32
    q, k, v = torch.randn(Sq, H, D), torch.randn(Sk, Hk, D), torch.randn(Sk, Hk, D)
    slopes = torch.rand(H) * 0.2 if alibi else None
    o, l = ao.attention_one(q, k, v, 0.2, 20, 0, slopes, softcap)
    oc, lc = native.c_oracle_attention(q.numpy(), k.numpy(), v.numpy(), 0.2, 20, 0,
                                       slopes.numpy() if alibi else None, softcap, threads=2)
    assert np.abs(o.numpy() - oc).max() < 1e-5
    assert np.abs(l.numpy() - lc).max() < 1e-5


def test_varlen_oracle_equals_per_sequence_dense():
    torch.manual_seed(4)
    lens = [5, 33, 1, 64]
    H, Hk, D = 4, 2, 16
    T = sum(lens)
    q, k, v = torch.randn(T, H, D), torch.randn(T, Hk, D), torch.randn(T, Hk, D)
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32)
    out, lse = ao.flash_attn_varlen_func_ref(q, k, v, cu, cu, max(lens), max(lens), causal=True)
    assert lse.shape == (H, T)
    for b, n in enumerate(lens):
        s, e = int(cu[b]), int(cu[b + 1])
        o, l = ao.flash_attn_func_ref(q[s:e][None], k[s:e][None], v[s:e][None], causal=(max(lens) != 1))
        assert (out[s:e] - o[0]).abs().max().item() < 1e-12
        assert (lse[:, s:e] - l[0]).abs().max().item() < 1e-12


def test_paged_oracle_equals_contiguous():
    torch.manual_seed(5)
    B, H, Hk, D, page = 3, 4, 2, 16, 256
    lens = [300, 17, 512]
    n_pages = 8
    kc, vc = torch.randn(n_pages, page, Hk, D), torch.randn(n_pages, page, Hk, D)
    bt = torch.tensor([[3, 1], [7, 0], [2, 5]], dtype=torch.int32)
    q = torch.randn(B, 1, H, D)
    out, lse, _, _ = ao.flash_attn_with_kvcache_ref(q, kc, vc, cache_seqlens=torch.tensor(lens, dtype=torch.int32),
                                                    block_table=bt)
    for b in range(B):
        kk = torch.cat([kc[int(p)] for p in bt[b]])[: lens[b]]
        vv = torch.cat([vc[int(p)] for p in bt[b]])[: lens[b]]
        o, _ = ao.flash_attn_func_ref(q[b][None], kk[None], vv[None])
        assert (out[b] - o[0]).abs().max().item() < 1e-12


@pytest.mark.parametrize("interleaved", [True, False])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_rope_python_matches_c_fmaf(interleaved, dtype):
    """Python RoPE (float64 emulation of the fused step) vs the C one that uses fmaf like the reference."""
    torch.manual_seed(6)
    S, H, D, rot = 9, 2, 64, 32
    x = torch.randn(S, H, D).to(dtype)
    ang = torch.rand(40, rot // 2) * 6.28
    cos, sin = ang.cos().to(dtype), ang.sin().to(dtype)
    pos = torch.arange(S) + 11
    y = ao.apply_rotary_ref(x, cos, sin, pos, interleaved)
    lib = native.load_c_oracle()
    yc = torch.empty(S, H, D)
    f32p = ctypes.POINTER(ctypes.c_float)
    for s in range(S):
        for h in range(H):
            xin = x[s, h].float().contiguous().numpy()
            yo = np.empty(D, dtype=np.float32)
            c = cos[pos[s]].float().contiguous().numpy()
            sn = sin[pos[s]].float().contiguous().numpy()
            lib.oracle_rope(xin.ctypes.data_as(f32p), yo.ctypes.data_as(f32p), c.ctypes.data_as(f32p),
                            sn.ctypes.data_as(f32p), D, rot, int(interleaved))
            yc[s, h] = torch.from_numpy(yo)
    yc = yc.to(dtype)
    # identical except (at most) a double-rounding tie: allow 1 ulp of the 16-bit type on < 0.1% of entries
    diff = (y.float() - yc.float()).abs()
    assert (diff > 0).float().mean().item() < 1e-3
    assert torch.equal(y[..., rot:], x[..., rot:])


def test_kvcache_oracle_appends_in_place_and_rotates():
    torch.manual_seed(7)
    B, H, Hk, D, cap = 2, 4, 2, 32, 64
    kc, vc = torch.randn(B, cap, Hk, D).half(), torch.randn(B, cap, Hk, D).half()
    lens = torch.tensor([10, 31], dtype=torch.int32)
    kn, vn = torch.randn(B, 2, Hk, D).half(), torch.randn(B, 2, Hk, D).half()
    ang = torch.rand(cap, D // 2) * 6.28
    cos, sin = ang.cos().half(), ang.sin().half()
    q = torch.randn(B, 2, H, D).half()
    out, lse, kc2, vc2 = ao.flash_attn_with_kvcache_ref(q, kc, vc, kn, vn, cos, sin, lens, causal=True,
                                                        rotary_interleaved=False)
    for b in range(B):
        L = int(lens[b])
        assert torch.equal(vc2[b, L:L + 2], vn[b])
        assert torch.equal(kc2[b, :L], kc[b, :L]) and torch.equal(kc2[b, L + 2:], kc[b, L + 2:])
        assert not torch.equal(kc2[b, L:L + 2], kn[b])  # rotated
    assert out.shape == (B, 2, H, D) and lse.shape == (B, H, 2)


def test_tolerance_rule_is_the_references():
    ref = torch.zeros(4)
    naive = torch.tensor([0.0, 0.01, 0.0, 0.0])
    ok, err, en = ao.fa_tolerance_ok(torch.tensor([0.0, 0.0, 0.02, 0.0]), ref, naive)
    assert ok and abs(err - 0.02) < 1e-9 and abs(en - 0.01) < 1e-9
    ok, _, _ = ao.fa_tolerance_ok(torch.tensor([0.0, 0.0, 0.0201, 0.0]), ref, naive)
    assert not ok
    ok, _, _ = ao.fa_tolerance_ok(torch.tensor([float("nan"), 0, 0, 0]), ref, naive)
    assert not ok


# ----------------------------------------------------------------------------- dropout (Philox)
def test_philox4x32_10_known_answers():
    """Random123 known-answer vectors for philox4x32-10 (kat_vectors of the published library): pins the
    generator the reference re-implements in include/philox.h."""
    import numpy as np

    def run(c, k):
        lo = np.array([c[0] | (c[1] << 32)], dtype=np.uint64)
        return [int(x) for x in ao.philox4x32_10(lo, k[0] | (k[1] << 32), c[2] | (c[3] << 32))[0]]

    assert run([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert run([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert run([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_dropout_keep_mask_properties():
    m = ao.dropout_keep_mask(0.25, 1234, 8, 0, 256, 512, 512)
    assert abs(m.float().mean().item() - 0.75) < 0.01
    # pure function of the flat index: a row offset of r equals starting r rows later
    m2 = ao.dropout_keep_mask(0.25, 1234, 8, 100, 50, 512, 512)
    assert torch.equal(m[100:150], m2)
    # counter offset of 1 = 4 flat indices
    m3 = ao.dropout_keep_mask(0.25, 1234, 9, 0, 1, 508, 512)
    assert torch.equal(m3[0], m[0, 4:])
    assert not torch.equal(ao.dropout_keep_mask(0.25, 1235, 8, 0, 8, 64, 64), m[:8, :64])


def test_dropout_oracle_is_unbiased_and_keeps_lse():
    torch.manual_seed(0)
    q, k, v = (torch.randn(1, 64, 2, 32) for _ in range(3))
    base, lse0 = ao.flash_attn_func_ref(q, k, v, causal=True)
    acc = torch.zeros_like(base)
    n = 200
    for t in range(n):
        o, lse = ao.flash_attn_func_ref(q, k, v, causal=True, dropout_p=0.3, rng_state=(7, 4096 * t))
        assert torch.equal(lse, lse0)
        acc += o
    assert (acc / n - base).abs().max().item() < 0.35  # Monte-Carlo mean of 200 masks
    assert (acc / n - base).abs().mean().item() < 0.05


# ----------------------------------------------------------------------------- backward
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "ref_mha_backward_*.npz"))))
def test_backward_oracle_matches_reference_ref_mha_backward(path):
    """Fixtures produced by executing the reference's `ref_mha_backward` (test.py:36-61) verbatim."""
    g = np.load(path)
    q, k, v, do = (torch.from_numpy(g[n]).permute(0, 2, 1, 3) for n in ("q", "k", "v", "do"))
    dq, dk, dv, _ = ao.flash_attn_bwd_ref(do, q, k, v, softmax_scale=float(g["scale"]), causal=bool(g["causal"]))
    for name, got in (("dq", dq), ("dk", dk), ("dv", dv)):
        want = torch.from_numpy(g[name]).permute(0, 2, 1, 3).double()
        assert (got - want).abs().max().item() < 5e-6, name  # the fixture is fp32 arithmetic


@pytest.mark.parametrize("kw", [
    dict(), dict(causal=True), dict(window_size=(7, 3)), dict(causal=True, alibi=True), dict(softcap=2.0),
    dict(causal=True, dropout_p=0.3, rng_state=(5, 8)),
])
def test_backward_oracle_formulas_match_autograd(kw):
    """The explicit restatement of the reference's gradient formulas equals autograd through the forward oracle."""
    torch.manual_seed(0)
    B, Sq, Sk, H, Hk, D = 2, 37, 53, 4, 2, 16
    q = torch.randn(B, Sq, H, D, dtype=torch.float64, requires_grad=True)
    k = torch.randn(B, Sk, Hk, D, dtype=torch.float64, requires_grad=True)
    v = torch.randn(B, Sk, Hk, D, dtype=torch.float64, requires_grad=True)
    do = torch.randn(B, Sq, H, D, dtype=torch.float64)
    kw = dict(kw)
    slopes = torch.rand(H) if kw.pop("alibi", False) else None
    p = kw.get("dropout_p", 0.0)
    wl, wr = ao.normalize_mask_args(Sq, Sk, kw.get("causal", False), kw.get("window_size", (-1, -1)), slopes is not None)
    keep = ao.dropout_keep_mask(p, 5, 8, 0, Sq, Sk, Sk) if p > 0 else None
    o = torch.stack([ao.attention_one(q[b], k[b], v[b], D ** -0.5, wl, wr, slopes, kw.get("softcap", 0.0),
                                      torch.float64, keep, p)[0] for b in range(B)])
    want = torch.autograd.grad(o, (q, k, v), do)
    got = ao.flash_attn_bwd_ref(do, q, k, v, alibi_slopes=slopes, **kw)
    for a, b in zip(got[:3], want):
        assert (a - b).abs().max().item() < 1e-12
    assert (got[3] - torch.einsum("bqhd,bqhd->bhq", do, o.detach())).abs().max().item() < 1e-12
