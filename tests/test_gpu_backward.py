"""GPU parity: the backward through the operator layer and autograd vs the CPU oracle.

Mirrors the reference's own backward test (test.py:226-233, 322-334): the raw `bwd` call with its 19
positional arguments in [B,H,M,D] layout on seed-421 fp16 inputs, pass rule
`err <= 3 * err_naive_16bit + 1e-4` per gradient -- plus everything test.py never covers (bf16, GQA, Sq != Sk,
windows, ALiBi, softcap, dropout, var-len, padded head dims, autograd wiring)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import attention_oracle as ao

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def api(fa_lib):
    import flash_attn_v100 as m

    return m


@pytest.fixture(scope="module")
def op(fa_lib):
    import flash_attn_v100_cuda as m

    return m


def _rule(got, ref, naive, name):
    """reference test.py:322-334"""
    err = (got.double().cpu() - ref).abs().max().item()
    err_naive = (naive.double().cpu() - ref).abs().max().item()
    assert bool(torch.isfinite(got.float()).all()), name
    assert err <= 3.0 * err_naive + 1e-4, f"{name}: err {err:.3e} > 3 * naive {err_naive:.3e} + 1e-4"


REF_SHAPES = [(1, 1, 16, 16, 16), (1, 1, 64, 64, 64), (1, 1, 128, 128, 128), (1, 1, 256, 256, 256), (1, 8, 512, 512, 32),
              (1, 8, 1024, 1024, 64), (1, 8, 1024, 1024, 128), (1, 4, 1024, 1024, 256)]


@pytest.mark.parametrize("B,H,M,N,D", REF_SHAPES)
@pytest.mark.parametrize("causal", [False, True])
def test_reference_test_py_shapes_raw_bwd_fp16(op, B, H, M, N, D, causal):
    torch.manual_seed(421)
    q = torch.randn(B, H, M, D, device="cuda", dtype=torch.float16)
    k = torch.randn(B, H, N, D, device="cuda", dtype=torch.float16)
    v = torch.randn(B, H, N, D, device="cuda", dtype=torch.float16)
    dO = torch.randn(B, H, M, D, device="cuda", dtype=torch.float16)
    scale = 1.0 / (D ** 0.5)
    o, lse, _, _ = op.fwd(q, k, v, None, None, 0.0, scale, causal, -1, -1, 0.0, False, None)
    before = op.launch_count()
    dq, dk, dv, sd = op.bwd(dO, q, k, v, o, lse, None, None, None, None, 0.0, scale, causal, -1, -1, 0.0, False, None, None)  # test.py:233
    torch.cuda.synchronize()
    assert op.launch_count() == before + 3  # row-dot, dK/dV pass, dQ pass
    assert dq.shape == q.shape and dk.shape == k.shape and dv.shape == v.shape and sd.shape == (B, H, M)
    t = lambda x: x.permute(0, 2, 1, 3)
    rq, rk, rv, rd = ao.flash_attn_bwd_ref(t(dO), t(q), t(k), t(v), softmax_scale=scale, causal=causal)
    nq, nk, nv = ao.naive_lowp_attention_bwd(t(q), t(k), t(v), t(dO), scale, causal)
    _rule(t(dq), rq, nq, "dQ")
    _rule(t(dk), rk, nk, "dK")
    _rule(t(dv), rv, nv, "dV")
    # softmax_d is computed from the 16-bit `out` the forward stored, so it carries that rounding
    assert (sd.double().cpu() - rd).abs().max().item() <= 2e-2


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "ref_mha_backward_*.npz"))))
def test_golden_ref_mha_backward_fixtures(op, path):
    """Kernel vs gradients the reference's own `ref_mha_backward` produced (tests/golden/make_golden.py)."""
    g = np.load(path)
    q, k, v, do = (torch.from_numpy(g[n]).cuda() for n in ("q", "k", "v", "do"))
    scale, causal = float(g["scale"]), bool(g["causal"])
    o, lse, _, _ = op.fwd(q, k, v, None, None, 0.0, scale, causal, -1, -1, 0.0, False, None)
    dq, dk, dv, _ = op.bwd(do, q, k, v, o, lse, None, None, None, None, 0.0, scale, causal, -1, -1, 0.0, False, None, None)
    for name, got in (("dq", dq), ("dk", dk), ("dv", dv)):
        err = (got.float().cpu() - torch.from_numpy(g[name])).abs().max().item()
        assert err < 1.5e-2, (name, err)  # fp16 P / dS operands, O(1)-magnitude gradients


def _autograd_case(api, dtype, B, Sq, Sk, H, Hk, D, tol_scale=1.0, **kw):
    torch.manual_seed(421)
    q = torch.randn(B, Sq, H, D, device="cuda", dtype=dtype, requires_grad=True)
    k = torch.randn(B, Sk, Hk, D, device="cuda", dtype=dtype, requires_grad=True)
    v = torch.randn(B, Sk, Hk, D, device="cuda", dtype=dtype, requires_grad=True)
    do = torch.randn(B, Sq, H, D, device="cuda", dtype=dtype)
    rng = None
    if kw.get("dropout_p", 0.0) > 0:
        gen = torch.cuda.default_generators[0]
        gen.manual_seed(77)
        rng = (gen.initial_seed(), gen.get_offset())
    out = api.flash_attn_func(q, k, v, **kw)
    dq, dk, dv = torch.autograd.grad(out, (q, k, v), do)
    okw = dict(kw)
    if rng is not None:
        okw["rng_state"] = rng
    rq, rk, rv, _ = ao.flash_attn_bwd_ref(do, q, k, v, **okw)
    tol = (4e-2 if dtype == torch.bfloat16 else 6e-3) * tol_scale
    for name, got, ref in (("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv)):
        assert got.shape == ref.shape and got.dtype == dtype and got.is_contiguous()
        err = (got.double().cpu() - ref).abs().max().item()
        bound = tol * max(1.0, ref.abs().max().item())
        assert err <= bound, f"{name}: {err:.3e} > {bound:.3e}"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,Sq,Sk,H,Hk,D,kw", [
    (2, 256, 256, 4, 4, 128, dict()),
    (2, 384, 384, 4, 2, 128, dict(causal=True)),                       # GQA: dK/dV summed over the group
    (1, 200, 328, 4, 1, 64, dict(causal=True)),                        # MQA, Sq < Sk, ragged tiles
    (1, 328, 200, 2, 2, 64, dict(causal=True)),                        # Sq > Sk: rows without any key
    (1, 512, 512, 2, 2, 128, dict(window_size=(100, 30))),
    (1, 512, 512, 2, 2, 64, dict(causal=True, window_size=(128, -1))),
    (1, 130, 77, 2, 2, 128, dict()),                                    # lengths that are no multiple of anything
    (2, 96, 96, 2, 2, 40, dict(causal=True)),                           # padded head dim
    (1, 256, 256, 4, 2, 128, dict(softcap=15.0)),
    (1, 256, 256, 4, 4, 64, dict(causal=True, softcap=5.0)),
    (1, 256, 256, 2, 2, 256, dict(causal=True)),                        # head dim 256: 64-row streamed tiles, D-split CTAs
    (2, 200, 328, 4, 2, 256, dict()),                                   # + GQA, ragged tiles, Sq < Sk
    (1, 328, 200, 2, 1, 192, dict(causal=True)),                        # 128 < D < 256, rows without any key
    (1, 512, 512, 2, 2, 256, dict(window_size=(100, 30))),
    (1, 256, 256, 2, 2, 256, dict(causal=True, softcap=10.0)),
])
def test_autograd_matches_oracle(api, dtype, B, Sq, Sk, H, Hk, D, kw):
    _autograd_case(api, dtype, B, Sq, Sk, H, Hk, D, **kw)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_autograd_alibi(api, dtype):
    slopes = torch.tensor([0.5, 0.25, 0.125, 0.0625], device="cuda", dtype=torch.float32)
    _autograd_case(api, dtype, 2, 256, 320, 4, 2, 64, causal=True, alibi_slopes=slopes)
    _autograd_case(api, dtype, 2, 256, 256, 4, 4, 128, alibi_slopes=slopes.repeat(2, 1).contiguous() * torch.tensor([[1.0], [0.5]], device="cuda"))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,Sq,Sk,H,Hk,D,causal,p", [
    (2, 256, 256, 4, 2, 128, True, 0.1),
    (1, 200, 328, 2, 2, 64, False, 0.3),
    (1, 130, 203, 2, 1, 64, True, 0.25),   # Sk % 4 != 0: unaligned Philox words in both passes
    (1, 200, 264, 2, 1, 256, True, 0.2),   # head dim 256
])
def test_autograd_dropout(api, dtype, B, Sq, Sk, H, Hk, D, causal, p):
    _autograd_case(api, dtype, B, Sq, Sk, H, Hk, D, tol_scale=1.0 / (1 - p), causal=causal, dropout_p=p)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("kw", [dict(causal=True), dict(), dict(causal=True, dropout_p=0.2), dict(window_size=(64, 0))])
def test_varlen_autograd_matches_oracle(api, dtype, kw):
    _varlen_autograd_case(api, dtype, kw, 128)


@pytest.mark.parametrize("kw", [dict(causal=True), dict(window_size=(64, 0))])
def test_varlen_autograd_head_dim_256(api, kw):
    _varlen_autograd_case(api, torch.bfloat16, kw, 256)


def _varlen_autograd_case(api, dtype, kw, D):
    torch.manual_seed(421)
    lens = [37, 256, 1, 300, 129]
    H, Hk = 4, 2
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    T = sum(lens)
    q = torch.randn(T, H, D, device="cuda", dtype=dtype, requires_grad=True)
    k = torch.randn(T, Hk, D, device="cuda", dtype=dtype, requires_grad=True)
    v = torch.randn(T, Hk, D, device="cuda", dtype=dtype, requires_grad=True)
    do = torch.randn(T, H, D, device="cuda", dtype=dtype)
    rng = None
    if kw.get("dropout_p", 0.0) > 0:
        gen = torch.cuda.default_generators[0]
        gen.manual_seed(78)
        rng = (gen.initial_seed(), gen.get_offset())
    out = api.flash_attn_varlen_func(q, k, v, cu, cu, max(lens), max(lens), **kw)
    dq, dk, dv = torch.autograd.grad(out, (q, k, v), do)
    okw = dict(kw)
    if rng is not None:
        okw["rng_state"] = rng
    rq, rk, rv, _ = ao.flash_attn_varlen_bwd_ref(do, q, k, v, cu, cu, max(lens), max(lens), **okw)
    tol = (4e-2 if dtype == torch.bfloat16 else 6e-3) / (1 - kw.get("dropout_p", 0.0))
    for name, got, ref in (("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv)):
        err = (got.double().cpu() - ref).abs().max().item()
        assert err <= tol * max(1.0, ref.abs().max().item()), f"{name}: {err:.3e}"


def test_varlen_different_q_and_k_lengths(api):
    torch.manual_seed(1)
    lq, lk = [64, 200, 5], [128, 131, 300]
    cuq = torch.tensor([0] + list(torch.tensor(lq).cumsum(0)), dtype=torch.int32, device="cuda")
    cuk = torch.tensor([0] + list(torch.tensor(lk).cumsum(0)), dtype=torch.int32, device="cuda")
    q = torch.randn(sum(lq), 4, 64, device="cuda", dtype=torch.float16, requires_grad=True)
    k = torch.randn(sum(lk), 4, 64, device="cuda", dtype=torch.float16, requires_grad=True)
    v = torch.randn(sum(lk), 4, 64, device="cuda", dtype=torch.float16, requires_grad=True)
    do = torch.randn_like(q)
    out = api.flash_attn_varlen_func(q, k, v, cuq, cuk, max(lq), max(lk), causal=True)
    dq, dk, dv = torch.autograd.grad(out, (q, k, v), do)
    rq, rk, rv, _ = ao.flash_attn_varlen_bwd_ref(do, q, k, v, cuq, cuk, max(lq), max(lk), causal=True)
    for got, ref in ((dq, rq), (dk, rk), (dv, rv)):
        assert (got.double().cpu() - ref).abs().max().item() <= 6e-3 * max(1.0, ref.abs().max().item())


def test_backward_is_deterministic_and_validates(op):
    torch.manual_seed(3)
    q = torch.randn(2, 4, 512, 128, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(2, 2, 512, 128, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(2, 2, 512, 128, device="cuda", dtype=torch.bfloat16)
    do = torch.randn_like(q)
    o, lse, _, _ = op.fwd(q, k, v, None, None, 0.0, 0.088, True, -1, -1, 0.0, False, None)
    a = op.bwd(do, q, k, v, o, lse, None, None, None, None, 0.0, 0.088, True, -1, -1, 0.0, True, None, None)
    b = op.bwd(do, q, k, v, o, lse, None, None, None, None, 0.0, 0.088, True, -1, -1, 0.0, False, None, None)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    # caller-provided gradient buffers are filled in place
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    c = op.bwd(do, q, k, v, o, lse, dq, dk, dv, None, 0.0, 0.088, True, -1, -1, 0.0, False, None, None)
    assert c[0].data_ptr() == dq.data_ptr() and torch.equal(dq, a[0]) and torch.equal(dk, a[1]) and torch.equal(dv, a[2])
    with pytest.raises(RuntimeError, match="rng_state required"):
        op.bwd(do, q, k, v, o, lse, None, None, None, None, 0.1, 0.088, True, -1, -1, 0.0, False, None, None)
    with pytest.raises(RuntimeError, match="softmax_lse must be fp32"):
        op.bwd(do, q, k, v, o, lse.half(), None, None, None, None, 0.0, 0.088, True, -1, -1, 0.0, False, None, None)
    with pytest.raises(RuntimeError, match="dq shape must match q shape"):
        op.bwd(do, q, k, v, o, lse, dk, None, None, None, 0.0, 0.088, True, -1, -1, 0.0, False, None, None)


def test_config2_geometry_backward_sampled(api):
    """BASELINE config-2 geometry (bf16 S=4096 D=128 causal, one batch element, GQA 8:2): gradients of a few
    sampled rows / keys recomputed exactly on the CPU, plus linearity in dO."""
    torch.manual_seed(421)
    S, H, Hk, D = 4096, 8, 2, 128
    q = torch.randn(1, S, H, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    k = torch.randn(1, S, Hk, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    v = torch.randn(1, S, Hk, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    do = torch.randn(1, S, H, D, device="cuda", dtype=torch.bfloat16)
    out = api.flash_attn_func(q, k, v, causal=True)
    dq, dk, dv = torch.autograd.grad(out, (q, k, v), do, retain_graph=True)
    dq2, dk2, dv2 = torch.autograd.grad(out, (q, k, v), 2 * do)
    for a, b in ((dq, dq2), (dk, dk2), (dv, dv2)):  # exact: doubling is a power-of-two scaling throughout
        assert torch.equal(2 * a.float(), b.float())
    # exact float64 gradients restricted to the first 640 positions (causal: they only see each other)
    n = 640
    rq, rk, rv, _ = ao.flash_attn_bwd_ref(do[:, :n], q[:, :n], k[:, :n], v[:, :n], causal=True)
    assert (dq[:, :n].double().cpu() - rq).abs().max().item() <= 4e-2 * max(1.0, rq.abs().max().item())
    # dK/dV of early keys also receive contributions from later queries, so compare the last keys instead:
    # keys [S-n, S) are only seen by queries [S-n, S)
    qq, kk, vv, dd = q[:, S - n:], k[:, S - n:], v[:, S - n:], do[:, S - n:]
    # their P needs the full-row normaliser: take LSE from the kernel's forward
    _, lse, _ = api.flash_attn_func(q.detach(), k.detach(), v.detach(), causal=True, dropout_p=0.0,
                                    return_attn_probs=False), None, None
    full_q, full_k, full_v = q.detach().double().cpu(), k.detach().double().cpu(), v.detach().double().cpu()
    g = H // Hk
    dk_ref = torch.zeros(n, Hk, D, dtype=torch.float64)
    dv_ref = torch.zeros(n, Hk, D, dtype=torch.float64)
    dof = do.double().cpu()
    for h in range(H):
        hk = h // g
        s = (full_q[0, S - n:, h] @ full_k[0, :, hk].T) * D ** -0.5          # [n, S]
        i = torch.arange(S - n, S).view(-1, 1)
        j = torch.arange(S).view(1, -1)
        s = s.masked_fill(j > i, float("-inf"))
        p = torch.softmax(s, dim=-1)
        o = p @ full_v[0, :, hk]
        delta = (dof[0, S - n:, h] * o).sum(-1, keepdim=True)
        dp = dof[0, S - n:, h] @ full_v[0, :, hk].T
        ds = p * (dp - delta) * D ** -0.5
        dk_ref[:, hk] += ds[:, S - n:].T @ full_q[0, S - n:, h]
        dv_ref[:, hk] += p[:, S - n:].T @ dof[0, S - n:, h]
    assert (dk[0, S - n:].double().cpu() - dk_ref).abs().max().item() <= 4e-2 * max(1.0, dk_ref.abs().max().item())
    assert (dv[0, S - n:].double().cpu() - dv_ref).abs().max().item() <= 4e-2 * max(1.0, dv_ref.abs().max().item())


def test_backward_called_first_from_a_fresh_thread(op):
    """PyTorch runs backward() on its own autograd thread, which may not have touched CUDA before our call:
    the C layer must bind the device's context itself (cuTensorMapEncodeTiled is a driver call)."""
    import threading

    torch.manual_seed(5)
    q = torch.randn(1, 2, 130, 64, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(1, 2, 200, 64, device="cuda", dtype=torch.bfloat16)
    v = torch.randn_like(k)
    do = torch.randn_like(q)
    out, lse, _, rng = op.fwd(q, k, v, None, None, 0.0, 0.125, True, -1, -1, 0.0, False, None)
    expect = op.bwd(do, q, k, v, out, lse, None, None, None, None, 0.0, 0.125, True, -1, -1, 0.0, False, None, rng)
    torch.cuda.synchronize()
    box = {}

    def work():
        try:
            box["res"] = op.bwd(do, q, k, v, out, lse, None, None, None, None, 0.0, 0.125, True, -1, -1, 0.0, False, None, rng)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            box["err"] = e

    t = threading.Thread(target=work)
    t.start()
    t.join()
    assert "err" not in box, box.get("err")
    for a, b in zip(box["res"][:3], expect[:3]):
        assert torch.equal(a, b)
