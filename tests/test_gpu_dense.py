"""GPU parity: dense forward through the public API / operator layer / C ABI vs the CPU oracle.

Mirrors the reference's own test (test.py::test_combined): same seed, randn inputs, 1/sqrt(D) scale,
causal in {False, True}, the raw-extension call of test.py:227 with 13 positional arguments in
[B,H,M,D] layout, and its pass rule (test.py:273-277) -- plus the features test.py never covers.
"""
import glob
import os

import numpy as np
import pytest
import torch

from gpu_utils import check_dense, rand_qkv, sampled_row_check

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


# shapes of reference test.py:115-139 that the CPU oracle finishes in seconds
REF_SHAPES = [(1, 1, 16, 16, 16), (1, 1, 32, 32, 32), (1, 1, 64, 64, 64), (1, 1, 128, 128, 128), (1, 1, 256, 256, 256),
              (1, 16, 1024, 1024, 16), (1, 16, 1024, 1024, 32), (1, 16, 1024, 1024, 64), (1, 16, 1024, 1024, 128),
              (1, 16, 1024, 1024, 256)]


@pytest.mark.parametrize("B,H,M,N,D", REF_SHAPES)
@pytest.mark.parametrize("causal", [False, True])
def test_reference_test_py_shapes_raw_fwd_fp16(op, B, H, M, N, D, causal):
    torch.manual_seed(421)
    q = torch.randn(B, H, M, D, device="cuda", dtype=torch.float16)
    k = torch.randn(B, H, N, D, device="cuda", dtype=torch.float16)
    v = torch.randn(B, H, N, D, device="cuda", dtype=torch.float16)
    scale = 1.0 / (D ** 0.5)
    before = op.launch_count()
    out, lse, dmask, rng = op.fwd(q, k, v, None, None, 0.0, scale, causal, -1, -1, 0.0, False, None)  # test.py:227
    torch.cuda.synchronize()
    assert op.launch_count() == before + 1
    assert out.shape == q.shape and lse.shape == (B, H, M) and lse.dtype == torch.float32
    assert dmask.numel() == 0 and rng.shape == (2,)
    t = lambda x: x.permute(0, 2, 1, 3)
    check_dense(t(out), lse, t(q), t(k), t(v), causal=causal, scale=scale)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "ref_mha_forward_*.npz"))))
def test_golden_ref_mha_forward_fixtures(op, path):
    """Kernel vs outputs the reference's own oracle produced (tests/golden/make_golden.py)."""
    g = np.load(path)
    q, k, v = (torch.from_numpy(g[n]).cuda() for n in ("q", "k", "v"))
    out, lse, _, _ = op.fwd(q, k, v, None, None, 0.0, float(g["scale"]), bool(g["causal"]), -1, -1, 0.0, False, None)
    err = (out.float().cpu() - torch.from_numpy(g["out"])).abs().max().item()
    assert err < 3e-3, err  # fp16 output rounding of O(1) values + fp16 P


def test_golden_reference_cpu_attention_known_answer(op):
    """The reference harness's case (forward_kernel.cu:439): D=128 causal 128x128, tol 5e-2 (:433)."""
    g = np.load(os.path.join(GOLD, "ref_cpu_attention_kat.npz"))
    q, k, v = (torch.from_numpy(g[n]).half().cuda().view(1, 1, 128, 128) for n in ("q", "k", "v"))
    out, _, _, _ = op.fwd(q, k, v, None, None, 0.0, 0.125, True, -1, -1, 0.0, False, None)
    err = (out.float().cpu().view(128, 128) - torch.from_numpy(g["out"])).abs().max().item()
    assert err < 5e-2 and err < 2e-3, err


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,Sq,Sk,H,Hk,D,causal", [
    (2, 512, 512, 8, 8, 64, False),      # BASELINE config 1 (fp16 is the reference dtype)
    (2, 512, 512, 4, 4, 128, True),
    (1, 1024, 1024, 8, 2, 128, True),    # GQA
    (1, 256, 256, 6, 1, 64, True),       # MQA
    (2, 333, 777, 4, 2, 128, True),      # ragged, Sq != Sk, bottom-right aligned
    (2, 777, 333, 4, 2, 128, True),      # Sq > Sk: leading rows see no key
    (1, 1, 1000, 4, 4, 128, True),       # Sq == 1 => causal normalised away
    (3, 130, 129, 2, 2, 64, False),      # N not a multiple of 16 (wrong in the reference's dense path)
    (1, 257, 255, 2, 1, 128, False),
])
def test_dense_api_vs_oracle(api, dtype, B, Sq, Sk, H, Hk, D, causal):
    q, k, v = rand_qkv(B, Sq, Sk, H, Hk, D, dtype)
    out = api.flash_attn_func(q, k, v, causal=causal)
    assert out.shape == q.shape and out.dtype == dtype and out.is_contiguous()
    check_dense(out, None, q, k, v, causal=causal)


@pytest.mark.parametrize("window", [(64, 0), (100, 30), (-1, 17), (300, -1), (0, 0), (5000, 5000)])
def test_sliding_window(api, window):
    q, k, v = rand_qkv(2, 600, 700, 4, 2, 128, torch.bfloat16)
    out = api.flash_attn_func(q, k, v, window_size=window)
    check_dense(out, None, q, k, v, window=window)
    out = api.flash_attn_func(q, k, v, causal=True, window_size=window)
    check_dense(out, None, q, k, v, causal=True, window=window)


@pytest.mark.parametrize("softcap,alibi,causal", [(30.0, False, False), (0.0, True, True), (0.0, True, False),
                                                 (20.0, True, True)])
@pytest.mark.parametrize("D", [64, 128, 256])
def test_softcap_and_alibi(api, softcap, alibi, causal, D):
    B, H = 2, 4
    q, k, v = rand_qkv(B, 384, 384, H, 2, D, torch.bfloat16)
    slopes = None
    if alibi:
        slopes = (torch.rand(B, H, device="cuda") * 0.3).float() if softcap == 0.0 else (torch.rand(H, device="cuda") * 0.3).float()
    out = api.flash_attn_func(q, k, v, causal=causal, softcap=softcap, alibi_slopes=slopes)
    check_dense(out, None, q, k, v, causal=causal, softcap=softcap, alibi=slopes)


@pytest.mark.parametrize("D", [16, 32, 40, 80, 96, 100, 136, 192, 250, 256])
def test_head_dims_any_multiple_of_8(api, D):
    q, k, v = rand_qkv(1, 200, 264, 4, 4, D, torch.float16)
    out = api.flash_attn_func(q, k, v, causal=True)
    assert out.shape == q.shape
    check_dense(out, None, q, k, v, causal=True)


@pytest.mark.parametrize("D,causal", [(136, True), (160, False), (184, True), (192, True), (200, False), (256, True)])
def test_wide_tile_head_dims_many_tiles(api, D, causal):
    """Head dims 129..256 share the 256-wide tile (one stage, K/V handed over in halves); up to 192 the tile's fourth
    64-column block is all padding and is neither loaded nor multiplied. Several tiles per item, several items, GQA."""
    q, k, v = rand_qkv(2, 700, 1100, 6, 2, D, torch.bfloat16)
    out = api.flash_attn_func(q, k, v, causal=causal)
    check_dense(out, None, q, k, v, causal=causal)


def test_lse_and_return_conventions(api, op):
    q, k, v = rand_qkv(2, 300, 300, 4, 4, 128, torch.bfloat16)
    # dense return_attn_probs=True with dropout_p=0 raises, like the reference (fused_mha_forward.cu:371)
    with pytest.raises(RuntimeError, match="return_softmax requires p_dropout > 0"):
        api.flash_attn_func(q, k, v, return_attn_probs=True)
    t = lambda x: x.permute(0, 2, 1, 3)
    out, lse, _, _ = op.fwd(t(q), t(k), t(v), None, None, 0.0, 128 ** -0.5, True, -1, -1, 0.0, False, None)
    check_dense(t(out), lse, q, k, v, causal=True)
    with pytest.warns(RuntimeWarning):
        api.flash_attn_func(q, k, v, deterministic=True)


def test_strided_inputs_and_out_argument(op):
    """[B,H,M,D] contract by strides: both a contiguous [B,H,M,D] tensor and a permuted [B,M,H,D] view."""
    B, H, M, D = 2, 4, 320, 128
    torch.manual_seed(421)
    q = torch.randn(B, H, M, D, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(B, H, M, D, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(B, H, M, D, device="cuda", dtype=torch.bfloat16)
    o1 = op.fwd(q, k, v, None, None, 0.0, D ** -0.5, True, -1, -1, 0.0, False, None)[0]
    qv, kv, vv = (x.permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3) for x in (q, k, v))
    out_buf = torch.empty_like(q)
    o2 = op.fwd(qv, kv, vv, out_buf, None, 0.0, D ** -0.5, True, -1, -1, 0.0, False, None)[0]
    assert o2.data_ptr() == out_buf.data_ptr()
    assert torch.equal(o1, o2)


def test_empty_kv_and_errors(op, api):
    q = torch.randn(1, 2, 8, 64, device="cuda", dtype=torch.float16)
    k = torch.zeros(1, 2, 0, 64, device="cuda", dtype=torch.float16)
    out, lse, _, _ = op.fwd(q, k, k, None, None, 0.0, 1.0, False, -1, -1, 0.0, False, None)
    assert out.abs().max().item() == 0 and torch.isinf(lse).all()
    with pytest.raises(RuntimeError, match="fp16 or bf16"):
        op.fwd(q.float(), q.float(), q.float(), None, None, 0.0, 1.0, False, -1, -1, 0.0, False, None)
    with pytest.raises(RuntimeError, match="divisible"):
        op.fwd(torch.zeros(1, 3, 8, 64, device="cuda", dtype=torch.float16), q, q, None, None, 0.0, 1.0, False, -1, -1, 0.0, False, None)
    with pytest.raises(RuntimeError, match="Softcapping does not support dropout"):
        api.flash_attn_func(q.permute(0, 2, 1, 3), q.permute(0, 2, 1, 3), q.permute(0, 2, 1, 3), dropout_p=0.1, softcap=10.0)


def test_config2_full_size_properties(api, op):
    """BASELINE config 2 at full size (bf16 B=8 H=32 S=4096 D=128 causal): size-independent checks."""
    B, S, H, D = 8, 4096, 32, 128
    q, k, v = rand_qkv(B, S, S, H, H, D, torch.bfloat16)
    t = lambda x: x.permute(0, 2, 1, 3)
    out_, lse, _, _ = op.fwd(t(q), t(k), t(v), None, None, 0.0, D ** -0.5, True, -1, -1, 0.0, False, None)
    out = t(out_)
    assert torch.isfinite(out.float()).all()
    # (1) sampled rows recomputed exactly on the CPU
    g = torch.Generator().manual_seed(0)
    rows = [(int(torch.randint(B, (1,), generator=g)), int(torch.randint(H, (1,), generator=g)),
             int(torch.randint(S, (1,), generator=g))) for _ in range(48)]
    rows += [(0, 0, 0), (B - 1, H - 1, S - 1), (3, 7, 127), (3, 7, 128), (5, 1, 255), (5, 1, 256)]
    sampled_row_check(out, lse, q, k, v, rows, causal=True)
    # (2) linearity in V: scaling V by 2 is exact in bf16, so the result must double bit-for-bit
    out2 = api.flash_attn_func(q, k, v * 2, causal=True)
    assert torch.equal(out2, out * 2)
    # (3) independence of (batch, head) problems: a slice computed alone is bit-identical
    sl = api.flash_attn_func(q[2:3, :, 5:6], k[2:3, :, 5:6], v[2:3, :, 5:6], causal=True)
    assert torch.equal(sl, out[2:3, :, 5:6])
    # (4) causality: perturbing keys/values after position p leaves rows <= p unchanged
    k2, v2 = k.clone(), v.clone()
    k2[:, 3000:] = torch.randn_like(k2[:, 3000:])
    v2[:, 3000:] = torch.randn_like(v2[:, 3000:])
    out3 = api.flash_attn_func(q, k2, v2, causal=True)
    assert torch.equal(out3[:, :3000], out[:, :3000])
    # (5) LSE is the log of a sum of S' positive terms: rows are bounded by max score + ln(i+1)
    assert torch.isfinite(lse).all()


def test_config5_window_slice_full_seqlen(api, op):
    """BASELINE config 5 geometry (S=8192, causal + window 4096) on a batch slice: sampled rows."""
    B, S, H, D = 1, 8192, 4, 128
    q, k, v = rand_qkv(B, S, S, H, H, D, torch.bfloat16)
    t = lambda x: x.permute(0, 2, 1, 3)
    out_, lse, _, _ = op.fwd(t(q), t(k), t(v), None, None, 0.0, D ** -0.5, True, 4096, 0, 0.0, False, None)
    rows = [(0, h, i) for h in (0, 3) for i in (0, 1, 4095, 4096, 4097, 5000, 8191, 4223, 4224)]
    sampled_row_check(t(out_), lse, q, k, v, rows, causal=True, window=(4096, 0))


def test_dispatcher_op_matches_the_module_function(op):
    """torch.ops.flash_attn_v100.fwd (reference kernel/fused_mha_api.cpp:308-315) runs the same kernel."""
    torch.manual_seed(421)
    q = torch.randn(1, 4, 256, 128, device="cuda", dtype=torch.bfloat16)
    k, v = torch.randn_like(q), torch.randn_like(q)
    a = op.fwd(q, k, v, None, None, 0.0, 128 ** -0.5, True, -1, -1, 0.0, False, None)
    b = torch.ops.flash_attn_v100.fwd(q, k, v, None, None, 0.0, 128 ** -0.5, True, -1, -1, 0.0, False, None)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_long_sequence_sampled_rows(api, op):
    """S = 65536 (512 KV tiles per block, 256 work items per head): index arithmetic and the persistent
    scheduler at a length far beyond the reference's own tests (max 8192, test.py:138)."""
    B, S, H, D = 1, 65536, 2, 128
    q, k, v = rand_qkv(B, S, S, H, 1, D, torch.bfloat16)
    t = lambda x: x.permute(0, 2, 1, 3)
    out_, lse, _, _ = op.fwd(t(q), t(k), t(v), None, None, 0.0, D ** -0.5, True, -1, -1, 0.0, False, None)
    rows = [(0, 0, 0), (0, 1, 65535), (0, 0, 32767), (0, 1, 32768), (0, 0, 40001), (0, 1, 127), (0, 0, 65280)]
    sampled_row_check(t(out_), lse, q, k, v, rows, causal=True)


def test_concurrent_streams_use_separate_scheduler_slots(api):
    """Two launches in flight on different streams must not share tile-scheduler state."""
    q1, k1, v1 = rand_qkv(2, 2048, 2048, 8, 8, 128, torch.bfloat16, seed=1)
    q2, k2, v2 = rand_qkv(4, 1024, 1024, 16, 4, 128, torch.bfloat16, seed=2)
    ref1 = api.flash_attn_func(q1, k1, v1, causal=True)
    ref2 = api.flash_attn_func(q2, k2, v2, causal=False)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs1, outs2 = [], []
    for _ in range(10):
        with torch.cuda.stream(s1):
            outs1.append(api.flash_attn_func(q1, k1, v1, causal=True))
        with torch.cuda.stream(s2):
            outs2.append(api.flash_attn_func(q2, k2, v2, causal=False))
    torch.cuda.synchronize()
    assert all(torch.equal(o, ref1) for o in outs1) and all(torch.equal(o, ref2) for o in outs2)
