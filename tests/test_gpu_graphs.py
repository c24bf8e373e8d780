"""CUDA-graph capture of the hot path: a serving loop replays one captured decode step (append + rotary +
split-KV attention + combine) and a captured dense forward; replays must match eager calls on the same data.
The library only enqueues kernels on the caller's stream (no allocation, no synchronisation after the first
call on a device), which is what makes its calls capturable."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(fa_lib):
    import flash_attn_v100 as m

    return m


def _rotary(seqlen, rot, dtype):
    inv = 1.0 / (10000 ** (torch.arange(0, rot, 2, dtype=torch.float32) / rot))
    ang = torch.outer(torch.arange(seqlen, dtype=torch.float32), inv)
    return ang.cos().to(dtype).cuda(), ang.sin().to(dtype).cuda()


def test_decode_step_replays_from_a_cuda_graph(api):
    torch.manual_seed(11)
    dt = torch.bfloat16
    B, H, Hk, D, cap, page = 4, 16, 4, 128, 2048, 256
    n_pages = B * (cap // page)
    kc = torch.randn(n_pages, page, Hk, D, device="cuda", dtype=dt)
    vc = torch.randn(n_pages, page, Hk, D, device="cuda", dtype=dt)
    bt = torch.randperm(n_pages, generator=torch.Generator().manual_seed(0)).view(B, -1).int().cuda()
    lens = torch.tensor([100, 1500, 7, 2000], dtype=torch.int32, device="cuda")
    q = torch.randn(B, 1, H, D, device="cuda", dtype=dt)
    kn = torch.randn(B, 1, Hk, D, device="cuda", dtype=dt)
    vn = torch.randn(B, 1, Hk, D, device="cuda", dtype=dt)
    cos, sin = _rotary(cap, D, dt)

    def step(kc_, vc_):
        return api.flash_attn_with_kvcache(q, kc_, vc_, kn, vn, rotary_cos=cos, rotary_sin=sin, cache_seqlens=lens,
                                           block_table=bt, causal=True, rotary_interleaved=False)

    step(kc.clone(), vc.clone())  # first call on the device: one-time set-up must not happen inside the capture
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            out_static = step(kc, vc)
    torch.cuda.current_stream().wait_stream(side)

    for it in range(3):  # three decode steps: new token, lengths advance, all through device memory only
        q.copy_(torch.randn_like(q))
        kn.copy_(torch.randn_like(kn))
        vn.copy_(torch.randn_like(vn))
        kc_ref, vc_ref = kc.clone(), vc.clone()
        ref = step(kc_ref, vc_ref)  # eager, on a copy of the cache as it is before this step
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out_static, ref), f"replay {it} differs from the eager call"
        assert torch.equal(kc, kc_ref) and torch.equal(vc, vc_ref), "the replay must append exactly like the eager call"
        lens.add_(1)


def test_dense_forward_replays_from_a_cuda_graph(api):
    torch.manual_seed(12)
    q = torch.randn(2, 700, 8, 128, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(2, 700, 2, 128, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(2, 700, 2, 128, device="cuda", dtype=torch.bfloat16)
    api.flash_attn_func(q, k, v, causal=True)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out_static = api.flash_attn_func(q, k, v, causal=True)
    for _ in range(3):
        q.copy_(torch.randn_like(q))
        k.copy_(torch.randn_like(k))
        g.replay()
        ref = api.flash_attn_func(q, k, v, causal=True)
        torch.cuda.synchronize()
        assert torch.equal(out_static, ref)


def test_graph_replay_overlapping_many_eager_launches_on_another_stream(api):
    """A captured forward keeps no per-launch scheduler state: replays may overlap any number of eager launches
    on other streams (the round-1 scheduler took a counter from a 1024-slot ring, which a replay running beside
    > 1024 eager launches could share; work is now handed out by cluster launch control, with no global state)."""
    torch.manual_seed(13)
    dt = torch.bfloat16
    q = torch.randn(2, 2048, 16, 128, device="cuda", dtype=dt)  # 256 work items: more than the SM count
    k = torch.randn(2, 2048, 4, 128, device="cuda", dtype=dt)
    v = torch.randn(2, 2048, 4, 128, device="cuda", dtype=dt)
    qs = torch.randn(1, 300, 2, 128, device="cuda", dtype=dt)
    ref_big = api.flash_attn_func(q, k, v, causal=True).clone()
    ref_small = api.flash_attn_func(qs, qs, qs, causal=True).clone()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out_static = api.flash_attn_func(q, k, v, causal=True)
    side = torch.cuda.Stream()
    for rep in range(3):
        out_static.zero_()
        torch.cuda.synchronize()
        outs = []
        for i in range(1200):
            if i % 40 == 0:
                g.replay()  # on the current stream, while the side stream keeps launching
            with torch.cuda.stream(side):
                o = api.flash_attn_func(qs, qs, qs, causal=True)
                if i % 100 == 0:
                    outs.append(o)
        torch.cuda.synchronize()
        assert torch.equal(out_static, ref_big), f"replay {rep}: captured forward corrupted by concurrent launches"
        for o in outs:
            assert torch.equal(o, ref_small)


# ------------------------------------------------------------------------------------------------------------------
# torch.compile: the reason SURVEY 8(f)3 lists the dispatcher registration
# ------------------------------------------------------------------------------------------------------------------
def test_torch_compile_fullgraph_forward_and_backward(api):
    import torch._dynamo

    torch._dynamo.reset()
    torch.manual_seed(21)
    dt = torch.bfloat16
    q = torch.randn(2, 300, 8, 128, device="cuda", dtype=dt, requires_grad=True)
    k = torch.randn(2, 300, 2, 128, device="cuda", dtype=dt, requires_grad=True)
    v = torch.randn(2, 300, 2, 128, device="cuda", dtype=dt, requires_grad=True)
    w = torch.randn(128, 128, device="cuda", dtype=dt)

    def block(q, k, v):
        o = api.flash_attn_func(q * 1.0, k, v, causal=True)  # surrounded by ordinary torch ops
        return (o @ w).float().square().mean()

    compiled = torch.compile(block, fullgraph=True)  # fullgraph: any graph break is an error
    loss_c = compiled(q, k, v)
    gc = torch.autograd.grad(loss_c, (q, k, v))
    loss_e = block(q, k, v)
    ge = torch.autograd.grad(loss_e, (q, k, v))
    assert torch.allclose(loss_c, loss_e, rtol=1e-3, atol=1e-5)
    for a, b in zip(gc, ge):
        assert (a.float() - b.float()).abs().max().item() <= 2e-2 * max(b.float().abs().max().item(), 1e-6) + 1e-6


def test_torch_compile_fullgraph_varlen_and_kvcache(api):
    import torch._dynamo

    torch._dynamo.reset()
    torch.manual_seed(22)
    dt = torch.float16
    lens = [37, 200, 1, 129]
    cu = torch.tensor([0, 37, 237, 238, 367], dtype=torch.int32, device="cuda")
    q = torch.randn(sum(lens), 4, 64, device="cuda", dtype=dt)
    k = torch.randn(sum(lens), 4, 64, device="cuda", dtype=dt)
    v = torch.randn(sum(lens), 4, 64, device="cuda", dtype=dt)
    f = torch.compile(lambda q, k, v: api.flash_attn_varlen_func(q, k, v, cu, cu, 200, 200, causal=True) + 0.0, fullgraph=True)
    assert torch.equal(f(q, k, v), api.flash_attn_varlen_func(q, k, v, cu, cu, 200, 200, causal=True))

    kc = torch.randn(2, 512, 2, 64, device="cuda", dtype=dt)
    vc = torch.randn(2, 512, 2, 64, device="cuda", dtype=dt)
    qd = torch.randn(2, 1, 8, 64, device="cuda", dtype=dt)
    kn = torch.randn(2, 1, 2, 64, device="cuda", dtype=dt)
    vn = torch.randn(2, 1, 2, 64, device="cuda", dtype=dt)
    sl = torch.tensor([100, 7], dtype=torch.int32, device="cuda")
    g = torch.compile(lambda qd, kc, vc: api.flash_attn_with_kvcache(qd, kc, vc, kn, vn, cache_seqlens=sl, causal=True) * 1.0,
                      fullgraph=True)
    kc1, vc1, kc2, vc2 = kc.clone(), vc.clone(), kc.clone(), vc.clone()
    out_c = g(qd, kc1, vc1)
    out_e = api.flash_attn_with_kvcache(qd, kc2, vc2, kn, vn, cache_seqlens=sl, causal=True)
    assert torch.equal(out_c, out_e) and torch.equal(kc1, kc2) and torch.equal(vc1, vc2)  # appended in place in both
