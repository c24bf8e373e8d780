"""Calls on a device that is not the current one (needs >= 2 GPUs; skipped otherwise): the C layer selects the
device of the tensors, sets the kernel attributes and takes scheduler slots per device, and restores the caller's
current device."""
import pytest
import torch

from oracle import attention_oracle as ao

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(fa_lib):
    import flash_attn_v100 as m

    return m


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_forward_backward_and_decode_on_a_non_current_device(api):
    assert torch.cuda.current_device() == 0
    dev = torch.device("cuda:1")
    torch.manual_seed(9)
    q = torch.randn(2, 300, 4, 128, device=dev, dtype=torch.bfloat16, requires_grad=True)
    k = torch.randn(2, 300, 2, 128, device=dev, dtype=torch.bfloat16, requires_grad=True)
    v = torch.randn(2, 300, 2, 128, device=dev, dtype=torch.bfloat16, requires_grad=True)
    do = torch.randn(2, 300, 4, 128, device=dev, dtype=torch.bfloat16)
    out = api.flash_attn_func(q, k, v, causal=True)
    dq, dk, dv = torch.autograd.grad(out, (q, k, v), do)
    assert torch.cuda.current_device() == 0 and out.device == dev and dq.device == dev
    ref, _ = ao.flash_attn_func_ref(q.detach(), k.detach(), v.detach(), causal=True)
    assert (out.double().cpu() - ref).abs().max().item() <= 2e-2
    rq, rk, rv, _ = ao.flash_attn_bwd_ref(do, q.detach(), k.detach(), v.detach(), causal=True)
    for got, r in ((dq, rq), (dk, rk), (dv, rv)):
        assert (got.double().cpu() - r).abs().max().item() <= 4e-2 * max(1.0, r.abs().max().item())
    # the same call on device 0 afterwards: per-device one-time set-up must not have been skipped for it
    q0, k0, v0 = (t.detach().to("cuda:0") for t in (q, k, v))
    out0 = api.flash_attn_func(q0, k0, v0, causal=True)
    assert torch.equal(out0.cpu(), out.detach().cpu())
    # decode with a kv-cache on device 1
    kc = torch.randn(2, 512, 2, 128, device=dev, dtype=torch.bfloat16)
    vc = torch.randn(2, 512, 2, 128, device=dev, dtype=torch.bfloat16)
    lens = torch.tensor([100, 511], dtype=torch.int32, device=dev)
    qd = torch.randn(2, 1, 4, 128, device=dev, dtype=torch.bfloat16)
    o = api.flash_attn_with_kvcache(qd, kc, vc, cache_seqlens=lens, causal=True)
    r, _, _, _ = ao.flash_attn_with_kvcache_ref(qd, kc, vc, cache_seqlens=lens, causal=True)
    assert (o.double().cpu() - r).abs().max().item() <= 2e-2
