#!/usr/bin/env python3
"""Where does the host time of one small forward call go (BASELINE config 1: fp16 B=2 H=8 S=512 D=64)?
Times, per call and without waiting for the GPU (the queue never fills: 1.07 GFLOP kernels): the public API, the
operator layer, and the bare C-ABI call with a pre-filled parameter block."""
import ctypes
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v100_b200"))
import flash_attn_v100_cuda as op  # noqa: E402
from flash_attn_v100 import flash_attn_func, flash_attn_with_kvcache  # noqa: E402


def per_call(fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6


q, k, v = (torch.randn(2, 512, 8, 64, device="cuda", dtype=torch.float16) for _ in range(3))
print("flash_attn_func             host %.1f us/call  (wall incl. drain %.1f)" % per_call(lambda: flash_attn_func(q, k, v)))
qt, kt, vt = (t.permute(0, 2, 1, 3) for t in (q, k, v))
print("op.fwd                      host %.1f us/call  (wall %.1f)" % per_call(lambda: op.fwd(qt, kt, vt, None, None, 0.0, 0.125, False, -1, -1, 0.0, False, None)))
out = torch.empty_like(q)
print("op.fwd with out=            host %.1f us/call  (wall %.1f)" % per_call(lambda: op.fwd(qt, kt, vt, out.permute(0, 2, 1, 3), None, 0.0, 0.125, False, -1, -1, 0.0, False, None)))
# bare C call
lib = op.load_library()
p = op.FaB200Params()
lse = torch.empty(2, 8, 512, device="cuda", dtype=torch.float32)
p.struct_bytes = ctypes.sizeof(op.FaB200Params)
p.dtype, p.device = 0, 0
p.batch, p.seqlen_q, p.seqlen_k, p.num_heads, p.num_heads_k, p.head_dim = 2, 512, 512, 8, 8, 64
p.q, p.k, p.v, p.out, p.lse = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr()
for name, t in (("q", qt), ("k", kt), ("v", vt), ("o", out.permute(0, 2, 1, 3))):
    setattr(p, f"{name}_stride_b", t.stride(0)); setattr(p, f"{name}_stride_h", t.stride(1)); setattr(p, f"{name}_stride_s", t.stride(2))
p.softmax_scale = 0.125
p.window_left = p.window_right = -1
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ref = ctypes.byref(p)
print("fa_b200_fwd (C ABI only)    host %.1f us/call  (wall %.1f)" % per_call(lambda: lib.fa_b200_fwd(ref, stream)))
print("torch.empty x2              host %.1f us/call" % per_call(lambda: (torch.empty_like(q), torch.empty(2, 8, 512, device="cuda", dtype=torch.float32)))[0])
print("FaB200Params() + 40 sets    host %.1f us/call" % per_call(lambda: [setattr(op.FaB200Params(), "batch", 1) for _ in range(1)] and [setattr(p, "batch", 2) for _ in range(40)])[0])
# decode B=1
dt = torch.bfloat16
kc = torch.randn(32, 256, 8, 128, device="cuda", dtype=dt); vc = torch.randn_like(kc)
bt = torch.arange(32, dtype=torch.int32, device="cuda").view(1, 32)
lens = torch.full((1,), 8191, dtype=torch.int32, device="cuda")
qd = torch.randn(1, 1, 32, 128, device="cuda", dtype=dt); kn = torch.randn(1, 1, 8, 128, device="cuda", dtype=dt); vn = torch.randn_like(kn)
ang = torch.rand(8192, 64, device="cuda"); cos, sin = ang.cos().to(dt), ang.sin().to(dt)
print("flash_attn_with_kvcache B=1 host %.1f us/call  (wall %.1f)" % per_call(lambda: flash_attn_with_kvcache(qd, kc, vc, kn, vn, rotary_cos=cos, rotary_sin=sin, cache_seqlens=lens, block_table=bt, causal=True, rotary_interleaved=False), 1000))
