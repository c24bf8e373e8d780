#!/usr/bin/env python3
"""Sustained (power-capped) A/B of library variants on BASELINE config 2: each library runs the forward back to back
for SECONDS seconds (default 2), in separate processes, alternating, ROUNDS times. The burst numbers of tests/gpu_quick.py
are taken over ~30 ms, before the GPU reaches its power limit; this is the seconds-scale figure.
    python tools/sustained_ab.py lib/libfa_b200.so lib/libfa_b200_x.so ..."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time, json
import torch
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "flash-attention-v100_b200"))
prov = os.environ.get("SUST_PROVIDER", "ours")
if prov == "cudnn":  # the library yardsticks, under the same seconds-scale loop
    from torch.nn.attention import SDPBackend, sdpa_kernel
    import torch.nn.functional as F
    def flash_attn_func(q, k, v, causal):
        with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
            return F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=causal).transpose(1, 2)
elif prov == "fa4":
    from vllm.vllm_flash_attn.cute.interface import flash_attn_func as _f4
    def flash_attn_func(q, k, v, causal):
        r = _f4(q, k, v, causal=causal)
        return r[0] if isinstance(r, tuple) else r
else:
    from flash_attn_v100 import flash_attn_func
causal = os.environ.get("SUST_CAUSAL", "1") == "1"
B, S, H, D = 8, 4096, 32, 128
q, k, v = (torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16) for _ in range(3))
for _ in range(5): flash_attn_func(q, k, v, causal=causal)
torch.cuda.synchronize()
secs = float(os.environ.get("SUST_SECONDS", "2"))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 0; t0 = time.time(); e0.record()
while time.time() - t0 < secs:
    for _ in range(50): flash_attn_func(q, k, v, causal=causal)
    n += 50
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"ms": round(ms, 4), "tflops": round(4 * B * H * S * S * D * (0.5 if causal else 1.0) / ms / 1e9, 1), "steps": n}))
''' % (ROOT, ROOT)
libs = sys.argv[1:]  # library paths, or the provider names "cudnn" / "fa4"
for r in range(int(os.environ.get("ROUNDS", "2"))):
    for lib in libs:
        env = dict(os.environ, SUST_PROVIDER=lib) if lib in ("cudnn", "fa4") else dict(os.environ, FA_B200_LIB=os.path.abspath(lib))
        out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        line = [x for x in out.stdout.splitlines() if x.startswith("{")]
        print(os.path.basename(lib), line[-1] if line else out.stderr[-300:], flush=True)
