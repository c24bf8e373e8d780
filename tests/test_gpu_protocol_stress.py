"""Protocol stress: the forward's warp-role hand-shakes under adversarial timing.

libfa_b200_jitter.so is the product source compiled with -DFA_JITTER (csrc/ptx_sm100.cuh): one time in four, every
mbarrier wait and arrival is preceded by a pseudo-random busy-wait of up to ~8k clocks, per warp and per call, which skews
the roles of a CTA against each other far beyond what the hardware does on its own. A hand-off that only works thanks to
"natural" timing then hangs (the in-kernel watchdog turns that into cudaErrorLaunchFailure) or fails parity. Round 2's
first shared-S forward had such a hole (a correction warp a tile ahead of a sibling); it passed every
other test. The cases run in a child process: a trapped kernel kills its CUDA context."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "flash-attention-v100_b200", "lib", "libfa_b200_jitter.so")

CHILD = r'''
import json, os, sys
import torch
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "flash-attention-v100_b200"))
from flash_attn_v100 import flash_attn_func, flash_attn_varlen_func
from oracle.attention_oracle import flash_attn_func_ref
torch.manual_seed(7)
res = []
def dense(name, B, Sq, Sk, H, Hk, D, causal, window=(-1, -1), dt=torch.bfloat16, reps=3):
    q = torch.randn(B, Sq, H, D, device="cuda", dtype=dt)
    k = torch.randn(B, Sk, Hk, D, device="cuda", dtype=dt)
    v = torch.randn(B, Sk, Hk, D, device="cuda", dtype=dt)
    for _ in range(reps):
        out = flash_attn_func(q, k, v, causal=causal, window_size=window)
    torch.cuda.synchronize()
    ref, _ = flash_attn_func_ref(q[:1], k[:1], v[:1], causal=causal, window_size=window)
    res.append({"name": name, "err": float((out[:1].double().cpu() - ref).abs().max())})
# few items per CTA, many items per CTA (persistent scheduler, item boundaries), stages with different tile ranges,
# ragged tails, GQA, head dims 64 / 128 / 256
dense("causal_512", 2, 512, 512, 4, 4, 128, True)
dense("full_many_items", 4, 2048, 2048, 48, 48, 128, False, reps=2)
dense("causal_many_items_gqa", 2, 4096, 4096, 64, 8, 128, True, reps=2)
dense("ragged_sq_ne_sk", 2, 333, 777, 4, 2, 128, True)
dense("window", 1, 1536, 1536, 8, 8, 128, True, window=(300, 0))
dense("d64", 2, 1024, 1024, 16, 16, 64, True, dt=torch.float16)
dense("d256", 1, 1024, 1024, 8, 8, 256, True)
dense("d256_full_many_items", 2, 1536, 1536, 24, 6, 256, False, reps=2)
dense("d192", 1, 1024, 1024, 8, 8, 192, True)
lens = [5, 333, 128, 1, 640, 257, 1900, 77]
cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
T = sum(lens)
q = torch.randn(T, 8, 128, device="cuda", dtype=torch.bfloat16)
k = torch.randn(T, 4, 128, device="cuda", dtype=torch.bfloat16)
v = torch.randn(T, 4, 128, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    out = flash_attn_varlen_func(q, k, v, cu, cu, max(lens), max(lens), causal=True)
torch.cuda.synchronize()
s0, s1 = int(cu[1]), int(cu[2])
ref, _ = flash_attn_func_ref(q[None, s0:s1], k[None, s0:s1], v[None, s0:s1], causal=True)
res.append({"name": "varlen", "err": float((out[s0:s1].double().cpu() - ref[0]).abs().max())})
print("RESULT " + json.dumps(res))
''' % (ROOT, ROOT)


@pytest.mark.gpu
def test_forward_hand_shakes_survive_jitter():
    if not os.path.exists(LIB):
        pytest.skip("libfa_b200_jitter.so not built (FA_BUILD_JITTER=1 bash csrc/build.sh, or __graft_entry__.build())")
    env = dict(os.environ, FA_B200_LIB=LIB)
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    line = [x for x in out.stdout.splitlines() if x.startswith("RESULT ")]
    assert out.returncode == 0 and line, f"jitter build failed or hung:\n{out.stdout[-2000:]}\n{out.stderr[-3000:]}"
    for r in json.loads(line[-1][7:]):
        assert r["err"] < 2e-2, r
