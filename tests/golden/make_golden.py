#!/usr/bin/env python3
"""Generates the committed golden fixtures from the REFERENCE itself.  Run in the dev container
(needs /root/reference and oracle/_ref built by `make -C oracle ref`):

    python tests/golden/make_golden.py

1. ref_cpu_attention_kat.npz -- the reference's known-answer case
   (utils/sass/mma_swizzle/forward_kernel.cu:394-407, :439): D=128, causal, M=N=128, scale 0.125,
   inputs from glibc srand(42)/rand(); output computed by the reference's own `cpu_attention`
   (:346-370) compiled from the reference sources (oracle/_ref).
2. ref_mha_forward_*.npz -- the reference's test oracle `ref_mha_forward` (test.py:18-34), executed
   verbatim (the function's source is extracted from /root/reference/test.py with `ast`, because
   importing test.py exits when the Volta extension is missing, test.py:12-16) on inputs drawn the
   way test.py draws them (seed 421, randn fp16, [B,H,M,D], test.py:151-157; CPU generator here).
"""
import ast
import ctypes
import ctypes.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import native  # noqa: E402
from oracle.attention_oracle import c_rand_uniform_pm1  # noqa: E402

REF = "/root/reference"


def kat():
    libc = ctypes.CDLL(ctypes.util.find_library("c"))
    libc.srand(42)
    M = N = D = 128
    q = c_rand_uniform_pm1(M * D).reshape(1, M, D)
    k = c_rand_uniform_pm1(N * D).reshape(1, N, D)
    v = c_rand_uniform_pm1(N * D).reshape(1, N, D)
    out = native.ref_cpu_attention(q, k, v, 0.125, True, threads=1)
    np.savez_compressed(os.path.join(HERE, "ref_cpu_attention_kat.npz"), q=q[0], k=k[0], v=v[0], out=out[0],
                        scale=np.float32(0.125), causal=np.int32(1))
    print("kat: out[0,:4] =", out[0, 0, :4])


def load_ref_fn(name):
    src = open(os.path.join(REF, "test.py")).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "reference/test.py", "exec"), ns)
    return ns[name]


def load_ref_mha_forward():
    return load_ref_fn("ref_mha_forward")


def mha_backward_fixtures():
    """3. ref_mha_backward_*.npz -- the reference's backward oracle `ref_mha_backward` (test.py:36-61),
    executed verbatim in fp32 (upcast=True) on seed-421 fp16 inputs and an fp16 dO drawn after them
    (test.py:151-158)."""
    ref_mha_backward = load_ref_fn("ref_mha_backward")
    for (B, H, M, N, D) in [(1, 1, 64, 64, 64), (1, 2, 128, 128, 128), (1, 2, 256, 256, 64), (1, 1, 32, 32, 32),
                            (1, 1, 256, 256, 256)]:  # the last two: test.py:117, :120
        for causal in (False, True):
            torch.manual_seed(421)
            q = torch.randn(B, H, M, D, dtype=torch.float16)
            k = torch.randn(B, H, N, D, dtype=torch.float16)
            v = torch.randn(B, H, N, D, dtype=torch.float16)
            do = torch.randn(B, H, M, D, dtype=torch.float16)
            scale = 1.0 / (D ** 0.5)
            dq, dk, dv = ref_mha_backward(q.float(), k.float(), v.float(), do.float(), scale=scale, causal=causal, upcast=True)
            name = f"ref_mha_backward_B{B}_H{H}_M{M}_N{N}_D{D}_{'causal' if causal else 'full'}.npz"
            np.savez_compressed(os.path.join(HERE, name), q=q.numpy(), k=k.numpy(), v=v.numpy(), do=do.numpy(),
                                dq=dq.numpy(), dk=dk.numpy(), dv=dv.numpy(), scale=np.float32(scale), causal=np.int32(causal))
            print(name, dq.abs().max().item(), dk.abs().max().item(), dv.abs().max().item())


def mha_fixtures():
    ref_mha_forward = load_ref_mha_forward()
    for (B, H, M, N, D) in [(1, 1, 16, 16, 16), (1, 1, 64, 64, 64), (1, 2, 128, 128, 128), (1, 2, 256, 256, 64),
                            (1, 1, 32, 32, 32), (1, 1, 256, 256, 256)]:  # the last two: test.py:117, :120
        for causal in (False, True):
            torch.manual_seed(421)
            q = torch.randn(B, H, M, D, dtype=torch.float16)
            k = torch.randn(B, H, N, D, dtype=torch.float16)
            v = torch.randn(B, H, N, D, dtype=torch.float16)
            scale = 1.0 / (D ** 0.5)
            o32 = ref_mha_forward(q.float(), k.float(), v.float(), scale=scale, causal=causal, upcast=True)
            name = f"ref_mha_forward_B{B}_H{H}_M{M}_N{N}_D{D}_{'causal' if causal else 'full'}.npz"
            np.savez_compressed(os.path.join(HERE, name), q=q.numpy(), k=k.numpy(), v=v.numpy(), out=o32.numpy(),
                                scale=np.float32(scale), causal=np.int32(causal))
            print(name, o32.abs().max().item())


if __name__ == "__main__":
    kat()
    mha_fixtures()
    mha_backward_fixtures()
