"""ctypes loaders for the oracle's native pieces.  TEST INFRASTRUCTURE ONLY (see attention_oracle.py).

  liboracle_c.so               oracle/attn_oracle.c, the C restatement ("port")
  _ref/libref_cpu_attention.so the reference's own cpu_attention compiled from /root/reference
                               (utils/sass/mma_swizzle/forward_kernel.cu:346-370) by oracle/Makefile
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = ctypes.POINTER(ctypes.c_float)


def build(ref_root: str = "/root/reference") -> None:
    """Compile the C restatement and, when the reference tree is present, oracle/_ref."""
    subprocess.run(["make", "-C", _HERE, "c"], check=True, capture_output=True)
    if os.path.isdir(ref_root):
        subprocess.run(["make", "-C", _HERE, "ref", f"REF_ROOT={ref_root}"], check=True, capture_output=True)


def _ptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(_f32p) if a is not None else None


def load_c_oracle() -> ctypes.CDLL:
    path = os.path.join(_HERE, "liboracle_c.so")
    if not os.path.exists(path):
        build()
    lib = ctypes.CDLL(path)
    lib.oracle_attention.restype = None
    lib.oracle_attention.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p] + [ctypes.c_int] * 5 + [
        ctypes.c_float, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_float, ctypes.c_int]
    lib.oracle_rope.restype = None
    lib.oracle_rope.argtypes = [_f32p, _f32p, _f32p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return lib


def c_oracle_attention(q: np.ndarray, k: np.ndarray, v: np.ndarray, scale: float, wl: int = -1, wr: int = -1,
                       slopes: Optional[np.ndarray] = None, softcap: float = 0.0, threads: int = 0):
    """q:[Sq,H,D] k,v:[Sk,Hk,D] fp32 -> (out [Sq,H,D], lse [H,Sq]) through liboracle_c.so."""
    lib = load_c_oracle()
    q, k, v = (np.ascontiguousarray(t, dtype=np.float32) for t in (q, k, v))
    Sq, H, D = q.shape
    Sk, Hk, _ = k.shape
    out = np.empty((Sq, H, D), dtype=np.float32)
    lse = np.empty((H, Sq), dtype=np.float32)
    sl = np.ascontiguousarray(slopes, dtype=np.float32) if slopes is not None else None
    lib.oracle_attention(_ptr(q), _ptr(k), _ptr(v), _ptr(out), _ptr(lse), Sq, Sk, H, Hk, D, scale, wl, wr,
                         _ptr(sl), softcap, threads or (os.cpu_count() or 1))
    return out, lse


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_cpu_attention.so"))


def load_ref() -> ctypes.CDLL:
    lib = ctypes.CDLL(os.path.join(_HERE, "_ref", "libref_cpu_attention.so"))
    lib.ref_cpu_attention.restype = None
    lib.ref_cpu_attention.argtypes = [_f32p, _f32p, _f32p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_float, ctypes.c_int]
    lib.ref_cpu_attention_batched.restype = None
    lib.ref_cpu_attention_batched.argtypes = [_f32p, _f32p, _f32p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int]
    return lib


def ref_cpu_attention(q: np.ndarray, k: np.ndarray, v: np.ndarray, scale: float, causal: bool,
                      threads: int = 0) -> np.ndarray:
    """The reference's cpu_attention on `heads` independent problems. q:[heads,M,D] k,v:[heads,N,D]."""
    lib = load_ref()
    q, k, v = (np.ascontiguousarray(t, dtype=np.float32) for t in (q, k, v))
    heads, M, D = q.shape
    N = k.shape[1]
    out = np.empty_like(q)
    lib.ref_cpu_attention_batched(_ptr(q), _ptr(k), _ptr(v), _ptr(out), heads, M, N, D, scale, int(causal),
                                  threads or (os.cpu_count() or 1))
    return out
