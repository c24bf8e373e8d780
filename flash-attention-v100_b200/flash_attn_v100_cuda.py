"""Operator layer of the B200 build: the module the Python API calls, with the reference's names.

Mirrors the pybind11 module `flash_attn_v100_cuda` of ai-bond/flash-attention-v100
(reference kernel/fused_mha_api.cpp:17-33; C++ signatures in reference include/mha.h): the same five
callables, the same positional argument order and the same return lists

    fwd(q, k, v, out, alibi_slopes, p_dropout, softmax_scale, is_causal, window_left, window_right,
        softcap, return_softmax, gen)                      -> [out, lse, dmask, rng_state]
    varlen_fwd(q, k, v, out, cu_seqlens_q, cu_seqlens_k, seqused_k, leftpad_k, block_table,
        alibi_slopes, max_seqlen_q, max_seqlen_k, p_dropout, softmax_scale, zero_tensors, is_causal,
        window_left, window_right, softcap, return_softmax, gen, num_splits)
                                                           -> [out, lse, dmask, rng_state]
    fwd_kvcache(q, kcache, vcache, k, v, seqlens_k, rotary_cos, rotary_sin, cache_batch_idx,
        leftpad_k, block_table, alibi_slopes, out, softmax_scale, is_causal, window_left,
        window_right, softcap, is_rotary_interleaved, num_splits) -> [out, lse]
    bwd(dout, q, k, v, out, softmax_lse, dq, dk, dv, alibi_slopes, p_dropout, softmax_scale, is_causal,
        window_left, window_right, softcap, deterministic, gen, rng_state) -> [dq, dk, dv, softmax_d]
    varlen_bwd(dout, q, k, v, out, softmax_lse, dq, dk, dv, cu_seqlens_q, cu_seqlens_k, alibi_slopes,
        max_seqlen_q, max_seqlen_k, p_dropout, softmax_scale, zero_tensors, is_causal, window_left,
        window_right, softcap, deterministic, gen, rng_state)      -> [dq, dk, dv, softmax_d]

Where the reference's wrappers (kernel/fused_mha_forward.cu:301-432, ..._varlen.cu:371-566,
..._kvcache.cu:416-652) validate with TORCH_CHECK and allocate outputs with ATen, this module
validates in Python (same messages where the reference has one), allocates with torch, and hands raw
pointers + strides to the C ABI in libfa_b200.so (include/fa_b200.h) through ctypes. There is no
CPU or PyTorch fallback: if the library is missing or the device is not sm_100 the call raises.

Layout contract kept from the reference: `fwd` takes q,k,v as [B, H, S, D] (include/mha.h:12-15) --
but by *strides*, so a permuted view of a [B, S, H, D] tensor is consumed in place (no copies).
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
def _find_library() -> str:
    """FA_B200_LIB, else the in-tree build (lib/ next to this module), else the installed layout (the library
    travels inside the flash_attn_v100 package: setup.py)."""
    env = os.environ.get("FA_B200_LIB")
    if env:
        return env
    for cand in (os.path.join(_HERE, "lib", "libfa_b200.so"), os.path.join(_HERE, "flash_attn_v100", "lib", "libfa_b200.so")):
        if os.path.exists(cand):
            return cand
    return os.path.join(_HERE, "lib", "libfa_b200.so")


_LIB_PATH = _find_library()

FA_B200_DTYPE_FP16 = 0
FA_B200_DTYPE_BF16 = 1
KIND_DENSE, KIND_VARLEN, KIND_KVCACHE = 0, 1, 2

_i32, _i64, _f32, _ptr = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p


class FaB200Params(ctypes.Structure):
    """ctypes mirror of fa_b200_params_t (include/fa_b200.h); tests/test_abi.py checks the match."""

    _fields_ = [
        ("struct_bytes", _i32), ("dtype", _i32), ("device", _i32), ("reserved0", _i32),
        ("batch", _i32), ("seqlen_q", _i32), ("seqlen_k", _i32), ("num_heads", _i32),
        ("num_heads_k", _i32), ("head_dim", _i32), ("total_q", _i32), ("total_k", _i32),
        ("q", _ptr), ("k", _ptr), ("v", _ptr), ("out", _ptr), ("lse", _ptr),
        ("q_stride_b", _i64), ("q_stride_s", _i64), ("q_stride_h", _i64),
        ("k_stride_b", _i64), ("k_stride_s", _i64), ("k_stride_h", _i64),
        ("v_stride_b", _i64), ("v_stride_s", _i64), ("v_stride_h", _i64),
        ("o_stride_b", _i64), ("o_stride_s", _i64), ("o_stride_h", _i64),
        ("batch_k", _i32), ("reserved1", _i32),
        ("cu_seqlens_q", _ptr), ("cu_seqlens_k", _ptr), ("seqused_k", _ptr),
        ("block_table", _ptr), ("block_table_stride", _i32), ("page_size", _i32),
        ("num_pages", _i32), ("reserved2", _i32),
        ("cache_seqlens", _ptr), ("cache_batch_idx", _ptr), ("cache_leftpad", _ptr),
        ("k_new", _ptr), ("v_new", _ptr),
        ("knew_stride_b", _i64), ("knew_stride_s", _i64), ("knew_stride_h", _i64),
        ("vnew_stride_b", _i64), ("vnew_stride_s", _i64), ("vnew_stride_h", _i64),
        ("seqlen_new", _i32), ("rotary_dim", _i32),
        ("rotary_cos", _ptr), ("rotary_sin", _ptr),
        ("rotary_seqlen", _i32), ("rotary_interleaved", _i32),
        ("alibi_slopes", _ptr), ("alibi_stride_b", _i64),
        ("softmax_scale", _f32), ("softcap", _f32),
        ("is_causal", _i32), ("window_left", _i32), ("window_right", _i32), ("num_splits", _i32),
        ("workspace", _ptr), ("workspace_bytes", _i64),
        ("p_dropout", _f32), ("reserved3", _i32), ("dropout_seed", ctypes.c_uint64), ("dropout_offset", ctypes.c_uint64),
        ("dmask", _ptr),
        ("dout", _ptr), ("dq", _ptr), ("dk", _ptr), ("dv", _ptr), ("softmax_d", _ptr),
        ("do_stride_b", _i64), ("do_stride_s", _i64), ("do_stride_h", _i64),
        ("dq_stride_b", _i64), ("dq_stride_s", _i64), ("dq_stride_h", _i64),
        ("dk_stride_b", _i64), ("dk_stride_s", _i64), ("dk_stride_h", _i64),
        ("dv_stride_b", _i64), ("dv_stride_s", _i64), ("dv_stride_h", _i64),
    ]


EXPORTED_SYMBOLS = (
    "fa_b200_abi_version", "fa_b200_last_error", "fa_b200_workspace_bytes", "fa_b200_fwd",
    "fa_b200_varlen_fwd", "fa_b200_kvcache_fwd", "fa_b200_launch_count", "fa_b200_bwd", "fa_b200_varlen_bwd",
    "fa_b200_init",
)

_lib = None


def load_library() -> ctypes.CDLL:
    """dlopen libfa_b200.so (built by `__graft_entry__.build()` / csrc/build.sh). Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"libfa_b200.so not found at {_LIB_PATH}: build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (there is no fallback path)")
    lib = ctypes.CDLL(_LIB_PATH)
    lib.fa_b200_abi_version.restype = ctypes.c_int
    lib.fa_b200_last_error.restype = ctypes.c_char_p
    lib.fa_b200_launch_count.restype = ctypes.c_int64
    if hasattr(lib, "fa_b200_init"):  # absent from round-1 builds loaded through FA_B200_LIB for A/B runs
        lib.fa_b200_init.restype = ctypes.c_int
        lib.fa_b200_init.argtypes = [ctypes.c_int]
    lib.fa_b200_workspace_bytes.restype = ctypes.c_int64
    lib.fa_b200_workspace_bytes.argtypes = [ctypes.POINTER(FaB200Params), ctypes.c_int]
    for name in ("fa_b200_fwd", "fa_b200_varlen_fwd", "fa_b200_kvcache_fwd", "fa_b200_bwd", "fa_b200_varlen_bwd"):
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.POINTER(FaB200Params), ctypes.c_void_p]
    if lib.fa_b200_abi_version() != 3:
        raise ImportError(f"libfa_b200.so ABI {lib.fa_b200_abi_version()} != 3")
    _lib = lib
    return lib


def launch_count() -> int:
    """CUDA kernels launched by libfa_b200.so so far in this process."""
    return int(load_library().fa_b200_launch_count())


def debug_set_counters(buf: Optional[torch.Tensor]) -> None:
    """Tests only: point the forward kernel's debug counters at `buf` (int64[2] on the GPU; [0] = softmax rows whose
    running maximum crossed the lazy-rescale threshold, [1] = rescales of the O accumulator), or None to switch off."""
    lib = load_library()
    lib.fa_b200_debug_set_counters.restype = ctypes.c_int
    lib.fa_b200_debug_set_counters.argtypes = [ctypes.c_void_p]
    if buf is not None:
        assert buf.is_cuda and buf.dtype == torch.int64 and buf.numel() >= 2 and buf.is_contiguous()
    lib.fa_b200_debug_set_counters(ctypes.c_void_p(buf.data_ptr() if buf is not None else None))


def _check(cond: bool, msg: str) -> None:
    # the reference raises c10::Error (a RuntimeError in Python) from TORCH_CHECK
    if not cond:
        raise RuntimeError(msg)


def _dtype_code(t: torch.Tensor) -> int:
    _check(t.dtype in (torch.float16, torch.bfloat16), "q must be fp16 or bf16")
    return FA_B200_DTYPE_FP16 if t.dtype == torch.float16 else FA_B200_DTYPE_BF16


def _call(fn_name: str, params: FaB200Params, device: torch.device) -> None:
    lib = load_library()
    params.struct_bytes = ctypes.sizeof(FaB200Params)
    # the C layer selects `params.device` itself (and restores the caller's); no torch device guard is needed
    stream = torch.cuda.current_stream(device).cuda_stream
    rc = getattr(lib, fn_name)(ctypes.byref(params), ctypes.c_void_p(stream))
    if rc != 0:
        msg = lib.fa_b200_last_error().decode("utf-8", "replace")
        if rc == -2:
            raise NotImplementedError(msg)
        raise RuntimeError(f"{fn_name} failed ({rc}): {msg}")


def _padded_dim(d: int) -> int:
    """Head dims the C layer takes: any multiple of 8 up to 256 (the reference's own checks,
    kernel/fused_mha_forward.cu:335-336); the kernels' tiles are 64 / 128 / 256 wide and TMA zero-fills the
    columns past `d`, so nothing is padded here."""
    _check(d <= 256, "head dimension must be <= 256")
    _check(d % 8 == 0, "head dimension must be multiple of 8")
    return d


def _pad_last(x: Optional[torch.Tensor], d_to: int) -> Optional[torch.Tensor]:
    if x is None or x.shape[-1] == d_to:
        return x
    return torch.nn.functional.pad(x, [0, d_to - x.shape[-1]])


def _aligned(x: torch.Tensor) -> torch.Tensor:
    """TMA needs a 16-byte aligned base and positive strides that are multiples of 8 elements."""
    # Expanded (stride-0) or negative-stride views cannot be described to TMA: copy them like the reference's
    # .contiguous() does. A dimension of extent 1 may carry any stride (it is never stepped over).
    ok = x.stride(-1) == 1 and x.data_ptr() % 16 == 0 and all(
        n == 1 or (s > 0 and s % 8 == 0) for n, s in zip(x.shape[:-1], x.stride()[:-1]))
    return x if ok else x.contiguous()


def _alibi(p: FaB200Params, alibi_slopes: Optional[torch.Tensor], batch: int, heads: int, keep: list) -> None:
    if alibi_slopes is None:
        return
    s = alibi_slopes
    _check(s.dtype == torch.float32 and s.is_cuda, "alibi_slopes must be fp32 on CUDA")
    _check(s.stride(-1) == 1, "alibi_slopes last dim must be contiguous")
    valid = (s.dim() == 1 and s.shape[0] == heads) or (s.dim() == 2 and tuple(s.shape) == (batch, heads))
    _check(valid, "alibi_slopes must be [H_Q] or [B, H_Q]")
    p.alibi_slopes = s.data_ptr()
    p.alibi_stride_b = s.stride(0) if s.dim() == 2 else 0
    keep.append(s)


def _check_dropout(p_dropout: float, return_softmax: bool, softcap: float) -> None:
    _check(0.0 <= p_dropout < 1.0, "p_dropout must be in [0, 1)")
    if softcap > 0.0:
        _check(p_dropout == 0.0, "Softcapping does not support dropout")
    _check((not return_softmax) or p_dropout > 0.0, "return_softmax requires p_dropout > 0")


def _dropout_state(p: FaB200Params, p_dropout: float, gen_, device: torch.device, batch: int, heads: int) -> torch.Tensor:
    """Seed / offset from the CUDA generator, advanced like the reference does
    (kernel/fused_mha_forward.cu:373-386): offset += B * H * 32. Returns rng_state = [seed, offset] (int64)."""
    if p_dropout <= 0.0:
        return torch.empty((2,), dtype=torch.int64, device=device)  # only meaningful with dropout
    gen = gen_ if gen_ is not None else torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    seed, offset = int(gen.initial_seed()), int(gen.get_offset())
    gen.set_offset(offset + batch * heads * 32)
    seed &= 2 ** 64 - 1
    p.p_dropout, p.dropout_seed, p.dropout_offset = float(p_dropout), seed, offset
    as_i64 = lambda x: x - 2 ** 64 if x >= 2 ** 63 else x  # rng_state is int64 like the reference's
    return torch.tensor([as_i64(seed), as_i64(offset)], dtype=torch.int64).to(device, non_blocking=True)


# ======================================================================================
# dense:  replaces flash_attention_forward (reference kernel/fused_mha_forward.cu:301-432)
# ======================================================================================
def fwd(q, k, v, out_, alibi_slopes_, p_dropout, softmax_scale, is_causal, window_left, window_right,
        softcap, return_softmax, gen_) -> List[torch.Tensor]:
    _check(q.is_cuda and k.is_cuda and v.is_cuda, "Tensors q, k, v must be on CUDA")
    dt = _dtype_code(q)
    _check(k.dtype == q.dtype and v.dtype == q.dtype, "k/v must have the same dtype as q")
    _check(q.stride(-1) == 1 and k.stride(-1) == 1 and v.stride(-1) == 1, "Last dim of q, k, v must be contiguous")
    B, H, M, D = q.shape
    Hk, N = k.shape[1], k.shape[2]
    _check(B > 0, "batch size must be positive")
    _check(H % Hk == 0, "H_Q must be divisible by H_K for GQA/MQA")
    _check_dropout(p_dropout, return_softmax, softcap)
    Dp = _padded_dim(D)

    lse = torch.empty((B, H, M), dtype=torch.float32, device=q.device)
    want_mask = return_softmax and p_dropout > 0.0  # reference :401-406
    # +1 kept / -1 dropped wherever the kernel evaluated the score tile; 0 in tiles it skipped (fully masked)
    dmask = torch.zeros((B, H, M, N), dtype=q.dtype, device=q.device) if want_mask else torch.empty((0,), dtype=q.dtype, device=q.device)
    rng_state = None
    if out_ is not None:
        _check(out_.dtype == q.dtype, "out must have the same dtype as q")
        _check(out_.is_cuda, "out must be on CUDA")
        _check(out_.stride(-1) == 1, "out must have contiguous last dimension")
        _check(out_.shape == q.shape, "out shape must match q shape")
    if N == 0 or M == 0:  # reference :409-413
        out = out_ if out_ is not None else torch.empty_like(q)
        out.zero_()
        lse.fill_(float("-inf"))
        return [out, lse, dmask, torch.empty((2,), dtype=torch.int64, device=q.device)]

    qp, kp, vp = (_aligned(_pad_last(t, Dp)) for t in (q, k, v))
    direct = out_ is not None and Dp == D and _aligned(out_) is out_
    out = out_ if direct else torch.empty_like(qp)
    if not (out.data_ptr() % 16 == 0 and all(s % 8 == 0 for s in out.stride()[:-1])):
        out = torch.empty(qp.shape, dtype=q.dtype, device=q.device)

    p = FaB200Params()
    keep = [qp, kp, vp, out, lse]
    p.dtype, p.device = dt, q.device.index if q.device.index is not None else torch.cuda.current_device()
    p.batch, p.seqlen_q, p.seqlen_k, p.num_heads, p.num_heads_k, p.head_dim = B, M, N, H, Hk, Dp
    p.q, p.k, p.v, p.out, p.lse = qp.data_ptr(), kp.data_ptr(), vp.data_ptr(), out.data_ptr(), lse.data_ptr()
    # tensors are [B, H, S, D]: batch stride 0, head stride 1, row stride 2
    p.q_stride_b, p.q_stride_h, p.q_stride_s = qp.stride(0), qp.stride(1), qp.stride(2)
    p.k_stride_b, p.k_stride_h, p.k_stride_s = kp.stride(0), kp.stride(1), kp.stride(2)
    p.v_stride_b, p.v_stride_h, p.v_stride_s = vp.stride(0), vp.stride(1), vp.stride(2)
    p.o_stride_b, p.o_stride_h, p.o_stride_s = out.stride(0), out.stride(1), out.stride(2)
    _alibi(p, alibi_slopes_, B, H, keep)
    p.softmax_scale, p.softcap = float(softmax_scale), float(softcap)
    p.is_causal, p.window_left, p.window_right = int(bool(is_causal)), int(window_left), int(window_right)
    rng_state = _dropout_state(p, p_dropout, gen_, q.device, B, H)
    if want_mask:
        p.dmask = dmask.data_ptr()
    _call("fa_b200_fwd", p, q.device)

    if not direct:
        res = out[..., :D]
        if out_ is not None:
            out_.copy_(res)
            res = out_
        out = res
    return [out, lse, dmask, rng_state]


# --------------------------------------------------------------------------------------
# Dense forward, lean path for the steady state of an inference loop: same call, same kernel, but everything that
# depends only on shapes / strides / dtypes / options is validated ONCE per distinct signature and kept as a
# pre-filled parameter block; a repeat call allocates `out`, patches four pointers and enters the C ABI.
# (tools/host_overhead.py: the general path above costs ~26 us of Python per call on top of the 6 us C call -- more
# than the 12 us the GPU needs for BASELINE config 1.) Per-thread, so concurrent callers never share a block.
# --------------------------------------------------------------------------------------
import threading  # noqa: E402

_fast_tls = threading.local()
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _fast_entry(q, k, v, softmax_scale, causal, wl, wr, softcap):
    """Validated parameter block for this signature, or False if the lean path does not apply (the general path
    then handles -- or rejects -- the call with its usual messages)."""
    if q.dim() != 4 or k.dim() != 4 or v.shape != k.shape or not (q.is_cuda and k.is_cuda and v.is_cuda):
        return False
    if q.dtype not in (torch.float16, torch.bfloat16) or k.dtype != q.dtype or v.dtype != q.dtype:
        return False
    if k.device != q.device or v.device != q.device:
        return False
    B, M, H, D = q.shape
    N, Hk = k.shape[1], k.shape[2]
    if min(B, M, N, H, Hk) <= 0 or k.shape[0] != B or k.shape[3] != D or H % Hk or D % 8 or D > 256:
        return False
    for t in (q, k, v):
        if t.stride(-1) != 1 or any(n != 1 and (s <= 0 or s % 8) for n, s in zip(t.shape[:-1], t.stride()[:-1])):
            return False
    lib = load_library()
    p = FaB200Params()
    p.struct_bytes = ctypes.sizeof(FaB200Params)
    p.dtype = FA_B200_DTYPE_FP16 if q.dtype == torch.float16 else FA_B200_DTYPE_BF16
    p.device = q.device.index if q.device.index is not None else torch.cuda.current_device()
    p.batch, p.seqlen_q, p.seqlen_k, p.num_heads, p.num_heads_k, p.head_dim = B, M, N, H, Hk, D
    p.q_stride_b, p.q_stride_s, p.q_stride_h = q.stride(0), q.stride(1), q.stride(2)
    p.k_stride_b, p.k_stride_s, p.k_stride_h = k.stride(0), k.stride(1), k.stride(2)
    p.v_stride_b, p.v_stride_s, p.v_stride_h = v.stride(0), v.stride(1), v.stride(2)
    p.o_stride_b, p.o_stride_s, p.o_stride_h = M * H * D, H * D, D  # a fresh contiguous (B, M, H, D) tensor
    p.softmax_scale = float(D ** -0.5 if softmax_scale is None else softmax_scale)
    p.softcap = float(softcap)
    p.is_causal, p.window_left, p.window_right = int(bool(causal)), int(wl), int(wr)
    return (p, ctypes.byref(p), lib.fa_b200_fwd, (B, H, M), (B, M, H, D), p.device)


def fwd_dense_fast(q, k, v, softmax_scale, causal, wl, wr, softcap):
    """flash_attn_func(q, k, v, softmax_scale=, causal=, window_size=, softcap=) on (B, S, H, D) tensors when nothing
    else is asked for (no dropout / ALiBi / returned statistics / autograd). Returns `out`, or None when the call is
    not eligible and must take the general path."""
    cache = getattr(_fast_tls, "cache", None)
    if cache is None:
        cache = _fast_tls.cache = {}
    key = (q.shape, k.shape, v.shape, q.stride(), k.stride(), v.stride(), q.dtype, q.device, softmax_scale, causal, wl, wr, softcap)
    ent = cache.get(key)
    if ent is None:
        if len(cache) > 256:
            cache.clear()
        ent = cache[key] = _fast_entry(q, k, v, softmax_scale, causal, wl, wr, softcap)
    if ent is False:
        return None
    p, ref, fn, lshape, oshape, dev = ent
    qp, kp, vp = q.data_ptr(), k.data_ptr(), v.data_ptr()
    if (qp | kp | vp) & 15:
        return None
    out = torch.empty(oshape, dtype=q.dtype, device=q.device)
    # the LSE is allocated per call even though nobody receives it here: a cached scratch buffer would be baked into
    # captured CUDA graphs and could not be dropped from the cache safely
    lse = torch.empty(lshape, dtype=torch.float32, device=q.device)
    p.q, p.k, p.v, p.out, p.lse = qp, kp, vp, out.data_ptr(), lse.data_ptr()
    stream = _raw_stream(dev) if _raw_stream is not None else torch.cuda.current_stream(q.device).cuda_stream
    rc = fn(ref, stream)
    if rc != 0:
        msg = load_library().fa_b200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"fa_b200_fwd failed ({rc}): {msg}")
    return out


def _sig(t):
    return None if t is None else (t.shape, t.stride(), t.dtype)


def fwd_kvcache_fast(q, kcache, vcache, k_, v_, seqlens_k_, rotary_cos_, rotary_sin_, cache_batch_idx_, leftpad_k_,
                     block_table_, alibi_slopes_, softmax_scale, is_causal, window_left, window_right, softcap,
                     is_rotary_interleaved, num_splits):
    """The lean path of `fwd_kvcache` for a decode loop: the first call with a given signature (shapes, strides,
    dtypes, which optional tensors are present, scalar options) goes through the general path once -- so every check and
    error message is the reference's -- and its filled parameter block is kept; repeat calls allocate out / lse /
    workspace, patch the pointers and enter the C ABI. Returns [out, lse], or None if the signature is not cached
    yet / not eligible (the caller then takes the general path)."""
    cache = getattr(_fast_tls, "kv_cache", None)
    if cache is None:
        cache = _fast_tls.kv_cache = {}
    key = (_sig(q), _sig(kcache), _sig(vcache), _sig(k_), _sig(v_), _sig(seqlens_k_), _sig(rotary_cos_), _sig(rotary_sin_),
           _sig(cache_batch_idx_), _sig(leftpad_k_), _sig(block_table_), _sig(alibi_slopes_), q.device, softmax_scale,
           is_causal, window_left, window_right, softcap, is_rotary_interleaved, num_splits)
    ent = cache.get(key)
    if ent is None:
        return None
    p, ref, fn, ws_fn, ws_bytes, oshape, lshape, dev = ent
    ptrs = [q.data_ptr(), kcache.data_ptr(), vcache.data_ptr()]
    p.q, p.k, p.v = ptrs
    if k_ is not None:
        p.k_new, p.v_new = k_.data_ptr(), v_.data_ptr()
        ptrs += [p.k_new, p.v_new]
    acc = 0
    for x in ptrs:
        acc |= x
    if acc & 15:
        return None
    if seqlens_k_ is not None:
        p.cache_seqlens = seqlens_k_.data_ptr()
    if cache_batch_idx_ is not None:
        p.cache_batch_idx = cache_batch_idx_.data_ptr()
    if leftpad_k_ is not None:
        p.cache_leftpad = leftpad_k_.data_ptr()
    if rotary_cos_ is not None:
        p.rotary_cos, p.rotary_sin = rotary_cos_.data_ptr(), rotary_sin_.data_ptr()
    if block_table_ is not None:
        p.block_table = block_table_.data_ptr()
    if alibi_slopes_ is not None:
        p.alibi_slopes = alibi_slopes_.data_ptr()
    out = torch.empty(oshape, dtype=q.dtype, device=q.device)
    lse = torch.empty(lshape, dtype=torch.float32, device=q.device)
    p.out, p.lse = out.data_ptr(), lse.data_ptr()
    if ws_bytes:
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=q.device)
        p.workspace = ws.data_ptr()
    stream = _raw_stream(dev) if _raw_stream is not None else torch.cuda.current_stream(q.device).cuda_stream
    rc = fn(ref, stream)
    if rc != 0:
        msg = load_library().fa_b200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"fa_b200_kvcache_fwd failed ({rc}): {msg}")
    return [out, lse]


def _remember_kvcache_signature(p, args):
    """Called by the general path after a successful call that used only directly consumable tensors."""
    (q, kcache, vcache, k_, v_, seqlens_k_, rotary_cos_, rotary_sin_, cache_batch_idx_, leftpad_k_, block_table_,
     alibi_slopes_, softmax_scale, is_causal, window_left, window_right, softcap, is_rotary_interleaved, num_splits) = args
    cache = getattr(_fast_tls, "kv_cache", None)
    if cache is None:
        cache = _fast_tls.kv_cache = {}
    if len(cache) > 256:
        cache.clear()
    key = (_sig(q), _sig(kcache), _sig(vcache), _sig(k_), _sig(v_), _sig(seqlens_k_), _sig(rotary_cos_), _sig(rotary_sin_),
           _sig(cache_batch_idx_), _sig(leftpad_k_), _sig(block_table_), _sig(alibi_slopes_), q.device, softmax_scale,
           is_causal, window_left, window_right, softcap, is_rotary_interleaved, num_splits)
    lib = load_library()
    B, Sq, H, D = q.shape
    cache[key] = (p, ctypes.byref(p), lib.fa_b200_kvcache_fwd, None, int(p.workspace_bytes), (B, Sq, H, D), (B, H, Sq), int(p.device))


# ======================================================================================
# varlen:  replaces flash_attention_varlen_forward (reference kernel/fused_mha_forward_varlen.cu:371-566)
# ======================================================================================
def varlen_fwd(q, k, v, out_, cu_seqlens_q, cu_seqlens_k, seqused_k_, leftpad_k_, block_table_, alibi_slopes_,
               max_seqlen_q, max_seqlen_k, p_dropout, softmax_scale, zero_tensors, is_causal, window_left,
               window_right, softcap, return_softmax, gen_, num_splits=0) -> List[torch.Tensor]:
    _check(q.is_cuda and k.is_cuda and v.is_cuda, "Tensors q, k, v must be on CUDA")
    dt = _dtype_code(q)
    _check(k.dtype == q.dtype and v.dtype == q.dtype, "k/v must have the same dtype as q")
    _check(q.stride(-1) == 1 and k.stride(-1) == 1 and v.stride(-1) == 1, "Last dim of q, k, v must be contiguous")
    _check(num_splits <= 1, "num_splits > 1 not supported")
    _check(leftpad_k_ is None, "leftpad_k is not supported by the varlen forward")  # reference ignores it
    for name, t in (("cu_seqlens_q", cu_seqlens_q), ("cu_seqlens_k", cu_seqlens_k)):
        _check(t.dtype == torch.int32 and t.dim() == 1 and t.is_cuda and t.is_contiguous(),
               f"{name} must be a contiguous 1-D int32 CUDA tensor")
    paged = block_table_ is not None
    T, H, D = q.shape
    B = cu_seqlens_q.numel() - 1
    _check(B > 0, "batch size must be positive")
    _check(cu_seqlens_k.numel() == B + 1, "cu_seqlens_k must have batch + 1 entries")
    if paged:
        _check(block_table_.dtype == torch.int32 and block_table_.is_cuda, "block_table must be int32 on CUDA")
        _check(block_table_.stride(-1) == 1, "block_table must have contiguous last dimension")
        _check(k.dim() == 4 and v.dim() == 4, "paged k/v must be [num_blocks, page_block_size, H_K, D]")
        num_pages, page, Hk = k.shape[0], k.shape[1], k.shape[2]
        _check(page % 128 == 0, "page_block_size must be a multiple of 128")  # reference: 256; the C ABI takes any multiple of 128
        _check(block_table_.shape[0] == B, "block_table must have one row per sequence")
    else:
        Hk = k.shape[1]
    _check(H % Hk == 0, "H_Q must be divisible by H_K for GQA/MQA")
    if seqused_k_ is not None:
        _check(seqused_k_.dtype == torch.int32 and seqused_k_.is_cuda and seqused_k_.is_contiguous()
               and seqused_k_.numel() == B, "seqused_k must be a contiguous int32 CUDA tensor of size batch")
    _check_dropout(p_dropout, return_softmax, softcap)
    Dp = _padded_dim(D)

    lse = torch.empty((H, T), dtype=torch.float32, device=q.device)
    want_mask = return_softmax and p_dropout > 0.0  # reference ..._varlen.cu:530-534
    dmask = (torch.zeros((T, H, int(max_seqlen_k)), dtype=q.dtype, device=q.device) if want_mask
             else torch.empty((0,), dtype=q.dtype, device=q.device))
    rng_state = torch.empty((2,), dtype=torch.int64, device=q.device)  # only meaningful with dropout
    if out_ is not None:
        _check(out_.dtype == q.dtype and out_.is_cuda and out_.stride(-1) == 1 and out_.shape == q.shape,
               "out must match q in dtype, device and shape with a contiguous last dimension")
    if T == 0 or max_seqlen_q == 0 or max_seqlen_k == 0:
        # reference ..._varlen.cu:537-545: zero_tensors zeroes out and fills lse with -inf before the early return;
        # without it the reference returns uninitialised buffers -- here they are always defined (out = 0, lse = -inf)
        out = out_ if out_ is not None else torch.empty_like(q)
        out.zero_()
        lse.fill_(float("-inf"))
        return [out, lse, dmask, rng_state]

    qp, kp, vp = (_aligned(_pad_last(t, Dp)) for t in (q, k, v))
    direct = out_ is not None and Dp == D and _aligned(out_) is out_
    out = out_ if direct else torch.empty((T, H, Dp), dtype=q.dtype, device=q.device)
    if zero_tensors:  # reference ..._varlen.cu:537-541 (the kernel overwrites every row of every sequence afterwards)
        out.zero_()
        lse.fill_(float("-inf"))

    p = FaB200Params()
    keep = [qp, kp, vp, out, lse, cu_seqlens_q, cu_seqlens_k]
    p.dtype, p.device = dt, q.device.index if q.device.index is not None else torch.cuda.current_device()
    p.batch, p.seqlen_q, p.seqlen_k, p.num_heads, p.num_heads_k, p.head_dim = B, int(max_seqlen_q), int(max_seqlen_k), H, Hk, Dp
    p.total_q = T
    p.q, p.k, p.v, p.out, p.lse = qp.data_ptr(), kp.data_ptr(), vp.data_ptr(), out.data_ptr(), lse.data_ptr()
    p.q_stride_s, p.q_stride_h = qp.stride(0), qp.stride(1)
    p.o_stride_s, p.o_stride_h = out.stride(0), out.stride(1)
    if paged:
        p.k_stride_b, p.k_stride_s, p.k_stride_h = kp.stride(0), kp.stride(1), kp.stride(2)
        p.v_stride_b, p.v_stride_s, p.v_stride_h = vp.stride(0), vp.stride(1), vp.stride(2)
        p.block_table, p.block_table_stride = block_table_.data_ptr(), block_table_.stride(0)
        p.page_size, p.num_pages = page, num_pages
        keep.append(block_table_)
    else:
        p.total_k = kp.shape[0]
        p.k_stride_s, p.k_stride_h = kp.stride(0), kp.stride(1)
        p.v_stride_s, p.v_stride_h = vp.stride(0), vp.stride(1)
    p.cu_seqlens_q, p.cu_seqlens_k = cu_seqlens_q.data_ptr(), cu_seqlens_k.data_ptr()
    if seqused_k_ is not None:
        p.seqused_k = seqused_k_.data_ptr()
        keep.append(seqused_k_)
    _alibi(p, alibi_slopes_, B, H, keep)
    p.softmax_scale, p.softcap = float(softmax_scale), float(softcap)
    p.is_causal, p.window_left, p.window_right = int(bool(is_causal)), int(window_left), int(window_right)
    p.num_splits = int(num_splits)
    rng_state = _dropout_state(p, p_dropout, gen_, q.device, B, H)
    if want_mask:
        p.dmask = dmask.data_ptr()
    _call("fa_b200_varlen_fwd", p, q.device)

    if not direct:
        res = out[..., :D]
        if out_ is not None:
            out_.copy_(res)
            res = out_
        out = res
    return [out, lse, dmask, rng_state]


# ======================================================================================
# kv-cache:  replaces flash_attention_kvcache (reference kernel/fused_mha_forward_kvcache.cu:416-652)
# ======================================================================================
def fwd_kvcache(q, kcache, vcache, k_, v_, seqlens_k_, rotary_cos_, rotary_sin_, cache_batch_idx_, leftpad_k_,
                block_table_, alibi_slopes_, out_, softmax_scale, is_causal, window_left, window_right, softcap,
                is_rotary_interleaved, num_splits) -> List[torch.Tensor]:
    _fast_args = (q, kcache, vcache, k_, v_, seqlens_k_, rotary_cos_, rotary_sin_, cache_batch_idx_, leftpad_k_,
                  block_table_, alibi_slopes_, softmax_scale, is_causal, window_left, window_right, softcap,
                  is_rotary_interleaved, num_splits)
    if out_ is None:
        res = fwd_kvcache_fast(*_fast_args)  # a signature seen (and fully validated) before
        if res is not None:
            return res
    _check(q.is_cuda and kcache.is_cuda and vcache.is_cuda, "q, kcache, vcache must be on CUDA")
    dt = _dtype_code(q)
    _check(kcache.dtype == q.dtype and vcache.dtype == q.dtype, "kcache/vcache must have the same dtype as q")
    _check(q.stride(-1) == 1 and kcache.stride(-1) == 1 and vcache.stride(-1) == 1, "Last dim must be contiguous")
    B, Sq, H, D = q.shape
    paged = block_table_ is not None
    _padded_dim(D)
    if paged:
        _check(block_table_.dtype == torch.int32 and block_table_.is_cuda, "block_table must be int32 on CUDA")
        _check(block_table_.stride(-1) == 1, "block_table must have contiguous last dimension")
        _check(cache_batch_idx_ is None, "Paged KV cache does not support cache_batch_idx")
        _check(leftpad_k_ is None, "Paged KV cache does not support cache_leftpad")
        num_pages, page, Hk = kcache.shape[0], kcache.shape[1], kcache.shape[2]
        _check(page % 128 == 0, "page_block_size must be a multiple of 128")  # reference: 256; the C ABI takes any multiple of 128
        _check(block_table_.shape[0] == B, "block_table must have one row per sequence")
        capacity = block_table_.shape[1] * page
        batch_c = 0
    else:
        batch_c, capacity, Hk = kcache.shape[0], kcache.shape[1], kcache.shape[2]
        if cache_batch_idx_ is None:
            _check(batch_c == B, "batch size of the cache must match q without cache_batch_idx")
    _check(H % Hk == 0, "H_Q must be divisible by H_K for GQA/MQA")
    _check(vcache.shape == kcache.shape, "kcache and vcache must have the same shape")
    _check(num_splits <= 1, "num_splits > 1 not supported")  # reference :462 (0/1 = library decides)
    if softcap > 0.0:
        _check(window_left < 0 and window_right < 0, "softcap does not support window")
        _check(alibi_slopes_ is None, "softcap does not support alibi")

    def _i32vec(t, name):
        _check(t.dtype == torch.int32 and t.is_cuda and t.is_contiguous() and t.numel() == B,
               f"{name} must be a contiguous int32 CUDA tensor of size batch")
        return t

    p = FaB200Params()
    keep = [q, kcache, vcache]
    seqlen_new = 0
    if k_ is not None or v_ is not None:
        _check(k_ is not None and v_ is not None, "k and v must be provided together")
        _check(seqlens_k_ is not None, "seqlens_k is required when appending k/v")
        _check(k_.dtype == q.dtype and v_.dtype == q.dtype, "k/v must have the same dtype as q")
        _check(k_.shape == v_.shape and k_.shape[0] == B and k_.shape[2] == Hk and k_.shape[3] == D,
               "k/v must be [B, S_new, H_K, D]")
        k_, v_ = _aligned(k_), _aligned(v_)
        seqlen_new = k_.shape[1]
        p.k_new, p.v_new, p.seqlen_new = k_.data_ptr(), v_.data_ptr(), seqlen_new
        p.knew_stride_b, p.knew_stride_s, p.knew_stride_h = k_.stride(0), k_.stride(1), k_.stride(2)
        p.vnew_stride_b, p.vnew_stride_s, p.vnew_stride_h = v_.stride(0), v_.stride(1), v_.stride(2)
        keep += [k_, v_]
    if seqlens_k_ is not None:
        p.cache_seqlens = _i32vec(seqlens_k_, "seqlens_k").data_ptr()
        keep.append(seqlens_k_)
    if cache_batch_idx_ is not None:
        p.cache_batch_idx = _i32vec(cache_batch_idx_, "cache_batch_idx").data_ptr()
        keep.append(cache_batch_idx_)
    if leftpad_k_ is not None:
        p.cache_leftpad = _i32vec(leftpad_k_, "leftpad_k").data_ptr()
        keep.append(leftpad_k_)
    if rotary_cos_ is not None or rotary_sin_ is not None:
        _check(rotary_cos_ is not None and rotary_sin_ is not None, "rotary_cos and rotary_sin must be given together")
        _check(k_ is not None, "rotary requires k/v to append")
        _check(rotary_cos_.dtype == q.dtype and rotary_sin_.dtype == q.dtype, "rotary cos/sin must have the dtype of q")
        _check(rotary_cos_.is_contiguous() and rotary_sin_.is_contiguous(), "rotary cos/sin must be contiguous")
        _check(rotary_cos_.shape == rotary_sin_.shape and rotary_cos_.dim() == 2, "rotary cos/sin must be [seqlen_ro, rotary_dim/2]")
        rotary_dim = 2 * rotary_cos_.shape[1]
        _check(rotary_dim <= D, "rotary_dim must be <= head_dim")
        _check(rotary_dim % 16 == 0, "rotary_dim must be multiple of 16")
        _check(rotary_cos_.shape[0] >= capacity, "rotary seqlen must cover the cache length")
        p.rotary_cos, p.rotary_sin = rotary_cos_.data_ptr(), rotary_sin_.data_ptr()
        p.rotary_dim, p.rotary_seqlen = rotary_dim, rotary_cos_.shape[0]
        p.rotary_interleaved = int(bool(is_rotary_interleaved))
        keep += [rotary_cos_, rotary_sin_]

    qa = _aligned(q)
    if out_ is not None:
        _check(out_.dtype == q.dtype and out_.is_cuda and out_.stride(-1) == 1 and out_.shape == q.shape,
               "out must match q in dtype, device and shape with a contiguous last dimension")
    direct = out_ is not None and _aligned(out_) is out_
    out = out_ if direct else torch.empty((B, Sq, H, D), dtype=q.dtype, device=q.device)
    lse = torch.empty((B, H, Sq), dtype=torch.float32, device=q.device)
    keep += [qa, out, lse]

    p.dtype, p.device = dt, q.device.index if q.device.index is not None else torch.cuda.current_device()
    p.batch, p.seqlen_q, p.seqlen_k, p.num_heads, p.num_heads_k, p.head_dim = B, Sq, capacity, H, Hk, D
    p.batch_k = batch_c
    p.q, p.k, p.v, p.out, p.lse = qa.data_ptr(), kcache.data_ptr(), vcache.data_ptr(), out.data_ptr(), lse.data_ptr()
    p.q_stride_b, p.q_stride_s, p.q_stride_h = qa.stride(0), qa.stride(1), qa.stride(2)
    p.o_stride_b, p.o_stride_s, p.o_stride_h = out.stride(0), out.stride(1), out.stride(2)
    p.k_stride_b, p.k_stride_s, p.k_stride_h = kcache.stride(0), kcache.stride(1), kcache.stride(2)
    p.v_stride_b, p.v_stride_s, p.v_stride_h = vcache.stride(0), vcache.stride(1), vcache.stride(2)
    if paged:
        p.block_table, p.block_table_stride = block_table_.data_ptr(), block_table_.stride(0)
        p.page_size, p.num_pages = page, num_pages
        keep.append(block_table_)
    _alibi(p, alibi_slopes_, B, H, keep)
    p.softmax_scale, p.softcap = float(softmax_scale), float(softcap)
    p.is_causal, p.window_left, p.window_right = int(bool(is_causal)), int(window_left), int(window_right)
    p.num_splits = int(num_splits)

    lib = load_library()
    p.struct_bytes = ctypes.sizeof(FaB200Params)
    ws_bytes = int(lib.fa_b200_workspace_bytes(ctypes.byref(p), KIND_KVCACHE))
    if ws_bytes > 0:
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=q.device)
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws_bytes
        keep.append(ws)
    _call("fa_b200_kvcache_fwd", p, q.device)
    if not direct and out_ is not None:
        out_.copy_(out)
        out = out_
    if out_ is None and qa is q and (k_ is None or (k_ is _fast_args[3] and v_ is _fast_args[4])):
        # every tensor was consumed in place (no .contiguous() copies): repeat calls with this signature can skip
        # the validation and fill-in above
        _remember_kvcache_signature(p, _fast_args)
    return [out, lse]


def _bwd_dropout(p: FaB200Params, p_dropout: float, softcap: float, rng_state_) -> None:
    _check(0.0 <= p_dropout < 1.0, "p_dropout must be in [0, 1)")
    if softcap > 0.0:
        _check(p_dropout == 0.0, "Softcapping does not support dropout")
    if p_dropout > 0.0:  # reference kernel/fused_mha_backward.cu:659-666
        _check(rng_state_ is not None and rng_state_.numel() == 2, "rng_state required when p_dropout > 0")
        seed, offset = (int(x) & (2 ** 64 - 1) for x in rng_state_.tolist())
        p.p_dropout, p.dropout_seed, p.dropout_offset = float(p_dropout), seed, offset


def _grad_out(given: Optional[torch.Tensor], like: torch.Tensor, name: str, ref: torch.Tensor) -> torch.Tensor:
    """A gradient buffer the kernel can write: the caller's tensor if it is TMA/vector-store friendly."""
    if given is not None:
        _check(given.dtype == ref.dtype, f"{name} must have the same dtype as q")
        _check(given.is_cuda, f"{name} must be on CUDA")
        _check(given.stride(-1) == 1, f"{name} must have contiguous last dimension")
        _check(given.shape == ref.shape, f"{name} shape must match {'q' if name == 'dq' else 'k'} shape")
        if given.shape == like.shape and _aligned(given) is given:
            return given
    return torch.empty_like(like)


def _finish_grad(buf: torch.Tensor, given: Optional[torch.Tensor], d: int) -> torch.Tensor:
    res = buf[..., :d] if buf.shape[-1] != d else buf
    if given is not None and given is not buf:
        given.copy_(res)
        return given
    return res


# ======================================================================================
# dense backward:  replaces flash_attention_backward (reference kernel/fused_mha_backward.cu:590-721)
# ======================================================================================
def bwd(dout, q, k, v, out, softmax_lse, dq_, dk_, dv_, alibi_slopes_, p_dropout, softmax_scale, is_causal,
        window_left, window_right, softcap, deterministic, gen_, rng_state_) -> List[torch.Tensor]:
    """Tensors are [B, H, S, D] by strides, like `fwd`. Always deterministic (no atomics), so the
    `deterministic` flag the reference rejects (:603) is accepted and has nothing to switch."""
    _check(q.is_cuda and k.is_cuda and v.is_cuda, "Tensors q, k, v must be on CUDA")
    _check(out.is_cuda and dout.is_cuda and softmax_lse.is_cuda, "out, dout, softmax_lse must be on CUDA")
    dt = _dtype_code(q)
    _check(k.dtype == q.dtype and v.dtype == q.dtype, "k/v must have the same dtype as q")
    _check(out.dtype == q.dtype, "out must have the same dtype as q")
    _check(dout.dtype == q.dtype, "dout must have the same dtype as q")
    _check(softmax_lse.dtype == torch.float32, "softmax_lse must be fp32")
    _check(q.stride(-1) == 1 and k.stride(-1) == 1 and v.stride(-1) == 1, "Last dim of q, k, v must be contiguous")
    _check(out.stride(-1) == 1 and dout.stride(-1) == 1, "Last dim of out, dout must be contiguous")
    B, H, M, D = q.shape
    Hk, N = k.shape[1], k.shape[2]
    _check(B > 0, "batch size must be positive")
    _check(H % Hk == 0, "H_Q must be divisible by H_K for GQA/MQA")
    _check(out.shape == q.shape and dout.shape == q.shape, "out and dout must have the shape of q")
    _check(tuple(softmax_lse.shape) == (B, H, M), "softmax_lse must be [B, H_Q, M]")
    Dp = _padded_dim(D)
    p = FaB200Params()
    _bwd_dropout(p, p_dropout, softcap, rng_state_)

    softmax_d = torch.empty((B, H, M), dtype=torch.float32, device=q.device)
    if M == 0 or N == 0:  # reference :696-701
        grads = [t if t is not None else torch.empty_like(r) for t, r in ((dq_, q), (dk_, k), (dv_, v))]
        for t in grads:
            t.zero_()
        softmax_d.zero_()
        return grads + [softmax_d]

    qp, kp, vp, op, dop = (_aligned(_pad_last(t, Dp)) for t in (q, k, v, out, dout))
    lse = softmax_lse.contiguous()
    dq, dk, dv = _grad_out(dq_, qp, "dq", q), _grad_out(dk_, kp, "dk", k), _grad_out(dv_, vp, "dv", k)
    keep = [qp, kp, vp, op, dop, lse, dq, dk, dv, softmax_d]
    p.dtype, p.device = dt, q.device.index if q.device.index is not None else torch.cuda.current_device()
    p.batch, p.seqlen_q, p.seqlen_k, p.num_heads, p.num_heads_k, p.head_dim = B, M, N, H, Hk, Dp
    p.q, p.k, p.v, p.out, p.lse = qp.data_ptr(), kp.data_ptr(), vp.data_ptr(), op.data_ptr(), lse.data_ptr()
    p.dout, p.dq, p.dk, p.dv, p.softmax_d = dop.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), softmax_d.data_ptr()
    for name, t in (("q", qp), ("k", kp), ("v", vp), ("o", op), ("do", dop), ("dq", dq), ("dk", dk), ("dv", dv)):
        setattr(p, f"{name}_stride_b", t.stride(0))
        setattr(p, f"{name}_stride_h", t.stride(1))
        setattr(p, f"{name}_stride_s", t.stride(2))
    _alibi(p, alibi_slopes_, B, H, keep)
    p.softmax_scale, p.softcap = float(softmax_scale), float(softcap)
    p.is_causal, p.window_left, p.window_right = int(bool(is_causal)), int(window_left), int(window_right)
    _call("fa_b200_bwd", p, q.device)
    return [_finish_grad(dq, dq_, D), _finish_grad(dk, dk_, D), _finish_grad(dv, dv_, D), softmax_d]


# ======================================================================================
# varlen backward:  replaces flash_attention_varlen_backward (reference kernel/fused_mha_backward_varlen.cu)
# ======================================================================================
def varlen_bwd(dout, q, k, v, out, softmax_lse, dq_, dk_, dv_, cu_seqlens_q, cu_seqlens_k, alibi_slopes_,
               max_seqlen_q, max_seqlen_k, p_dropout, softmax_scale, zero_tensors, is_causal, window_left,
               window_right, softcap, deterministic, gen_, rng_state_) -> List[torch.Tensor]:
    """Packed layouts: q,out,dout,dq (T_q, H, D); k,v,dk,dv (T_k, H_K, D); softmax_lse / softmax_d [H, T_q]."""
    _check(q.is_cuda and k.is_cuda and v.is_cuda, "Tensors q, k, v must be on CUDA")
    _check(out.is_cuda and dout.is_cuda and softmax_lse.is_cuda, "out, dout, softmax_lse must be on CUDA")
    dt = _dtype_code(q)
    _check(k.dtype == q.dtype and v.dtype == q.dtype, "k/v must have the same dtype as q")
    _check(out.dtype == q.dtype and dout.dtype == q.dtype, "out and dout must have the same dtype as q")
    _check(softmax_lse.dtype == torch.float32, "softmax_lse must be fp32")
    _check(q.stride(-1) == 1 and k.stride(-1) == 1 and v.stride(-1) == 1, "Last dim of q, k, v must be contiguous")
    _check(out.stride(-1) == 1 and dout.stride(-1) == 1, "Last dim of out, dout must be contiguous")
    for name, t in (("cu_seqlens_q", cu_seqlens_q), ("cu_seqlens_k", cu_seqlens_k)):
        _check(t.dtype == torch.int32 and t.dim() == 1 and t.is_cuda and t.is_contiguous(),
               f"{name} must be a contiguous 1-D int32 CUDA tensor")
    T, H, D = q.shape
    Tk, Hk = k.shape[0], k.shape[1]
    B = cu_seqlens_q.numel() - 1
    _check(B > 0, "batch size must be positive")
    _check(cu_seqlens_k.numel() == B + 1, "cu_seqlens_k must have batch + 1 entries")
    _check(H % Hk == 0, "H_Q must be divisible by H_K for GQA/MQA")
    _check(out.shape == q.shape and dout.shape == q.shape, "out and dout must have the shape of q")
    _check(tuple(softmax_lse.shape) == (H, T), "softmax_lse must be [H_Q, total_q]")
    Dp = _padded_dim(D)
    p = FaB200Params()
    _bwd_dropout(p, p_dropout, softcap, rng_state_)

    softmax_d = torch.zeros((H, T), dtype=torch.float32, device=q.device)
    if T == 0 or Tk == 0 or max_seqlen_q == 0 or max_seqlen_k == 0:
        grads = [t if t is not None else torch.empty_like(r) for t, r in ((dq_, q), (dk_, k), (dv_, v))]
        for t in grads:
            t.zero_()
        return grads + [softmax_d]

    qp, kp, vp, op, dop = (_aligned(_pad_last(t, Dp)) for t in (q, k, v, out, dout))
    lse = softmax_lse.contiguous()
    dq, dk, dv = _grad_out(dq_, qp, "dq", q), _grad_out(dk_, kp, "dk", k), _grad_out(dv_, vp, "dv", k)
    if zero_tensors:
        dq.zero_(), dk.zero_(), dv.zero_()
    keep = [qp, kp, vp, op, dop, lse, dq, dk, dv, softmax_d, cu_seqlens_q, cu_seqlens_k]
    p.dtype, p.device = dt, q.device.index if q.device.index is not None else torch.cuda.current_device()
    p.batch, p.seqlen_q, p.seqlen_k, p.num_heads, p.num_heads_k, p.head_dim = B, int(max_seqlen_q), int(max_seqlen_k), H, Hk, Dp
    p.total_q, p.total_k = T, Tk
    p.q, p.k, p.v, p.out, p.lse = qp.data_ptr(), kp.data_ptr(), vp.data_ptr(), op.data_ptr(), lse.data_ptr()
    p.dout, p.dq, p.dk, p.dv, p.softmax_d = dop.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), softmax_d.data_ptr()
    for name, t in (("q", qp), ("k", kp), ("v", vp), ("o", op), ("do", dop), ("dq", dq), ("dk", dk), ("dv", dv)):
        setattr(p, f"{name}_stride_s", t.stride(0))
        setattr(p, f"{name}_stride_h", t.stride(1))
    p.cu_seqlens_q, p.cu_seqlens_k = cu_seqlens_q.data_ptr(), cu_seqlens_k.data_ptr()
    _alibi(p, alibi_slopes_, B, H, keep)
    p.softmax_scale, p.softcap = float(softmax_scale), float(softcap)
    p.is_causal, p.window_left, p.window_right = int(bool(is_causal)), int(window_left), int(window_right)
    _call("fa_b200_varlen_bwd", p, q.device)
    return [_finish_grad(dq, dq_, D), _finish_grad(dk, dk_, D), _finish_grad(dv, dv_, D), softmax_d]


# ======================================================================================
# torch.ops.flash_attn_v100.* : the same five operators as dispatcher ops, with the reference's schemas
# (reference kernel/fused_mha_api.cpp:308-358), so torch.compile / custom-op users find them.
# ======================================================================================
_OP_SCHEMAS = {
    "fwd": "(Tensor(a!) q, Tensor k, Tensor v, Tensor? out, Tensor? alibi_slopes, float p_dropout, "
           "float softmax_scale, bool is_causal, int window_left, int window_right, float softcap, "
           "bool return_softmax, Generator? gen) -> Tensor[]",
    "bwd": "(Tensor dout, Tensor q, Tensor k, Tensor v, Tensor out, Tensor softmax_lse, Tensor? dq, Tensor? dk, "
           "Tensor? dv, Tensor? alibi_slopes, float p_dropout, float softmax_scale, bool is_causal, int window_left, "
           "int window_right, float softcap, bool deterministic, Generator? gen, Tensor? rng_state) -> Tensor[]",
    "varlen_fwd": "(Tensor(a!) q, Tensor k, Tensor v, Tensor? out, Tensor cu_seqlens_q, Tensor cu_seqlens_k, "
                  "Tensor? seqused_k, Tensor? leftpad_k, Tensor? block_table, Tensor? alibi_slopes, int max_seqlen_q, "
                  "int max_seqlen_k, float p_dropout, float softmax_scale, bool zero_tensors, bool is_causal, "
                  "int window_left, int window_right, float softcap, bool return_softmax, Generator? gen, "
                  "int num_splits) -> Tensor[]",
    "varlen_bwd": "(Tensor dout, Tensor q, Tensor k, Tensor v, Tensor out, Tensor softmax_lse, Tensor? dq, Tensor? dk, "
                  "Tensor? dv, Tensor cu_seqlens_q, Tensor cu_seqlens_k, Tensor? alibi_slopes, int max_seqlen_q, "
                  "int max_seqlen_k, float p_dropout, float softmax_scale, bool zero_tensors, bool is_causal, "
                  "int window_left, int window_right, float softcap, bool deterministic, Generator? gen, "
                  "Tensor? rng_state) -> Tensor[]",
    "fwd_kvcache": "(Tensor(a!) q, Tensor kcache, Tensor vcache, Tensor? k, Tensor? v, Tensor? seqlens_k, "
                   "Tensor? rotary_cos, Tensor? rotary_sin, Tensor? cache_batch_idx, Tensor? leftpad_k, "
                   "Tensor? block_table, Tensor? alibi_slopes, Tensor? out, float softmax_scale, bool is_causal, "
                   "int window_left, int window_right, float softcap, bool is_rotary_interleaved, "
                   "int num_splits) -> Tensor[]",
}
_ops_registered = False


def register_torch_ops() -> bool:
    """Define torch.ops.flash_attn_v100.{fwd,bwd,varlen_fwd,varlen_bwd,fwd_kvcache} (CUDA implementations
    only: there is no CPU kernel to dispatch to). Returns False if the namespace is already taken
    (e.g. the reference extension itself is loaded in this process)."""
    global _ops_registered
    if _ops_registered:
        return True
    impls = {"fwd": fwd, "bwd": bwd, "varlen_fwd": varlen_fwd, "varlen_bwd": varlen_bwd, "fwd_kvcache": fwd_kvcache}
    try:
        for name, schema in _OP_SCHEMAS.items():
            torch.library.define(f"flash_attn_v100::{name}", schema)
            torch.library.impl(f"flash_attn_v100::{name}", "CUDA")(impls[name])
    except RuntimeError:
        return False
    _ops_registered = True
    return True


register_torch_ops()


# ======================================================================================
# torch.compile support. The reference registers its five operators with schemas that mark `q` as mutated
# (`Tensor(a!) q`, kernel/fused_mha_api.cpp:308-358) and gives them neither a meta kernel nor an autograd formula, so
# they cannot be traced or differentiated by the dispatcher; `torch.ops.flash_attn_v100.*` above mirrors that surface
# (plus fake kernels, so at least shape propagation works). For `torch.compile(fullgraph=True)` the Python API routes
# through FUNCTIONAL twins in the `fa_b200` namespace: same kernels, no Generator / optional-output arguments, fake
# (meta) implementations, and autograd formulas registered with torch.library.register_autograd.
# ======================================================================================
def _like_strided(x: torch.Tensor, last: Optional[int] = None) -> torch.Tensor:
    return torch.empty_like(x) if last is None or last == x.shape[-1] else x.new_empty((*x.shape[:-1], last))


def _fake_fwd(q, k, v, out_, alibi_slopes_, p_dropout, softmax_scale, is_causal, window_left, window_right, softcap,
              return_softmax, gen_=None):
    B, H, M, _ = q.shape
    N = k.shape[2]
    out = out_ if out_ is not None else torch.empty_like(q)
    lse = q.new_empty((B, H, M), dtype=torch.float32)
    dmask = q.new_empty((B, H, M, N)) if (return_softmax and p_dropout > 0.0) else q.new_empty((0,))
    return [out, lse, dmask, q.new_empty((2,), dtype=torch.int64)]


def _fake_bwd(dout, q, k, v, out, softmax_lse, dq_, dk_, dv_, alibi_slopes_, p_dropout, softmax_scale, is_causal,
              window_left, window_right, softcap, deterministic, gen_=None, rng_state_=None):
    B, H, M, _ = q.shape
    return [dq_ if dq_ is not None else torch.empty_like(q), dk_ if dk_ is not None else torch.empty_like(k),
            dv_ if dv_ is not None else torch.empty_like(v), q.new_empty((B, H, M), dtype=torch.float32)]


def _fake_varlen_fwd(q, k, v, out_, cu_seqlens_q, cu_seqlens_k, seqused_k_, leftpad_k_, block_table_, alibi_slopes_,
                     max_seqlen_q, max_seqlen_k, p_dropout, softmax_scale, zero_tensors, is_causal, window_left,
                     window_right, softcap, return_softmax, gen_=None, num_splits=0):
    T, H, _ = q.shape
    out = out_ if out_ is not None else torch.empty_like(q)
    lse = q.new_empty((H, T), dtype=torch.float32)
    dmask = q.new_empty((T, H, max_seqlen_k)) if (return_softmax and p_dropout > 0.0) else q.new_empty((0,))
    return [out, lse, dmask, q.new_empty((2,), dtype=torch.int64)]


def _fake_varlen_bwd(dout, q, k, v, out, softmax_lse, dq_, dk_, dv_, cu_seqlens_q, cu_seqlens_k, alibi_slopes_,
                     max_seqlen_q, max_seqlen_k, p_dropout, softmax_scale, zero_tensors, is_causal, window_left,
                     window_right, softcap, deterministic, gen_=None, rng_state_=None):
    T, H, _ = q.shape
    return [dq_ if dq_ is not None else torch.empty_like(q), dk_ if dk_ is not None else torch.empty_like(k),
            dv_ if dv_ is not None else torch.empty_like(v), q.new_empty((H, T), dtype=torch.float32)]


def _fake_fwd_kvcache(q, kcache, vcache, k_, v_, seqlens_k_, rotary_cos_, rotary_sin_, cache_batch_idx_, leftpad_k_,
                      block_table_, alibi_slopes_, out_, softmax_scale, is_causal, window_left, window_right, softcap,
                      is_rotary_interleaved, num_splits):
    B, Sq, H, _ = q.shape
    out = out_ if out_ is not None else q.new_empty(q.shape)
    return [out, q.new_empty((B, H, Sq), dtype=torch.float32)]


_FUNCTIONAL_SCHEMAS = {
    "fwd": "(Tensor q, Tensor k, Tensor v, Tensor? alibi_slopes, float p_dropout, float softmax_scale, bool is_causal, "
           "int window_left, int window_right, float softcap, bool return_softmax) -> (Tensor, Tensor, Tensor, Tensor)",
    "bwd": "(Tensor dout, Tensor q, Tensor k, Tensor v, Tensor out, Tensor softmax_lse, Tensor? alibi_slopes, "
           "float p_dropout, float softmax_scale, bool is_causal, int window_left, int window_right, float softcap, "
           "Tensor? rng_state) -> (Tensor, Tensor, Tensor, Tensor)",
    "varlen_fwd": "(Tensor q, Tensor k, Tensor v, Tensor cu_seqlens_q, Tensor cu_seqlens_k, Tensor? block_table, "
                  "Tensor? alibi_slopes, int max_seqlen_q, int max_seqlen_k, float p_dropout, float softmax_scale, "
                  "bool is_causal, int window_left, int window_right, float softcap, bool return_softmax) "
                  "-> (Tensor, Tensor, Tensor, Tensor)",
    "varlen_bwd": "(Tensor dout, Tensor q, Tensor k, Tensor v, Tensor out, Tensor softmax_lse, Tensor cu_seqlens_q, "
                  "Tensor cu_seqlens_k, Tensor? alibi_slopes, int max_seqlen_q, int max_seqlen_k, float p_dropout, "
                  "float softmax_scale, bool is_causal, int window_left, int window_right, float softcap, "
                  "Tensor? rng_state) -> (Tensor, Tensor, Tensor, Tensor)",
    "fwd_kvcache": "(Tensor q, Tensor(a!) kcache, Tensor(b!) vcache, Tensor? k, Tensor? v, Tensor? seqlens_k, "
                   "Tensor? rotary_cos, Tensor? rotary_sin, Tensor? cache_batch_idx, Tensor? leftpad_k, "
                   "Tensor? block_table, Tensor? alibi_slopes, float softmax_scale, bool is_causal, int window_left, "
                   "int window_right, float softcap, bool is_rotary_interleaved, int num_splits) -> (Tensor, Tensor)",
}
_functional_ops_registered = False


def register_functional_ops() -> bool:
    """torch.ops.fa_b200.{fwd,bwd,varlen_fwd,varlen_bwd,fwd_kvcache}: traceable, differentiable twins of the operators
    (see the comment above). Also gives torch.ops.flash_attn_v100.* fake kernels. Idempotent."""
    global _functional_ops_registered
    if _functional_ops_registered:
        return True
    L = torch.library

    def f_fwd(q, k, v, alibi_slopes, p_dropout, softmax_scale, is_causal, window_left, window_right, softcap, return_softmax):
        return tuple(fwd(q, k, v, None, alibi_slopes, p_dropout, softmax_scale, is_causal, window_left, window_right,
                         softcap, return_softmax, None))

    def f_bwd(dout, q, k, v, out, softmax_lse, alibi_slopes, p_dropout, softmax_scale, is_causal, window_left,
              window_right, softcap, rng_state):
        return tuple(bwd(dout, q, k, v, out, softmax_lse, None, None, None, alibi_slopes, p_dropout, softmax_scale,
                         is_causal, window_left, window_right, softcap, False, None, rng_state))

    def f_varlen_fwd(q, k, v, cu_seqlens_q, cu_seqlens_k, block_table, alibi_slopes, max_seqlen_q, max_seqlen_k,
                     p_dropout, softmax_scale, is_causal, window_left, window_right, softcap, return_softmax):
        return tuple(varlen_fwd(q, k, v, None, cu_seqlens_q, cu_seqlens_k, None, None, block_table, alibi_slopes,
                                max_seqlen_q, max_seqlen_k, p_dropout, softmax_scale, False, is_causal, window_left,
                                window_right, softcap, return_softmax, None, 0))

    def f_varlen_bwd(dout, q, k, v, out, softmax_lse, cu_seqlens_q, cu_seqlens_k, alibi_slopes, max_seqlen_q,
                     max_seqlen_k, p_dropout, softmax_scale, is_causal, window_left, window_right, softcap, rng_state):
        return tuple(varlen_bwd(dout, q, k, v, out, softmax_lse, None, None, None, cu_seqlens_q, cu_seqlens_k,
                                alibi_slopes, max_seqlen_q, max_seqlen_k, p_dropout, softmax_scale, False, is_causal,
                                window_left, window_right, softcap, False, None, rng_state))

    def f_kvcache(q, kcache, vcache, k, v, seqlens_k, rotary_cos, rotary_sin, cache_batch_idx, leftpad_k, block_table,
                  alibi_slopes, softmax_scale, is_causal, window_left, window_right, softcap, is_rotary_interleaved,
                  num_splits):
        return tuple(fwd_kvcache(q, kcache, vcache, k, v, seqlens_k, rotary_cos, rotary_sin, cache_batch_idx, leftpad_k,
                                 block_table, alibi_slopes, None, softmax_scale, is_causal, window_left, window_right,
                                 softcap, is_rotary_interleaved, num_splits))

    impls = {"fwd": f_fwd, "bwd": f_bwd, "varlen_fwd": f_varlen_fwd, "varlen_bwd": f_varlen_bwd, "fwd_kvcache": f_kvcache}
    fakes = {
        "fwd": lambda q, k, v, a, p, s, c, wl, wr, sc, rs: tuple(_fake_fwd(q, k, v, None, a, p, s, c, wl, wr, sc, rs)),
        "bwd": lambda do, q, k, v, o, lse, a, p, s, c, wl, wr, sc, rng: tuple(
            _fake_bwd(do, q, k, v, o, lse, None, None, None, a, p, s, c, wl, wr, sc, False)),
        "varlen_fwd": lambda q, k, v, cq, ck, bt, a, mq, mk, p, s, c, wl, wr, sc, rs: tuple(
            _fake_varlen_fwd(q, k, v, None, cq, ck, None, None, bt, a, mq, mk, p, s, False, c, wl, wr, sc, rs)),
        "varlen_bwd": lambda do, q, k, v, o, lse, cq, ck, a, mq, mk, p, s, c, wl, wr, sc, rng: tuple(
            _fake_varlen_bwd(do, q, k, v, o, lse, None, None, None, cq, ck, a, mq, mk, p, s, False, c, wl, wr, sc, False)),
        "fwd_kvcache": lambda q, kc, vc, k, v, sl, rc, rs, cbi, lp, bt, a, s, c, wl, wr, sc, ri, ns: tuple(
            _fake_fwd_kvcache(q, kc, vc, k, v, sl, rc, rs, cbi, lp, bt, a, None, s, c, wl, wr, sc, ri, ns)),
    }
    try:
        for name, schema in _FUNCTIONAL_SCHEMAS.items():
            L.define(f"fa_b200::{name}", schema)
            L.impl(f"fa_b200::{name}", "CUDA")(impls[name])
            L.register_fake(f"fa_b200::{name}")(fakes[name])
        if _ops_registered:
            for name, fake in (("fwd", _fake_fwd), ("bwd", _fake_bwd), ("varlen_fwd", _fake_varlen_fwd),
                               ("varlen_bwd", _fake_varlen_bwd), ("fwd_kvcache", _fake_fwd_kvcache)):
                L.register_fake(f"flash_attn_v100::{name}")(fake)
    except RuntimeError:
        return False

    # autograd formulas: the same saved state and gradient calls as the autograd.Function wrappers of the API
    def fwd_setup(ctx, inputs, output):
        q, k, v, alibi, p_drop, scale, causal, wl, wr, softcap, _ = inputs
        out, lse, _, rng = output
        ctx.save_for_backward(q, k, v, out, lse, rng)
        ctx.alibi, ctx.opts = alibi, (p_drop, scale, causal, wl, wr, softcap)

    def fwd_backward(ctx, dout, dlse, ddmask, drng):
        q, k, v, out, lse, rng = ctx.saved_tensors
        p_drop, scale, causal, wl, wr, softcap = ctx.opts
        dq, dk, dv, _ = torch.ops.fa_b200.bwd(dout, q, k, v, out, lse, ctx.alibi, p_drop, scale, causal, wl, wr, softcap, rng)
        return dq, dk, dv, None, None, None, None, None, None, None, None

    def varlen_setup(ctx, inputs, output):
        q, k, v, cq, ck, bt, alibi, mq, mk, p_drop, scale, causal, wl, wr, softcap, _ = inputs
        out, lse, _, rng = output
        ctx.save_for_backward(q, k, v, out, lse, cq, ck, rng)
        ctx.alibi, ctx.paged, ctx.opts = alibi, bt is not None, (mq, mk, p_drop, scale, causal, wl, wr, softcap)

    def varlen_backward(ctx, dout, dlse, ddmask, drng):
        if ctx.paged:
            raise RuntimeError("the backward has no paged-KV form (the reference's varlen_bwd takes no block_table)")
        q, k, v, out, lse, cq, ck, rng = ctx.saved_tensors
        mq, mk, p_drop, scale, causal, wl, wr, softcap = ctx.opts
        dq, dk, dv, _ = torch.ops.fa_b200.varlen_bwd(dout, q, k, v, out, lse, cq, ck, ctx.alibi, mq, mk, p_drop, scale,
                                                     causal, wl, wr, softcap, rng)
        return (dq, dk, dv) + (None,) * 13

    L.register_autograd("fa_b200::fwd", fwd_backward, setup_context=fwd_setup)
    L.register_autograd("fa_b200::varlen_fwd", varlen_backward, setup_context=varlen_setup)
    _functional_ops_registered = True
    return True


register_functional_ops()
