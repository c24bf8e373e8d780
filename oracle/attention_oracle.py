"""CPU oracle for the attention forward hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module; the product path (flash-attention-v100_b200/) never does and has no CPU fallback.

It restates, tile-free and in float64 (or float32 on request), what the reference's forward computes:

  * scores, masks, bias, softcap ............ reference include/mat_mul.h:82-157
      S = (Q K^T) * scale;  causal/window mask bottom-right aligned with off = Sk - Sq
      (masked iff j - off > i + WR, or j - off < i - WL);  ALiBi  S -= slope * |i + off - j|;
      softcap S = cap * tanh(S / cap) applied AFTER ALiBi (the reference's order, :111-117)
  * softmax / output / LSE .................. reference include/softmax.h:80-95,186-188,
      kernel/fused_mha_forward.cu:220-223, include/gemm_smem.h:143-146
      online softmax == exact softmax;  out = P V;  lse = m + ln(l)   (natural log)
  * GQA head map ............................ reference include/template.h:71-73  (kv_head = h // (H/Hk))
  * host normalisations ..................... reference kernel/fused_mha_forward.cu:343,351-352,
      kernel/fused_mha_forward_kvcache.cu:465-466  (Sq==1 && no alibi => non-causal; window >= Sk => -1;
      causal => window_right = 0)
  * varlen layouts .......................... reference include/template.h:199-230,
      kernel/fused_mha_forward_varlen.cu:100-111,519  (packed (T,H,D); lse [H,T]; no-key tiles => 0/-1e30)
  * kv-cache append / lengths / paging ...... reference kernel/fused_mha_forward_kvcache.cu:79-86,
      include/rotary.h:53-76;  paged addressing kernel/fused_mha_forward_varlen.cu:184-193
  * RoPE .................................... reference include/rotary.h:89-143,176-257
      y0 = fma(x0, c, -(x1*s)), y1 = fma(x0, s, x1*c) in fp32, rounded to the 16-bit dtype
  * dropout ................................. reference include/softmax.h:96-125, include/philox.h
      (Philox4x32-10; pinned on the Random123 known-answer vectors)
  * backward ................................ reference include/softmax.h:205-406, include/product.h,
      kernel/fused_mha_backward.cu (explicit formulas in attention_bwd_one; pinned on fixtures produced by
      executing the reference's `ref_mha_backward`, test.py:36-61, and cross-checked against autograd)

Documented positions where the reference is internally inconsistent (SURVEY 8a "quirks"): rows with no
visible key give out = 0 and lse = -1e30 (the reference does this for wholly skipped tiles and the
kv-cache kernel, ..._kvcache.cu:288-294); ALiBi always uses the bottom-right offset; any Sk is handled.

Parity pinning: tests/test_oracle.py checks this file against (1) the reference's own CPU
implementation `cpu_attention` (utils/sass/mma_swizzle/forward_kernel.cu:346-370) compiled from the
reference sources into oracle/_ref, on its known-answer case (:394-407,:433,:439), via the committed
fixture tests/golden/ref_cpu_attention_kat.npz; (2) fixtures produced by executing the reference's
`ref_mha_forward` (test.py:18-34) on its seed-421 inputs (tests/golden/ref_mha_forward_*.npz);
(3) torch SDPA on CPU for the subset SDPA expresses.  Features the reference never tests (varlen,
kv-cache, paged, rotary, ALiBi, softcap, window, GQA, Sq != Sk) are "parity unpinned" by the
reference itself; they are pinned here only against (3) and against upstream-API semantics.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch

NEG_SENTINEL = -1e30  # reference include/kernel.h:20


def normalize_mask_args(seqlen_q: int, seqlen_k: int, causal: bool, window: Tuple[int, int],
                        has_alibi: bool, kvcache: bool = False) -> Tuple[int, int]:
    """Host-side normalisation -> (window_left, window_right) with the causal mask folded in."""
    wl, wr = int(window[0]), int(window[1])
    if seqlen_q == 1 and not has_alibi:
        causal = False
    if kvcache and causal:
        wr = 0
    if wl >= seqlen_k:
        wl = -1
    if wr >= seqlen_k:
        wr = -1
    if causal:
        wr = 0  # kernel applies both masks; causal is the tighter right bound
    return wl, wr


# ----------------------------------------------------------------------------- dropout (Philox4x32-10)
_PHILOX_M0, _PHILOX_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85
_U32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter_lo64: np.ndarray, key: int, counter_hi64: int = 0) -> np.ndarray:
    """Philox4x32-10 (Salmon et al., SC'11; reference include/philox.h:13-64). `counter_lo64`: uint64 array
    holding counter words (x, y); `counter_hi64` words (z, w); `key` 64-bit. Returns uint32 [..., 4]."""
    c = np.asarray(counter_lo64, dtype=np.uint64)
    x0, x1 = c & _U32, c >> np.uint64(32)
    x2 = np.full_like(x0, counter_hi64 & 0xFFFFFFFF)
    x3 = np.full_like(x0, (counter_hi64 >> 32) & 0xFFFFFFFF)
    k0, k1 = key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _PHILOX_M0 * x0, _PHILOX_M1 * x2          # 32x32 -> 64 bit, exact in uint64
        x0, x1, x2, x3 = ((p1 >> np.uint64(32)) ^ x1 ^ np.uint64(k0), p1 & _U32,
                          (p0 >> np.uint64(32)) ^ x3 ^ np.uint64(k1), p0 & _U32)
        k0, k1 = (k0 + _PHILOX_W0) & 0xFFFFFFFF, (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return np.stack([x0, x1, x2, x3], axis=-1).astype(np.uint32)


def dropout_keep_mask(p_dropout: float, seed: int, offset: int, row0: int, rows: int, cols: int,
                      row_len: int) -> torch.Tensor:
    """bool [rows, cols]: element (row0 + r, c) survives. Reference include/softmax.h:50-51,96-109: flat index
    idx = row * row_len + col (no batch / head term), counter = offset + (idx >> 2), word = idx & 3, keep iff
    word <= uint32((1 - p) * 4294967295.0f) evaluated in float32. (For row_len % 4 != 0 the reference reuses one
    counter for 4 consecutive columns of a row; this restatement -- and the CUDA path -- use the pure
    function of idx, identical whenever row_len % 4 == 0, the only case the reference's dense path handles.)"""
    thr = int(np.uint32(np.float32(np.float32(1.0) - np.float32(p_dropout)) * np.float32(4294967295.0)))
    r = np.arange(rows, dtype=np.uint64).reshape(-1, 1) + np.uint64(row0)
    idx = r * np.uint64(row_len) + np.arange(cols, dtype=np.uint64).reshape(1, -1)
    ctr = (np.uint64(offset) + (idx >> np.uint64(2))).reshape(-1)
    uniq, inv = np.unique(ctr, return_inverse=True)
    words = philox4x32_10(uniq, seed & (2 ** 64 - 1))[inv.reshape(-1)]
    sel = np.take_along_axis(words, (idx.reshape(-1, 1) & np.uint64(3)).astype(np.int64), axis=1).reshape(rows, cols)
    return torch.from_numpy(sel <= np.uint32(thr))


def attention_one(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float, wl: int, wr: int,
                  slopes: Optional[torch.Tensor], softcap: float,
                  dtype=torch.float64, keep: Optional[torch.Tensor] = None,
                  p_dropout: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """One sequence. q:[Sq,H,D] k,v:[Sk,Hk,D] -> out [Sq,H,D] (dtype), lse [H,Sq] (dtype).
    `keep` [Sq,Sk] bool + p_dropout: dropped P entries are zeroed, kept ones scaled by 1/(1-p); the
    normaliser l and the LSE use P before dropout (reference include/softmax.h:94,111-114)."""
    Sq, H, D = q.shape
    Sk, Hk, _ = k.shape
    g = H // Hk
    qf, kf, vf = q.to(dtype), k.to(dtype), v.to(dtype)
    kf = kf.repeat_interleave(g, dim=1)
    vf = vf.repeat_interleave(g, dim=1)
    out = torch.zeros((Sq, H, D), dtype=dtype)
    lse = torch.full((H, Sq), NEG_SENTINEL, dtype=dtype)
    if Sk == 0 or Sq == 0:
        return out, lse
    s = torch.einsum("qhd,khd->hqk", qf, kf) * scale
    i = torch.arange(Sq).view(Sq, 1)
    j = torch.arange(Sk).view(1, Sk)
    off = Sk - Sq
    if slopes is not None:
        s = s - slopes.to(dtype).view(H, 1, 1) * (i + off - j).abs().to(dtype)
    if softcap > 0.0:
        s = softcap * torch.tanh(s / softcap)
    masked = torch.zeros((Sq, Sk), dtype=torch.bool)
    if wr >= 0:
        masked |= j > i + off + wr
    if wl >= 0:
        masked |= j < i + off - wl
    s = s.masked_fill(masked.view(1, Sq, Sk), float("-inf"))
    m = s.max(dim=-1).values  # [H,Sq]
    has_key = torch.isfinite(m)
    m_safe = torch.where(has_key, m, torch.zeros_like(m))
    p = torch.exp(s - m_safe.unsqueeze(-1))
    l = p.sum(dim=-1)
    if keep is not None:
        p = p * keep.to(dtype).view(1, Sq, Sk) / (1.0 - p_dropout)
    o = torch.einsum("hqk,khd->qhd", p, vf)
    l_safe = torch.where(has_key, l, torch.ones_like(l))
    out = o / l_safe.t().unsqueeze(-1)
    out = torch.where(has_key.t().unsqueeze(-1), out, torch.zeros_like(out))
    lse = torch.where(has_key, m_safe + torch.log(l_safe), torch.full_like(m, NEG_SENTINEL))
    return out, lse


def attention_bwd_one(q, k, v, dout, scale: float, wl: int, wr: int, slopes, softcap: float,
                      dtype=torch.float64, keep: Optional[torch.Tensor] = None, p_dropout: float = 0.0):
    """Backward of one sequence, restating the reference's formulas (not autograd):
    q,dout:[Sq,H,D] k,v:[Sk,Hk,D] -> dq [Sq,H,D], dk,dv [Sk,Hk,D], delta [H,Sq].

      delta_i = sum_d dO[i,d] O[i,d]                         reference include/product.h:9-96
      P = exp(S - lse), P_drop = keep ? P/(1-p) : 0          include/softmax.h:270-291
      dP = dO V^T                                            kernel/fused_mha_backward.cu:160-164 (dOV)
      dS = (P_drop * dP - P * delta) * scale                 include/softmax.h:293-294
      softcap: dS *= 1 - (S/softcap)^2, S the capped score   include/softmax.h:296-299
      dQ = dS K;  dK = sum_group dS^T Q;  dV = sum_group P_drop^T dO   kernel/fused_mha_backward.cu:201-204, 351-470
    ALiBi is a constant bias: it shapes P but has no gradient path of its own."""
    Sq, H, D = q.shape
    Sk, Hk, _ = k.shape
    g = H // Hk
    qf, kf, vf, dof = q.to(dtype), k.to(dtype), v.to(dtype), dout.to(dtype)
    dq = torch.zeros((Sq, H, D), dtype=dtype)
    dk = torch.zeros((Sk, Hk, D), dtype=dtype)
    dv = torch.zeros((Sk, Hk, D), dtype=dtype)
    delta = torch.zeros((H, Sq), dtype=dtype)
    if Sk == 0 or Sq == 0:
        return dq, dk, dv, delta
    kr = kf.repeat_interleave(g, dim=1)
    vr = vf.repeat_interleave(g, dim=1)
    s = torch.einsum("qhd,khd->hqk", qf, kr) * scale
    i = torch.arange(Sq).view(Sq, 1)
    j = torch.arange(Sk).view(1, Sk)
    off = Sk - Sq
    if slopes is not None:
        s = s - slopes.to(dtype).view(H, 1, 1) * (i + off - j).abs().to(dtype)
    if softcap > 0.0:
        s = softcap * torch.tanh(s / softcap)
    masked = torch.zeros((Sq, Sk), dtype=torch.bool)
    if wr >= 0:
        masked |= j > i + off + wr
    if wl >= 0:
        masked |= j < i + off - wl
    sm = s.masked_fill(masked.view(1, Sq, Sk), float("-inf"))
    m = sm.max(dim=-1).values
    has_key = torch.isfinite(m)
    m_safe = torch.where(has_key, m, torch.zeros_like(m))
    e = torch.exp(sm - m_safe.unsqueeze(-1))
    l = torch.where(has_key, e.sum(dim=-1), torch.ones_like(m))
    p = e / l.unsqueeze(-1)                       # == exp(S - lse); 0 on masked entries and key-less rows
    p_drop = p if keep is None else p * keep.to(dtype).view(1, Sq, Sk) / (1.0 - p_dropout)
    o = torch.einsum("hqk,khd->qhd", p_drop, vr)
    delta = torch.einsum("qhd,qhd->hq", dof, o)
    dp = torch.einsum("qhd,khd->hqk", dof, vr)
    ds = (p_drop * dp - p * delta.unsqueeze(-1)) * scale
    if softcap > 0.0:
        ds = ds * (1.0 - (s / softcap) ** 2)
    ds = ds.masked_fill(masked.view(1, Sq, Sk), 0.0)
    dq = torch.einsum("hqk,khd->qhd", ds, kr)
    dk = torch.einsum("hqk,qhd->khd", ds, qf).view(Sk, Hk, g, D).sum(dim=2)
    dv = torch.einsum("hqk,qhd->khd", p_drop, dof).view(Sk, Hk, g, D).sum(dim=2)
    return dq, dk, dv, delta


def flash_attn_bwd_ref(dout, q, k, v, softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
                       alibi_slopes=None, dtype=torch.float64, dropout_p=0.0, rng_state=None):
    """Dense backward. (B,S,H,D) layouts -> dq, dk, dv (same layouts), softmax_d (B,H,Sq)."""
    dout, q, k, v = (t.detach().cpu() for t in (dout, q, k, v))
    B, Sq, H, D = q.shape
    Sk = k.shape[1]
    scale = D ** -0.5 if softmax_scale is None else softmax_scale
    wl, wr = normalize_mask_args(Sq, Sk, causal, window_size, alibi_slopes is not None)
    keep = None
    if dropout_p > 0.0:
        keep = dropout_keep_mask(dropout_p, int(rng_state[0]), int(rng_state[1]), 0, Sq, Sk, Sk)
    res = [attention_bwd_one(q[b], k[b], v[b], dout[b], scale, wl, wr, _slopes_for(alibi_slopes, b), softcap,
                             dtype, keep, dropout_p) for b in range(B)]
    return tuple(torch.stack([r[n] for r in res]) for n in range(4))


def flash_attn_varlen_bwd_ref(dout, q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                              softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
                              alibi_slopes=None, dtype=torch.float64, dropout_p=0.0, rng_state=None):
    """Packed backward. q,dout:(T,H,D) k,v:(Tk,Hk,D) -> dq, dk, dv, softmax_d (H,T)."""
    dout, q, k, v = (t.detach().cpu() for t in (dout, q, k, v))
    cu_q, cu_k = cu_seqlens_q.detach().cpu().long(), cu_seqlens_k.detach().cpu().long()
    T, H, D = q.shape
    B = cu_q.numel() - 1
    scale = D ** -0.5 if softmax_scale is None else softmax_scale
    causal_eff = causal and not (max_seqlen_q == 1 and alibi_slopes is None)
    wl, wr = int(window_size[0]), int(window_size[1])
    if wl >= max_seqlen_k:
        wl = -1
    if wr >= max_seqlen_k:
        wr = -1
    if causal_eff:
        wr = 0
    dq = torch.zeros((T, H, D), dtype=dtype)
    dk = torch.zeros(tuple(k.shape), dtype=dtype)
    dv = torch.zeros(tuple(k.shape), dtype=dtype)
    delta = torch.zeros((H, T), dtype=dtype)
    for b in range(B):
        qs, qe, ks, ke = int(cu_q[b]), int(cu_q[b + 1]), int(cu_k[b]), int(cu_k[b + 1])
        keep = None
        if dropout_p > 0.0 and qe > qs and ke > ks:
            keep = dropout_keep_mask(dropout_p, int(rng_state[0]), int(rng_state[1]), qs, qe - qs, ke - ks, int(max_seqlen_k))
        a, b_, c, d = attention_bwd_one(q[qs:qe], k[ks:ke], v[ks:ke], dout[qs:qe], scale, wl, wr,
                                        _slopes_for(alibi_slopes, b), softcap, dtype, keep, dropout_p)
        dq[qs:qe], dk[ks:ke], dv[ks:ke], delta[:, qs:qe] = a, b_, c, d
    return dq, dk, dv, delta


def naive_lowp_attention_bwd(q, k, v, dout, scale, causal):
    """The same-precision yardstick of reference test.py:36-61 (`ref_mha_backward`, upcast=False): autograd
    through einsum / softmax / einsum in the tensors' own 16-bit dtype. (B,S,H,D), MHA only, top-left causal
    as in test.py (it only uses Sq == Sk)."""
    q, k, v = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * scale
    if causal:
        mask = torch.triu(torch.ones(s.shape[-2], s.shape[-1], device=s.device, dtype=torch.bool), diagonal=1)
        s = s.masked_fill(mask, float("-inf"))
    o = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, dim=-1), v)
    return torch.autograd.grad(o, (q, k, v), dout)


def _slopes_for(alibi_slopes: Optional[torch.Tensor], b: int) -> Optional[torch.Tensor]:
    if alibi_slopes is None:
        return None
    a = alibi_slopes.detach().cpu()
    return a[b] if a.dim() == 2 else a


def flash_attn_func_ref(q, k, v, softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
                        alibi_slopes=None, dtype=torch.float64, dropout_p=0.0, rng_state=None):
    """Dense. q:(B,Sq,H,D) k,v:(B,Sk,Hk,D) -> out (B,Sq,H,D), lse (B,H,Sq).
    dropout_p > 0 needs rng_state = (seed, offset); every (batch, head) shares one mask, as in the reference."""
    q, k, v = q.detach().cpu(), k.detach().cpu(), v.detach().cpu()
    B, Sq, H, D = q.shape
    Sk = k.shape[1]
    scale = D ** -0.5 if softmax_scale is None else softmax_scale
    if Sk == 0:  # reference kernel/fused_mha_forward.cu:409-413
        return torch.zeros((B, Sq, H, D), dtype=dtype), torch.full((B, H, Sq), float("-inf"), dtype=dtype)
    wl, wr = normalize_mask_args(Sq, Sk, causal, window_size, alibi_slopes is not None)
    keep = None
    if dropout_p > 0.0:
        keep = dropout_keep_mask(dropout_p, int(rng_state[0]), int(rng_state[1]), 0, Sq, Sk, Sk)
    outs, lses = [], []
    for b in range(B):
        o, l = attention_one(q[b], k[b], v[b], scale, wl, wr, _slopes_for(alibi_slopes, b), softcap, dtype,
                             keep, dropout_p)
        outs.append(o)
        lses.append(l)
    return torch.stack(outs), torch.stack(lses)


def _gather_paged(cache: torch.Tensor, block_table_row: torch.Tensor, length: int) -> torch.Tensor:
    """cache:(num_pages, page, Hk, D); logical rows [0,length) through one row of the block table."""
    page = cache.shape[1]
    n_pages = (length + page - 1) // page
    if n_pages == 0:
        return cache.new_zeros((0,) + tuple(cache.shape[2:]))
    pages = cache[block_table_row[:n_pages].long()]
    return pages.reshape(-1, *cache.shape[2:])[:length]


def flash_attn_varlen_func_ref(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                               softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
                               alibi_slopes=None, block_table=None, seqused_k=None, dtype=torch.float64,
                               dropout_p=0.0, rng_state=None):
    """Packed. q:(T,H,D); k,v:(Tk,Hk,D) or paged (num_pages,page,Hk,D) -> out (T,H,D), lse (H,T).
    Dropout index: row = packed q row, row length = max_seqlen_k (reference ..._varlen.cu:235)."""
    q, k, v = q.detach().cpu(), k.detach().cpu(), v.detach().cpu()
    cu_q = cu_seqlens_q.detach().cpu().long()
    cu_k = cu_seqlens_k.detach().cpu().long()
    T, H, D = q.shape
    B = cu_q.numel() - 1
    scale = D ** -0.5 if softmax_scale is None else softmax_scale
    has_alibi = alibi_slopes is not None
    causal_eff = causal and not (max_seqlen_q == 1 and not has_alibi)
    wl, wr = int(window_size[0]), int(window_size[1])
    if wl >= max_seqlen_k:
        wl = -1
    if wr >= max_seqlen_k:
        wr = -1
    if causal_eff:
        wr = 0
    out = torch.zeros((T, H, D), dtype=dtype)
    lse = torch.full((H, T), NEG_SENTINEL, dtype=dtype)
    for b in range(B):
        qs, qe = int(cu_q[b]), int(cu_q[b + 1])
        ks, ke = int(cu_k[b]), int(cu_k[b + 1])
        lk = ke - ks
        if seqused_k is not None:  # reference include/template.h:65-68
            u = int(seqused_k[b])
            lk = min(lk, u) if u > 0 else 0
        if block_table is not None:
            bt = block_table.detach().cpu()[b]
            kb, vb = _gather_paged(k, bt, lk), _gather_paged(v, bt, lk)
        else:
            kb, vb = k[ks:ks + lk], v[ks:ks + lk]
        keep = None
        if dropout_p > 0.0 and qe > qs and lk > 0:
            keep = dropout_keep_mask(dropout_p, int(rng_state[0]), int(rng_state[1]), qs, qe - qs, lk, int(max_seqlen_k))
        o, l = attention_one(q[qs:qe], kb, vb, scale, wl, wr, _slopes_for(alibi_slopes, b), softcap, dtype,
                             keep, dropout_p)
        out[qs:qe] = o
        lse[:, qs:qe] = l
    return out, lse


def _round_like(x32: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    return x32.to(like.dtype)


def apply_rotary_ref(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, positions: torch.Tensor,
                     interleaved: bool) -> torch.Tensor:
    """x:(S,H,D) 16-bit; cos,sin:(seqlen_ro, rotary_dim/2) 16-bit; positions:(S,) -> rotated, dtype of x.

    fp32 arithmetic with the reference's operation order (include/rotary.h:95-105,108-137); numpy has no
    fma, so the fused step is evaluated in float64 and rounded once to float32 (exact for these operand
    widths: a 24x24-bit product plus a float32 addend fits the float64 significand except in rare
    double-rounding ties; tests allow 1 ulp of the 16-bit result for that reason).
    """
    S, H, D = x.shape
    rot = 2 * cos.shape[1]
    xf = x.float().double()
    c = cos[positions.long()].float().double().unsqueeze(1)  # (S,1,rot/2)
    s = sin[positions.long()].float().double().unsqueeze(1)
    y = xf.clone()
    if interleaved:
        x0, x1 = xf[..., 0:rot:2], xf[..., 1:rot:2]
    else:
        x0, x1 = xf[..., : rot // 2], xf[..., rot // 2: rot]
    t0 = (x1 * s).float().double()  # x1*s rounded to fp32 first
    t1 = (x1 * c).float().double()
    y0 = (x0 * c - t0).float()
    y1 = (x0 * s + t1).float()
    y = y.float()
    if interleaved:
        y[..., 0:rot:2], y[..., 1:rot:2] = y0, y1
    else:
        y[..., : rot // 2], y[..., rot // 2: rot] = y0, y1
    return y.to(x.dtype)


def flash_attn_with_kvcache_ref(q, k_cache, v_cache, k=None, v=None, rotary_cos=None, rotary_sin=None,
                                cache_seqlens=None, cache_batch_idx=None, cache_leftpad=None, block_table=None,
                                softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
                                rotary_interleaved=True, alibi_slopes=None, dtype=torch.float64):
    """KV-cache forward. Returns (out (B,Sq,H,D), lse (B,H,Sq), k_cache_after, v_cache_after).

    The caches are cloned, updated like the reference does in place, and returned for comparison.
    """
    q = q.detach().cpu()
    kc, vc = k_cache.detach().cpu().clone(), v_cache.detach().cpu().clone()
    B, Sq, H, D = q.shape
    paged = block_table is not None
    bt = block_table.detach().cpu() if paged else None
    page = kc.shape[1] if paged else None
    capacity = bt.shape[1] * page if paged else kc.shape[1]
    scale = D ** -0.5 if softmax_scale is None else softmax_scale
    has_alibi = alibi_slopes is not None
    if isinstance(cache_seqlens, int):
        cache_seqlens = torch.full((B,), cache_seqlens, dtype=torch.int32)
    lens = cache_seqlens.detach().cpu().long() if cache_seqlens is not None else None
    s_new = k.shape[1] if k is not None else 0
    causal_eff = causal and not (Sq == 1 and not has_alibi)
    wl, wr = int(window_size[0]), int(window_size[1])
    if causal_eff:
        wr = 0
    if wl >= capacity:
        wl = -1
    if wr >= capacity:
        wr = -1
    per_row_pos = causal_eff or wl >= 0 or wr >= 0
    outs, lses = [], []
    for b in range(B):
        # absent cache_seqlens => the whole cache is valid (upstream API)
        len_b = int(lens[b]) if lens is not None else capacity
        pad_b = int(cache_leftpad[b]) if cache_leftpad is not None else 0
        cb = int(cache_batch_idx[b]) if cache_batch_idx is not None else b
        if k is not None:
            pos = torch.arange(s_new) + len_b + pad_b  # reference include/rotary.h:62
            kn, vn = k[b].detach().cpu(), v[b].detach().cpu()
            if rotary_cos is not None:
                kn = apply_rotary_ref(kn, rotary_cos.detach().cpu(), rotary_sin.detach().cpu(), pos, rotary_interleaved)
            for r in range(s_new):
                pr = int(pos[r])
                if paged:
                    kc[int(bt[b, pr // page]), pr % page] = kn[r]
                    vc[int(bt[b, pr // page]), pr % page] = vn[r]
                else:
                    kc[cb, pr] = kn[r]
                    vc[cb, pr] = vn[r]
        total = len_b + s_new if lens is not None else capacity
        qb = q[b]
        if rotary_cos is not None:
            qpos = torch.full((Sq,), len_b + pad_b) + (torch.arange(Sq) if per_row_pos else 0)
            qb = apply_rotary_ref(qb, rotary_cos.detach().cpu(), rotary_sin.detach().cpu(), qpos, rotary_interleaved)
        if paged:
            kb, vb = _gather_paged(kc, bt[b], total), _gather_paged(vc, bt[b], total)
        else:
            kb, vb = kc[cb, pad_b:pad_b + total], vc[cb, pad_b:pad_b + total]
        o, l = attention_one(qb, kb, vb, scale, wl, wr, _slopes_for(alibi_slopes, b), softcap, dtype)
        outs.append(o)
        lses.append(l)
    return torch.stack(outs), torch.stack(lses), kc, vc


# ----------------------------------------------------------------------------- tolerance rule
def naive_lowp_attention(q, k, v, scale, wl, wr, dtype=None):
    """The 'naive same-precision PyTorch implementation' yardstick of reference test.py:18-34 with
    upcast=False: einsum / softmax / einsum evaluated in the tensors' own 16-bit dtype. (B,S,H,D)."""
    B, Sq, H, D = q.shape
    Sk, Hk = k.shape[1], k.shape[2]
    g = H // Hk
    kk = k.repeat_interleave(g, dim=2)
    vv = v.repeat_interleave(g, dim=2)
    s = torch.einsum("bqhd,bkhd->bhqk", q, kk) * scale
    i = torch.arange(Sq, device=q.device).view(Sq, 1)
    j = torch.arange(Sk, device=q.device).view(1, Sk)
    off = Sk - Sq
    masked = torch.zeros((Sq, Sk), dtype=torch.bool, device=q.device)
    if wr >= 0:
        masked |= j > i + off + wr
    if wl >= 0:
        masked |= j < i + off - wl
    s = s.masked_fill(masked, float("-inf"))
    p = torch.softmax(s, dim=-1)
    p = torch.nan_to_num(p, nan=0.0)
    return torch.einsum("bhqk,bkhd->bqhd", p, vv)


def fa_tolerance_ok(out_test: torch.Tensor, out_ref: torch.Tensor, out_naive: torch.Tensor,
                    factor: float = 2.0, atol: float = 1e-5) -> Tuple[bool, float, float]:
    """Pass rule of reference test.py:273-277:  max|test-ref| <= 2 * max|naive-ref| + 1e-5, all finite."""
    ref = out_ref.double().cpu()
    err = (out_test.double().cpu() - ref).abs().max().item()
    err_naive = (out_naive.double().cpu() - ref).abs().max().item()
    ok = bool(torch.isfinite(out_test.float()).all().item()) and err <= factor * err_naive + atol
    return ok, err, err_naive


def c_rand_uniform_pm1(n: int, state: Optional[list] = None) -> np.ndarray:
    """glibc srand(42)/rand() stream used by the reference's known-answer harness
    (utils/sass/mma_swizzle/forward_kernel.cu:394-407): ((float)rand()/RAND_MAX - 0.5f) * 2.0f.
    Implemented through libc so the values are the very ones the harness generates."""
    import ctypes
    import ctypes.util

    libc = ctypes.CDLL(ctypes.util.find_library("c"))
    libc.rand.restype = ctypes.c_int
    out = np.empty(n, dtype=np.float32)
    rand_max = np.float32(2147483647)
    for t in range(n):
        r = np.float32(libc.rand())
        out[t] = (np.float32(r / rand_max) - np.float32(0.5)) * np.float32(2.0)
    return out
