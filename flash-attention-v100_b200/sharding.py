"""Multi-GPU plumbing for the attention forward: batch/head sharding and timing reductions.

The reference has no distributed code at all; its kernels are independent per
(batch, query-head) -- the grid z-dimension is B*H_Q (reference kernel/fused_mha_forward.cu:260) and KV
heads are shared only inside a GQA group (reference include/template.h:73). So N GPUs need no data-path
collective: each rank owns a contiguous slice of the batch (or, when batch < ranks, of whole GQA groups).
torch.distributed (NCCL on the GPU box, gloo in the CPU tests) is used for barriers and for the
MAX-reduce of per-rank elapsed time only.
"""
from __future__ import annotations

from typing import List, Tuple


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) slice of `total` units for `rank` (first ranks get the remainder)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch_or_heads(batch: int, heads: int, heads_k: int, world: int, rank: int):
    """Returns ((b0, b1), (h0, h1)): batch slice first; if batch < world, split whole GQA groups instead
    so every query head stays on the GPU that holds its KV head."""
    if batch >= world:
        return shard_range(batch, world, rank), (0, heads)
    if batch * heads_k < world:
        raise ValueError("fewer (batch, kv-head) groups than ranks")
    per_b = world // batch  # ranks per batch element
    if world % batch:
        raise ValueError("world must be a multiple of batch when batch < world")
    b = rank // per_b
    g0, g1 = shard_range(heads_k, per_b, rank % per_b)
    group = heads // heads_k
    return (b, b + 1), (g0 * group, g1 * group)


def balanced_varlen_shards(seqlens: List[int], world: int, causal: bool = True) -> List[List[int]]:
    """Assign packed sequences to ranks balancing attention cost (sum of s^2 for causal/full): greedy LPT."""
    cost = [(s * s, i) for i, s in enumerate(seqlens)]
    loads = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for c, i in sorted(cost, reverse=True):
        r = loads.index(min(loads))
        out[r].append(i)
        loads[r] += c
    return [sorted(x) for x in out]


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """MAX-reduce a per-rank scalar (elapsed milliseconds) -- the only reduction the benchmark needs."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_checksums(value: float, dist=None, device=None) -> List[float]:
    """all_gather of one per-rank scalar so rank 0 can print a single parity verdict."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(value)]
    import torch

    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    outs = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return [float(x.item()) for x in outs]
