"""Padding helpers under the upstream names (`flash_attn.bert_padding`), pure PyTorch.

The reference ships a copy of upstream's helpers (reference flash_attn/bert_padding.py:1-147); these are
written from the documented behaviour: pack the valid tokens of a padded batch into (total, ...) plus
cu_seqlens for `flash_attn_varlen_func`, and scatter them back.
"""
import torch
import torch.nn.functional as F


def index_first_axis(x: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """x[indices] along dim 0 for an arbitrary trailing shape."""
    return x.reshape(x.shape[0], -1).index_select(0, indices).reshape(-1, *x.shape[1:])


def index_put_first_axis(values: torch.Tensor, indices: torch.Tensor, first_axis_dim: int) -> torch.Tensor:
    """Inverse of index_first_axis: zeros(first_axis_dim, ...) with `values` written at `indices`."""
    out = torch.zeros(first_axis_dim, *values.shape[1:], device=values.device, dtype=values.dtype)
    out[indices] = values
    return out


def unpad_input(hidden_states: torch.Tensor, attention_mask: torch.Tensor, unused_mask: torch.Tensor = None):
    """(batch, seqlen, ...) + bool/int mask (batch, seqlen) ->
    (tokens (total, ...), indices (total,), cu_seqlens (batch+1,) int32, max_seqlen_in_batch, seqused (batch,))."""
    all_masks = attention_mask if unused_mask is None else attention_mask + unused_mask
    seqlens_in_batch = all_masks.sum(dim=-1, dtype=torch.int32)
    used_seqlens_in_batch = attention_mask.sum(dim=-1, dtype=torch.int32)
    indices = torch.nonzero(all_masks.flatten(), as_tuple=False).flatten()
    max_seqlen_in_batch = int(seqlens_in_batch.max().item()) if seqlens_in_batch.numel() else 0
    cu_seqlens = F.pad(torch.cumsum(seqlens_in_batch, dim=0, dtype=torch.int32), (1, 0))
    b, s = hidden_states.shape[:2]
    tokens = index_first_axis(hidden_states.reshape(b * s, *hidden_states.shape[2:]), indices)
    return tokens, indices, cu_seqlens, max_seqlen_in_batch, used_seqlens_in_batch


def pad_input(hidden_states: torch.Tensor, indices: torch.Tensor, batch: int, seqlen: int) -> torch.Tensor:
    """(total, ...) -> (batch, seqlen, ...) with zeros at the padded positions."""
    out = index_put_first_axis(hidden_states, indices, batch * seqlen)
    return out.reshape(batch, seqlen, *hidden_states.shape[1:])
