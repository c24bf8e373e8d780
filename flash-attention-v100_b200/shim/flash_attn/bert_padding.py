"""Padding helpers under the upstream names (`flash_attn.bert_padding`), pure PyTorch.

The reference ships these names (reference flash_attn/bert_padding.py:9-147): three autograd Functions with their
`.apply` aliases -- IndexFirstAxis / index_first_axis, IndexPutFirstAxis / index_put_first_axis,
IndexFirstAxisResidual / index_first_axis_residual -- and unpad_input, unpad_input_for_concatenated_sequences,
pad_input. They are re-implemented here from that documented behaviour (row gather / scatter along the first axis with
hand-written backward passes, so the gradient of a gather is a scatter into a zero buffer instead of autograd's
generic index kernels): pack the valid tokens of a padded batch into (total, ...) plus cu_seqlens for
`flash_attn_varlen_func`, and scatter them back.
"""
import torch
import torch.nn.functional as F


def _rows(x: torch.Tensor) -> torch.Tensor:
    return x.reshape(x.shape[0], -1)


class IndexFirstAxis(torch.autograd.Function):
    """out = input[indices] along dim 0 (any trailing shape); backward scatters the rows back into zeros."""

    @staticmethod
    def forward(ctx, input: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:  # noqa: A002
        assert input.ndim >= 2
        ctx.save_for_backward(indices)
        ctx.first_axis_dim = input.shape[0]
        return _rows(input).index_select(0, indices).reshape(-1, *input.shape[1:])

    @staticmethod
    def backward(ctx, grad_output: torch.Tensor):
        (indices,) = ctx.saved_tensors
        assert grad_output.ndim >= 2
        tail = grad_output.shape[1:]
        flat = _rows(grad_output)
        grad_input = flat.new_zeros((ctx.first_axis_dim, flat.shape[1]))
        grad_input.index_copy_(0, indices, flat)  # indices are unique row numbers (positions of valid tokens)
        return grad_input.reshape(ctx.first_axis_dim, *tail), None


index_first_axis = IndexFirstAxis.apply


class IndexPutFirstAxis(torch.autograd.Function):
    """Inverse of IndexFirstAxis: zeros(first_axis_dim, ...) with `values` written at rows `indices`."""

    @staticmethod
    def forward(ctx, values: torch.Tensor, indices: torch.Tensor, first_axis_dim: int) -> torch.Tensor:
        assert indices.ndim == 1 and values.ndim >= 2
        ctx.save_for_backward(indices)
        out = values.new_zeros((first_axis_dim, *values.shape[1:]))
        out.index_copy_(0, indices, values)
        return out

    @staticmethod
    def backward(ctx, grad_output: torch.Tensor):
        (indices,) = ctx.saved_tensors
        return grad_output.index_select(0, indices), None, None


index_put_first_axis = IndexPutFirstAxis.apply


class IndexFirstAxisResidual(torch.autograd.Function):
    """Returns (input[indices], input.detach()): the gathered rows plus the untouched input as a residual branch
    whose gradient is accumulated in place with the scattered row gradients (reference bert_padding.py:54-74)."""

    @staticmethod
    def forward(ctx, input: torch.Tensor, indices: torch.Tensor):  # noqa: A002
        assert input.ndim >= 2
        ctx.save_for_backward(indices)
        ctx.first_axis_dim = input.shape[0]
        return input.index_select(0, indices), input.detach()

    @staticmethod
    def backward(ctx, grad_output: torch.Tensor, grad_residual: torch.Tensor):
        (indices,) = ctx.saved_tensors
        assert grad_output.ndim >= 2 and grad_residual.shape[1:] == grad_output.shape[1:]
        grad_input = grad_residual  # accumulated in place, like the reference
        grad_input.index_add_(0, indices, grad_output)
        return grad_input.reshape(ctx.first_axis_dim, *grad_output.shape[1:]), None


index_first_axis_residual = IndexFirstAxisResidual.apply


def _flatten_tokens(hidden_states: torch.Tensor) -> torch.Tensor:
    b, s = hidden_states.shape[:2]
    return hidden_states.reshape(b * s, *hidden_states.shape[2:])


def unpad_input(hidden_states: torch.Tensor, attention_mask: torch.Tensor, unused_mask: torch.Tensor = None):
    """(batch, seqlen, ...) + bool/int mask (batch, seqlen) [+ mask of allocated-but-unused slots] ->
    (tokens (total, ...), indices (total,), cu_seqlens (batch+1,) int32, max_seqlen_in_batch, seqlens (batch,) int32).
    The fifth value counts attention_mask + unused_mask, as the reference returns it (reference :93-105)."""
    all_masks = attention_mask if unused_mask is None else attention_mask + unused_mask
    seqlens_in_batch = all_masks.sum(dim=-1, dtype=torch.int32)
    indices = torch.nonzero(all_masks.flatten(), as_tuple=False).flatten()
    max_seqlen_in_batch = int(seqlens_in_batch.max().item()) if seqlens_in_batch.numel() else 0
    cu_seqlens = F.pad(torch.cumsum(seqlens_in_batch, dim=0, dtype=torch.int32), (1, 0))
    return index_first_axis(_flatten_tokens(hidden_states), indices), indices, cu_seqlens, max_seqlen_in_batch, seqlens_in_batch


def unpad_input_for_concatenated_sequences(hidden_states: torch.Tensor, attention_mask_in_length: torch.Tensor):
    """Several short samples packed into each row: `attention_mask_in_length[b]` lists the lengths of the samples
    concatenated in row b (zeros elsewhere), e.g. [2, 3, 0, 0, 0, 0] = a 2-token and a 3-token sample followed by
    padding. Returns (tokens (total, ...), indices (total,), cu_seqlens (num_samples+1,) int32, max_seqlen) with one
    cu_seqlens entry per SAMPLE, so each sample attends only to itself (reference :107-133)."""
    row_tokens = attention_mask_in_length.sum(dim=-1)  # tokens in use per row
    seqlen = attention_mask_in_length.shape[-1]
    positions = torch.arange(seqlen, device=row_tokens.device, dtype=row_tokens.dtype)
    in_use = positions.unsqueeze(0) < row_tokens.unsqueeze(1)  # (batch, seqlen): the row's first `row_tokens` slots
    lengths = attention_mask_in_length.flatten()
    seqlens_in_batch = lengths[torch.nonzero(lengths, as_tuple=False).flatten()]  # sample lengths, row by row
    indices = torch.nonzero(in_use.flatten(), as_tuple=False).flatten()
    max_seqlen_in_batch = int(seqlens_in_batch.max().item()) if seqlens_in_batch.numel() else 0
    cu_seqlens = F.pad(torch.cumsum(seqlens_in_batch, dim=0, dtype=torch.int32), (1, 0))
    return index_first_axis(_flatten_tokens(hidden_states), indices), indices, cu_seqlens, max_seqlen_in_batch


def pad_input(hidden_states: torch.Tensor, indices: torch.Tensor, batch: int, seqlen: int) -> torch.Tensor:
    """(total, ...) -> (batch, seqlen, ...) with zeros at the padded positions."""
    out = index_put_first_axis(hidden_states, indices, batch * seqlen)
    return out.reshape(batch, seqlen, *hidden_states.shape[1:])
