"""Packed <-> padded conversions.  Parity: `realhf/impl/model/utils/padding.py` (unpad_input / pad_input /
pad_sequence_parallel_input and the index helpers behind them).

Everything in this framework runs on PACKED batches (`[T, ...]` tokens of all sequences back to back + `cu_seqlens`); padded
`[B, S, ...]` tensors only appear at the border with code that wants them (HuggingFace models inside custom interfaces, user
analysis code).  All functions are plain indexing: differentiable, no host sync except where a length has to become a python int
(`max_seqlen`)."""

from __future__ import annotations

from typing import Tuple

import torch


def index_first_axis(x: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """x[indices] over the first axis (gradients scatter back)."""
    return x.index_select(0, indices)


def index_put_first_axis(values: torch.Tensor, indices: torch.Tensor, first_axis_dim: int) -> torch.Tensor:
    """Zeros of shape [first_axis_dim, ...] with rows `indices` set to `values` (inverse of `index_first_axis`)."""
    out = values.new_zeros((first_axis_dim,) + tuple(values.shape[1:]))
    return out.index_copy(0, indices, values)


def unpad_input(hidden: torch.Tensor, attention_mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int]:
    """[B, S, ...] + mask [B, S] (1 = token) -> (packed [T, ...], flat indices [T] into B*S, cu_seqlens int32 [B + 1], max_seqlen).
    Tokens keep their order inside a sequence; the mask may have holes (left padding, padding in the middle)."""
    mask = attention_mask.bool()
    lens = mask.sum(dim=1, dtype=torch.int32)
    indices = torch.nonzero(mask.flatten(), as_tuple=False).flatten()
    cu = torch.zeros(mask.shape[0] + 1, dtype=torch.int32, device=mask.device)
    cu[1:] = lens.cumsum(0)
    flat = hidden.reshape((mask.numel(),) + tuple(hidden.shape[2:]))
    return index_first_axis(flat, indices), indices, cu, int(lens.max()) if lens.numel() else 0


def pad_input(packed: torch.Tensor, indices: torch.Tensor, batch: int, seqlen: int) -> torch.Tensor:
    """Inverse of `unpad_input`: packed [T, ...] -> [batch, seqlen, ...] with zeros at the padding positions."""
    out = index_put_first_axis(packed, indices, batch * seqlen)
    return out.view((batch, seqlen) + tuple(packed.shape[1:]))


def pack_padded(padded: torch.Tensor, lens: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Right-padded [B, S, ...] + lengths [B] -> (packed [T, ...], cu_seqlens int32 [B + 1])."""
    S = padded.shape[1]
    mask = torch.arange(S, device=padded.device)[None, :] < lens[:, None]
    packed, _, cu, _ = unpad_input(padded, mask)
    return packed, cu


def pad_packed(packed: torch.Tensor, cu_seqlens: torch.Tensor, seqlen: int = 0, left: bool = False, value=0) -> torch.Tensor:
    """packed [T, ...] -> [B, seqlen, ...], right- (default) or left-padded with `value` (seqlen 0: the longest sequence)."""
    cu = cu_seqlens.long()
    lens = cu[1:] - cu[:-1]
    B = lens.numel()
    S = int(seqlen or (int(lens.max()) if B else 0))
    pos = torch.arange(S, device=packed.device)[None, :]
    if left:
        mask = pos >= (S - lens)[:, None]
    else:
        mask = pos < lens[:, None]
    out = packed.new_full((B * S,) + tuple(packed.shape[1:]), value)
    out[torch.nonzero(mask.flatten(), as_tuple=False).flatten()] = packed
    return out.view((B, S) + tuple(packed.shape[1:]))


def pad_sequence_parallel_input(ids: torch.Tensor, cu_seqlens: torch.Tensor, max_seqlen: int, tp: int, pad_id: int = 0):
    """Sequence parallelism shards the packed token axis over the TP group: T must be a multiple of tp.  Appends ONE fake sequence
    of `pad_id` tokens; returns (ids, cu_seqlens, max_seqlen, n_pad) -- the caller drops the last n_pad rows of the output."""
    T = ids.shape[0]
    pad = (-T) % tp
    if pad == 0:
        return ids, cu_seqlens, max_seqlen, 0
    ids = torch.cat([ids, ids.new_full((pad,), pad_id)])
    cu = torch.cat([cu_seqlens, (cu_seqlens[-1:] + pad)])
    return ids, cu, max(max_seqlen, pad), pad
