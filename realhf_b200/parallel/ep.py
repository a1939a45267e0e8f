"""Expert parallelism: all-to-all token dispatch -> grouped expert GEMMs -> all-to-all combine.

The reference has no expert parallelism: every rank holds all experts, sliced along the FFN dim by TP, and its
"AlltoAll" dispatcher never leaves the rank (modules/moe/token_dispatcher.py:17-27).  Here `moe.expert_parallel=True`
partitions the experts over the tensor-parallel group (rank r owns experts [r*E/t, (r+1)*E/t), each kept whole):

  * sequence-parallel activations (tokens sharded over the group): tokens travel to their experts' owners with a
    variable-split all-to-all, are processed by the local grouped GEMMs and travel back (this module);
  * replicated activations: every rank computes only its own experts' assignments and the outputs are all-reduced
    (`models/moe.py`), no token exchange needed.
"""

from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


class _AllToAll(torch.autograd.Function):
    """Variable-split all-to-all along dim 0; backward is the all-to-all with the split lists swapped."""

    @staticmethod
    def forward(ctx, x, in_splits: List[int], out_splits: List[int], group):
        ctx.in_splits, ctx.out_splits, ctx.group = in_splits, out_splits, group
        out = x.new_empty((sum(out_splits), *x.shape[1:]))
        dist.all_to_all_single(out, x.contiguous(), output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        out = g.new_empty((sum(ctx.in_splits), *g.shape[1:]))
        dist.all_to_all_single(out, g.contiguous(), output_split_sizes=ctx.in_splits, input_split_sizes=ctx.out_splits, group=ctx.group)
        return out, None, None, None


def all_to_all(x, in_splits, out_splits, group):
    return _AllToAll.apply(x, in_splits, out_splits, group)


def dispatch_compute_combine(x_sorted: torch.Tensor, expert_sorted: torch.Tensor, n_experts: int, w_gate_up: torch.Tensor,
                             w_down: torch.Tensor, act: str, group) -> torch.Tensor:
    """x_sorted [A, H]: this rank's token copies sorted by (global) expert id `expert_sorted` [A].
    w_gate_up [E/t, 2F, H], w_down [E/t, H, F]: the local experts.  Returns y [A, H] in the same order as x_sorted."""
    from realhf_b200.models.moe import grouped_mlp
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    e_local = n_experts // world
    counts = torch.bincount(expert_sorted, minlength=n_experts)                 # assignments per global expert
    # exchange the per-expert counts so that every owner knows what arrives from whom
    recv_counts = torch.empty(world * e_local, dtype=counts.dtype, device=counts.device)
    dist.all_to_all_single(recv_counts, counts.contiguous(), group=group)       # [src, local expert]
    send_splits = counts.view(world, e_local).sum(1).tolist()
    recv_mat = recv_counts.view(world, e_local)
    recv_splits = recv_mat.sum(1).tolist()
    x_recv = all_to_all(x_sorted, send_splits, recv_splits, group)              # grouped by source rank, then local expert
    # regroup by local expert (stable): position of every received row
    src_ids = torch.repeat_interleave(torch.arange(world, device=x_recv.device), torch.tensor(recv_splits, device=x_recv.device))
    le = torch.repeat_interleave(torch.arange(e_local, device=x_recv.device).repeat(world), recv_mat.reshape(-1))
    order = torch.argsort(le, stable=True)
    x_by_e = x_recv.index_select(0, order)
    per_e = recv_mat.sum(0).tolist()
    y_by_e = grouped_mlp(x_by_e, per_e, w_gate_up, w_down, act)
    inv = torch.empty_like(order)
    inv[order] = torch.arange(order.numel(), device=order.device)
    y_recv = y_by_e.index_select(0, inv)
    return all_to_all(y_recv, recv_splits, send_splits, group)
