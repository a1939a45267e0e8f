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


# ------------------------------------------------------------------------------------------------ fused (peer-store) path


class FusedEP:
    """Device-driven expert-parallel exchange over symmetric memory (`csrc/ep.cu`): token counts are exchanged and turned into
    segment tables on the GPU, rows are stored straight into the owners' receive buffers in the (local expert, source rank)
    order the grouped tcgen05 GEMM consumes, and the outputs are stored back to where their inputs came from.  No host sync,
    no `all_to_all` split lists, no re-sorting of received rows.

    Buffers (per rank, symmetric): the count post region, TWO receive buffers of `cap_rows` rows (forward activations /
    backward gradients are both resident during the expert backward) and one return buffer of `max_rows` rows."""

    def __init__(self, group, n_experts: int, hidden: int, max_rows: int, dtype=torch.bfloat16, device=None, cap_factor: float = 2.0):
        from realhf_b200.ops import lib
        from realhf_b200.parallel.symm_mem import SymmetricBuffer, VmmSymmetricBuffer, multicast_supported
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        assert n_experts % self.world == 0
        self.E, self.e_local, self.H = n_experts, n_experts // self.world, hidden
        self.dtype, self.es = dtype, torch.tensor([], dtype=dtype).element_size()
        self.device = torch.device(device if device is not None else ("cuda", torch.cuda.current_device()))
        self.max_rows = max_rows
        self.cap_rows = int(cap_factor * max_rows)
        row = hidden * self.es
        assert row % 16 == 0
        post = (2 * self.world * self.e_local * 2 * 4 + 4095) // 4096 * 4096
        self.post_off = 0
        self.recv_off = [post, post + self.cap_rows * row]
        self.ret_off = post + 2 * self.cap_rows * row
        total = self.ret_off + max_rows * row
        flags = [None] * self.world
        dist.all_gather_object(flags, bool(multicast_supported(self.device)), group=group)
        cls = VmmSymmetricBuffer if all(flags) else SymmetricBuffer
        self.symm = cls(total, group=group, device=self.device)
        self.symm.data()[:].zero_()
        self.overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._calls = 0
        self._lib = lib

    # ---- views into my own buffers
    def _recv(self, which: int) -> torch.Tensor:
        n = self.cap_rows * self.H * self.es
        return self.symm.data()[self.recv_off[which]: self.recv_off[which] + n].view(self.dtype).view(self.cap_rows, self.H)

    def _ret(self, rows: int) -> torch.Tensor:
        n = rows * self.H * self.es
        return self.symm.data()[self.ret_off: self.ret_off + n].view(self.dtype).view(rows, self.H)

    def plan(self, counts: torch.Tensor):
        """counts [E] (device): my assignments per global expert -> (send_tab, ret_tab, per_expert [e_local + 1])."""
        par = self._calls & 1
        self._calls += 1
        return self._lib().ep_plan(counts.int().contiguous(), self.symm.data_ptrs, self.symm.pad_ptrs, self.post_off, self.e_local, self.rank, par,
                                   self.overflow, self.cap_rows)

    def dispatch(self, x_sorted: torch.Tensor, send_tab: torch.Tensor, which: int = 0) -> torch.Tensor:
        """Rows of x_sorted (grouped by global expert) -> receive buffer `which` of their experts' owners.  Returns MY receive
        buffer [cap_rows, H]: rows [0, per_expert[-1]) are valid, grouped by (local expert, source rank)."""
        self._lib().ep_move_rows(x_sorted, send_tab, self.symm.data_ptrs, self.symm.pad_ptrs, self.recv_off[which], self.cap_rows, self.rank, 64)
        return self._recv(which)

    def combine(self, y_recv: torch.Tensor, ret_tab: torch.Tensor, rows: int) -> torch.Tensor:
        """Rows of y_recv (the layout `dispatch` produced) -> back to the source ranks, at the positions their inputs had in the
        sender's sorted order.  Returns MY return buffer [rows, H] (a view: clone it to keep it beyond the next call)."""
        self._lib().ep_move_rows(y_recv, ret_tab, self.symm.data_ptrs, self.symm.pad_ptrs, self.ret_off, self.max_rows, self.rank, 64)
        return self._ret(rows)

    def raise_if_overflow(self):
        if int(self.overflow.item()) != 0:
            raise RuntimeError(f"expert-parallel receive buffer overflow: a rank was sent more than {self.cap_rows} rows; "
                               "raise REAL_EP_CAP_FACTOR (routing is badly imbalanced) or REAL_EP_MAX_ROWS")


class _FusedEPExperts(torch.autograd.Function):
    """y_sorted = combine(expert_mlp(dispatch(x_sorted))).  The expert MLP's activations are NOT kept: the backward pass
    re-dispatches x (into receive buffer 0) next to the gradient (receive buffer 1), re-runs the grouped expert MLP under
    autograd on the received rows and back-propagates through it, so only two shared receive buffers exist for the whole
    model, whatever its depth; expert weight gradients accumulate into the flat gradient buffer as usual."""

    @staticmethod
    def forward(ctx, x_sorted, counts, w_gu, w_dn, act: str, ep: FusedEP):
        from realhf_b200.models.moe import grouped_mlp_device
        send_tab, ret_tab, per_e = ep.plan(counts)
        x_recv = ep.dispatch(x_sorted, send_tab, 0)
        with torch.no_grad():
            y_recv = grouped_mlp_device(x_recv, per_e[: ep.e_local], w_gu, w_dn, act)
        y = ep.combine(y_recv, ret_tab, x_sorted.shape[0]).clone()
        ctx.save_for_backward(x_sorted, send_tab, ret_tab, per_e, w_gu, w_dn)
        ctx.act, ctx.ep = act, ep
        return y

    @staticmethod
    def backward(ctx, dy_sorted):
        from realhf_b200.models.moe import grouped_mlp_device
        x_sorted, send_tab, ret_tab, per_e, w_gu, w_dn = ctx.saved_tensors
        ep: FusedEP = ctx.ep
        x_recv = ep.dispatch(x_sorted, send_tab, 0)
        dy_recv = ep.dispatch(dy_sorted.contiguous(), send_tab, 1)
        with torch.enable_grad():
            x_in = x_recv.detach().requires_grad_(True)
            y = grouped_mlp_device(x_in, per_e[: ep.e_local], w_gu, w_dn, ctx.act)
        torch.autograd.backward(y, dy_recv)  # accumulates d(w_gu), d(w_dn) in place; leaves d(x_in)
        dx = ep.combine(x_in.grad, ret_tab, x_sorted.shape[0]).clone()
        return dx, None, None, None, None, None


def fused_ep_for(ctx, n_experts: int, hidden: int, dtype, device) -> "FusedEP | None":
    """The FusedEP workspace of a parallel context (created collectively on first use); None when peer memory is unavailable."""
    import os
    ep = getattr(ctx, "_fused_ep", None)
    if ep is None and not getattr(ctx, "_fused_ep_failed", False):
        try:
            ep = FusedEP(ctx.tp_group, n_experts, hidden, int(os.environ.get("REAL_EP_MAX_ROWS", "65536")), dtype=dtype, device=device,
                         cap_factor=float(os.environ.get("REAL_EP_CAP_FACTOR", "2.0")))
            ctx._fused_ep = ep
        except Exception as e:  # e.g. IPC not permitted in this container
            import warnings
            warnings.warn(f"fused expert-parallel exchange disabled: {e}")
            ctx._fused_ep_failed = True
    return ep


def dispatch_compute_combine_fused(x_sorted, expert_sorted, n_experts, w_gate_up, w_down, act, ep: FusedEP):
    from realhf_b200.models.moe import _count_experts
    counts = _count_experts(expert_sorted, n_experts).int()
    return _FusedEPExperts.apply(x_sorted.contiguous(), counts, w_gate_up, w_down, act, ep)
