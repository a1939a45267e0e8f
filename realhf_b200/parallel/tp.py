"""Tensor / sequence-parallel primitives.

Parity: `realhf/impl/model/parallelism/model_parallel/mappings.py` (copy / reduce / scatter / gather
regions and their sequence-parallel variants) and `modules.py` (Column/Row parallel linear,
vocab-parallel embedding and cross entropy).  All functions take the shard's `ParallelContext`
explicitly and are the identity when `tp_size == 1`.

`col_linear` / `row_linear` are the two places where a GEMM is adjacent to a collective: when the
context carries a symmetric-memory workspace (`ctx.symm`) and the tensors qualify, they route to the
fused peer-memory kernels in `realhf_b200.parallel.fused_tp` (all-gather->GEMM, GEMM->reduce-scatter,
GEMM->all-reduce); otherwise they run GEMM + NCCL collective (the baseline path, also used on gloo/CPU).
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from realhf_b200.base.topology import ParallelContext
from realhf_b200.ops import functional as OF


def _tp(ctx: Optional[ParallelContext]) -> int:
    return 1 if ctx is None else ctx.tp_size


# ------------------------------------------------------------------------------------------- raw collectives


def _all_reduce(x, ctx, op=dist.ReduceOp.SUM):
    if _tp(ctx) == 1:
        return x
    x = x.contiguous()
    fused = getattr(ctx, "symm", None)
    if fused is not None and op == dist.ReduceOp.SUM and x.is_cuda and not torch.is_grad_enabled() \
            and x.dtype in (torch.bfloat16, torch.float16) and x.numel() * x.element_size() <= (1 << 20):
        return fused.all_reduce(x)  # decode-sized message: peer-memory / in-switch all-reduce instead of an NCCL launch
    dist.all_reduce(x, op=op, group=ctx.tp_group)
    return x


def _gather_first_dim(x, ctx):
    t = _tp(ctx)
    if t == 1:
        return x
    x = x.contiguous()
    out = torch.empty((x.shape[0] * t, *x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x, group=ctx.tp_group)
    return out


def _reduce_scatter_first_dim(x, ctx):
    t = _tp(ctx)
    if t == 1:
        return x
    x = x.contiguous()
    assert x.shape[0] % t == 0, x.shape
    out = torch.empty((x.shape[0] // t, *x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.reduce_scatter_tensor(out, x, group=ctx.tp_group)
    return out


def _split_first_dim(x, ctx):
    t = _tp(ctx)
    if t == 1:
        return x
    n = x.shape[0] // t
    return x[ctx.tp_rank * n:(ctx.tp_rank + 1) * n].contiguous()


def _gather_last_dim(x, ctx):
    t = _tp(ctx)
    if t == 1:
        return x
    parts = [torch.empty_like(x) for _ in range(t)]
    dist.all_gather(parts, x.contiguous(), group=ctx.tp_group)
    return torch.cat(parts, dim=-1)


def _split_last_dim(x, ctx):
    t = _tp(ctx)
    if t == 1:
        return x
    n = x.shape[-1] // t
    return x[..., ctx.tp_rank * n:(ctx.tp_rank + 1) * n].contiguous()


# ------------------------------------------------------------------------------------------- autograd regions


class _CopyToTP(torch.autograd.Function):
    """identity forward, all-reduce backward (input of a column-parallel linear)."""

    @staticmethod
    def forward(ctx, x, pctx):
        ctx.pctx = pctx
        return x

    @staticmethod
    def backward(ctx, g):
        return _all_reduce(g, ctx.pctx), None


class _ReduceFromTP(torch.autograd.Function):
    """all-reduce forward, identity backward (output of a row-parallel linear)."""

    @staticmethod
    def forward(ctx, x, pctx):
        return _all_reduce(x, pctx)

    @staticmethod
    def backward(ctx, g):
        return g, None


class _GatherFromSP(torch.autograd.Function):
    """all-gather along tokens forward, reduce-scatter backward (SP -> column-parallel input)."""

    @staticmethod
    def forward(ctx, x, pctx, rs_bwd):
        ctx.pctx, ctx.rs_bwd = pctx, rs_bwd
        return _gather_first_dim(x, pctx)

    @staticmethod
    def backward(ctx, g):
        if ctx.rs_bwd:
            return _reduce_scatter_first_dim(g, ctx.pctx), None, None
        return _split_first_dim(g, ctx.pctx), None, None


class _ReduceScatterToSP(torch.autograd.Function):
    """reduce-scatter along tokens forward, all-gather backward (row-parallel output -> SP)."""

    @staticmethod
    def forward(ctx, x, pctx):
        ctx.pctx = pctx
        return _reduce_scatter_first_dim(x, pctx)

    @staticmethod
    def backward(ctx, g):
        return _gather_first_dim(g, ctx.pctx), None


class _ScatterToSP(torch.autograd.Function):
    """split along tokens forward, all-gather backward (embedding output -> SP)."""

    @staticmethod
    def forward(ctx, x, pctx):
        ctx.pctx = pctx
        return _split_first_dim(x, pctx)

    @staticmethod
    def backward(ctx, g):
        return _gather_first_dim(g, ctx.pctx), None


class _GatherLastDim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pctx):
        ctx.pctx = pctx
        return _gather_last_dim(x, pctx)

    @staticmethod
    def backward(ctx, g):
        return _split_last_dim(g, ctx.pctx), None


def copy_to_tp(x, ctx):
    return x if _tp(ctx) == 1 else _CopyToTP.apply(x, ctx)


def reduce_from_tp(x, ctx):
    return x if _tp(ctx) == 1 else _ReduceFromTP.apply(x, ctx)


def gather_from_sp(x, ctx, reduce_scatter_bwd: bool = True):
    return x if _tp(ctx) == 1 else _GatherFromSP.apply(x, ctx, reduce_scatter_bwd)


def reduce_scatter_to_sp(x, ctx):
    return x if _tp(ctx) == 1 else _ReduceScatterToSP.apply(x, ctx)


def scatter_to_sp(x, ctx):
    return x if _tp(ctx) == 1 else _ScatterToSP.apply(x, ctx)


def gather_last_dim(x, ctx):
    return x if _tp(ctx) == 1 else _GatherLastDim.apply(x, ctx)


# ------------------------------------------------------------------------------------------- parallel linears


def col_linear(x, w, b, ctx: Optional[ParallelContext], sp: bool = False):
    """Column-parallel: w is the local [out/t, in] shard.  With SP the input arrives token-sharded."""
    if _tp(ctx) > 1:
        fused = getattr(ctx, "symm", None)
        if sp:
            if fused is not None and fused.can_ag_gemm(x, w):
                return fused.ag_gemm(x, w, b)
            x = gather_from_sp(x, ctx)
        else:
            x = copy_to_tp(x, ctx)
    return OF.linear(x, w, b)


def row_linear(x, w, b, ctx: Optional[ParallelContext], sp: bool = False):
    """Row-parallel: w is the local [out, in/t] shard; partial sums are reduced (all-reduce, or RS with SP)."""
    if _tp(ctx) == 1:
        return OF.linear(x, w, b)
    fused = getattr(ctx, "symm", None)
    if fused is not None and fused.can_gemm_reduce(x, w):
        y = fused.gemm_rs(x, w) if sp else fused.gemm_ar(x, w)
    else:
        y = OF.linear(x, w, None)
        y = reduce_scatter_to_sp(y, ctx) if sp else reduce_from_tp(y, ctx)
    return y if b is None else y + b


# ------------------------------------------------------------------------------------------- vocab parallel


def vocab_range(vocab_size: int, ctx: Optional[ParallelContext]) -> Tuple[int, int]:
    t = _tp(ctx)
    if t == 1:
        return 0, vocab_size
    per = vocab_size // t
    return ctx.tp_rank * per, (ctx.tp_rank + 1) * per


def vocab_parallel_embedding(ids, weight, ctx: Optional[ParallelContext], sp: bool = False):
    """weight is the local [V/t, H] shard. Out-of-shard ids contribute zeros, then reduce over TP."""
    if _tp(ctx) == 1:
        return torch.nn.functional.embedding(ids, weight)
    lo, hi = ctx.tp_rank * weight.shape[0], (ctx.tp_rank + 1) * weight.shape[0]
    mask = (ids < lo) | (ids >= hi)
    local = (ids - lo).masked_fill(mask, 0)
    out = torch.nn.functional.embedding(local, weight)
    out = out.masked_fill(mask.unsqueeze(-1), 0.0)
    return reduce_scatter_to_sp(out, ctx) if sp else reduce_from_tp(out, ctx)


class _VocabParallelLogProb(torch.autograd.Function):
    """log p(label) from vocab-sharded logits: one fused kernel pass per rank + ONE packed all-reduce
    (the reference issues three: max, sum-exp, target logit; modules.py:1056,1091,1101)."""

    @staticmethod
    def forward(ctx, logits, labels, pctx, inv_temp, mask_bits=None):
        V = logits.shape[-1]
        lo = pctx.tp_rank * V
        x = logits.float() * inv_temp
        if mask_bits is not None:  # bit-packed over the FULL vocabulary: take this rank's byte range
            assert lo % 8 == 0 and V % 8 == 0, "vocab shard must be byte aligned for the packed logits mask"
            local = OF.unpack_mask_bits(mask_bits[:, lo // 8:(lo + V) // 8], V)
            x = x.masked_fill(local, float("-inf"))
        m = x.max(dim=-1).values
        gm = m.clone()
        dist.all_reduce(gm, op=dist.ReduceOp.MAX, group=pctx.tp_group)
        ex = torch.exp(x - gm.unsqueeze(-1))
        own = (labels >= lo) & (labels < lo + V)
        loc = (labels - lo).masked_fill(~own, 0)
        tgt = x.gather(-1, loc.unsqueeze(-1)).squeeze(-1).masked_fill(~own, 0.0)
        packed = torch.stack([ex.sum(-1), tgt], dim=0)
        dist.all_reduce(packed, group=pctx.tp_group)
        lse = gm + packed[0].log()
        ctx.save_for_backward(ex / packed[0].unsqueeze(-1), loc, own)
        ctx.inv_temp = inv_temp
        return packed[1] - lse

    @staticmethod
    def backward(ctx, g):
        p, loc, own = ctx.saved_tensors
        grad = -p
        grad.scatter_add_(-1, loc.unsqueeze(-1), own.float().unsqueeze(-1))
        grad = grad * (g * ctx.inv_temp).unsqueeze(-1)
        return grad, None, None, None, None


def vocab_parallel_logprobs(logits, labels, ctx: Optional[ParallelContext], temperature: float = 1.0, mask_bits=None):
    if _tp(ctx) == 1:
        return OF.logprob_from_logits_ref(logits, labels, mask_bits, 1.0 / temperature)[0] if logits.requires_grad \
            else OF.logprob_from_logits(logits, labels, mask_bits, 1.0 / temperature)[0]
    return _VocabParallelLogProb.apply(logits, labels, ctx, 1.0 / temperature, mask_bits)
