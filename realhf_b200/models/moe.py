"""Mixture-of-experts MLP: router, token permutation, grouped expert GEMMs, expert parallelism.

Parity: `realhf/impl/model/modules/moe/{router,experts,token_dispatcher,layer}.py` and `utils/moe.py`
(top-k softmax routing, aux load-balancing loss, z-loss, capacity-factor token dropping, sinkhorn).
Beyond the reference: real **expert parallelism** — with `ep_group` set, experts are partitioned across
ranks and tokens travel by all-to-all (`dispatch -> grouped GEMM -> combine`); the reference's
"AlltoAll" dispatcher never leaves the rank (token_dispatcher.py:17-27).
"""

from __future__ import annotations

import os
from typing import Dict, List, Tuple

import torch
import torch.distributed as dist
import torch.nn.functional as F

from realhf_b200.ops import functional as OF
from realhf_b200.parallel import tp as TP

MOE_STATS: Dict[str, List[torch.Tensor]] = {"aux_loss": [], "z_loss": []}


def pop_moe_losses() -> Dict[str, torch.Tensor]:
    out = {k: (torch.stack(v).sum() if v else None) for k, v in MOE_STATS.items()}
    for v in MOE_STATS.values():
        v.clear()
    return out


def sinkhorn(cost: torch.Tensor, tol: float = 1e-4, max_iter: int = 100) -> torch.Tensor:
    cost = torch.exp(cost.float())
    d0 = torch.ones(cost.size(0), device=cost.device)
    d1 = torch.ones(cost.size(1), device=cost.device)
    eps, err, it = 1e-8, 1e9, 0
    while err > tol and it < max_iter:
        d0 = (1.0 / d0.size(0)) / (torch.sum(d1.unsqueeze(0) * cost, 1) + eps)
        d1_new = (1.0 / d1.size(0)) / (torch.sum(d0.unsqueeze(1) * cost, 0) + eps)
        err = torch.mean(torch.abs(d1 - d1_new)).item()
        d1, it = d1_new, it + 1
    return d1 * cost * d0.unsqueeze(1)


def route(h: torch.Tensor, w_router: torch.Tensor, mcfg, training: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """h [T,H] -> (probs [T,k] fp32 normalised over the chosen experts, expert ids [T,k])."""
    if training and mcfg.input_jitter_eps > 0:
        h = h * torch.empty_like(h).uniform_(1 - mcfg.input_jitter_eps, 1 + mcfg.input_jitter_eps)
    logits = F.linear(h.float(), w_router.float())
    if training and mcfg.z_loss_coeff > 0:
        MOE_STATS["z_loss"].append(torch.logsumexp(logits, -1).square().mean() * mcfg.z_loss_coeff)
    if mcfg.routing_type == "sinkhorn" and training:
        with torch.no_grad():
            _, idx = torch.topk(sinkhorn(logits), mcfg.top_k, dim=-1)
        probs = torch.sigmoid(logits).gather(-1, idx) if mcfg.top_k == 1 else torch.softmax(logits, -1).gather(-1, idx)
        return probs, idx
    full = torch.softmax(logits, dim=-1)
    probs, idx = torch.topk(full, mcfg.top_k, dim=-1)
    probs = probs / probs.sum(-1, keepdim=True)
    if training and mcfg.routing_type == "aux_loss" and mcfg.aux_loss_coeff > 0:
        E = logits.shape[-1]
        frac_tokens = F.one_hot(idx, E).float().sum(1).mean(0)     # share of assignments per expert
        frac_probs = full.mean(0)
        MOE_STATS["aux_loss"].append((frac_tokens * frac_probs).sum() * E * mcfg.aux_loss_coeff / mcfg.top_k)
    return probs, idx


def _apply_capacity(probs, idx, n_experts: int, mcfg):
    """Zero the routing weight of assignments beyond each expert's capacity."""
    if mcfg.capacity_factor is None:
        return probs
    T, k = idx.shape
    cap = int(mcfg.capacity_factor * T * k / n_experts + 0.999)
    flat_e = idx.reshape(-1)
    if mcfg.token_drop_policy == "position":
        order = torch.arange(T * k, device=idx.device)
    else:
        order = torch.argsort(probs.reshape(-1), descending=True, stable=True)
    e_sorted = flat_e[order]
    onehot = F.one_hot(e_sorted, n_experts)
    rank_in_e = (onehot.cumsum(0) * onehot).sum(-1) - 1
    keep_sorted = rank_in_e < cap
    keep = torch.empty_like(keep_sorted)
    keep[order] = keep_sorted
    return probs * keep.view(T, k).to(probs.dtype)


def _count_experts(flat_e: torch.Tensor, n_experts: int) -> torch.Tensor:
    """Assignments per expert without a host sync (`torch.bincount` reads its input's maximum back to size the output, which also
    makes it illegal inside CUDA-graph capture)."""
    return torch.zeros(n_experts, dtype=torch.long, device=flat_e.device).scatter_add_(0, flat_e, torch.ones_like(flat_e))


def grouped_mlp_device(x_sorted: torch.Tensor, counts_dev: torch.Tensor, w_gate_up: torch.Tensor, w_down: torch.Tensor, act: str):
    """Expert MLP over tokens sorted by expert with the counts left on the DEVICE: two grouped tcgen05 GEMM launches and the
    gated activation, no `tokens_per_expert.cpu()` (reference: moe/experts.py:186) and no per-expert Python loop."""
    from realhf_b200.ops import gemm as G
    offsets = torch.zeros(counts_dev.numel() + 1, dtype=torch.int32, device=x_sorted.device)
    offsets[1:] = counts_dev.cumsum(0)
    h = G.grouped_linear(x_sorted, w_gate_up, offsets)
    return G.grouped_linear(OF.gated_act(h, act), w_down, offsets)


def _use_grouped_kernel(x: torch.Tensor, w_gate_up: torch.Tensor, w_down: torch.Tensor) -> bool:
    if not x.is_cuda or OF.gemm_impl() is None:
        return False
    from realhf_b200.ops import gemm as G
    return G.grouped_supported(x, w_gate_up) and G.grouped_supported(x, w_down)


def grouped_mlp(x_sorted: torch.Tensor, counts: List[int], w_gate_up: torch.Tensor, w_down: torch.Tensor, act: str):
    """Tokens sorted by expert; expert e owns rows [sum(counts[:e]), +counts[e]).  w_gate_up [E,2F,H], w_down [E,H,F]."""
    outs, off = [], 0
    for e, n in enumerate(counts):
        if n == 0:
            continue
        xe = x_sorted[off: off + n]
        outs.append(OF.linear(OF.gated_act(OF.linear(xe, w_gate_up[e]), act), w_down[e]))
        off += n
    if not outs:
        return x_sorted.new_zeros(0, w_down.shape[1])
    return torch.cat(outs, 0)


def moe_forward(model, i: int, h: torch.Tensor) -> torch.Tensor:
    """MoE MLP of block i on normalised hidden states h ([T, H], or [T/t, H] under sequence parallelism)."""
    c, ctx = model.config, model.ctx
    mcfg = c.moe
    E, k = mcfg.num_experts, mcfg.top_k
    w_gu, w_dn = model.p[f"{i}.mlp.experts.gate_up.weight"], model.p[f"{i}.mlp.experts.down.weight"]
    ep = mcfg.expert_parallel and ctx.tp_size > 1
    sp = model.sequence_parallel
    if ep and sp:
        pass  # tokens stay sharded: they travel to the experts' owners by all-to-all
    elif sp:
        h = TP.gather_from_sp(h, ctx)
    else:
        h = TP.copy_to_tp(h, ctx)
    T, H = h.shape
    probs, idx = route(h, model.p[f"{i}.mlp.router.weight"], mcfg, model.training)
    probs = _apply_capacity(probs, idx, E, mcfg)
    flat_e = idx.reshape(-1)
    order = torch.argsort(flat_e, stable=True)
    tok = torch.arange(T, device=h.device).repeat_interleave(k)[order]
    w_sorted = probs.reshape(-1)[order]
    if ep and sp:
        from realhf_b200.parallel import ep as EP
        x_sorted = h.index_select(0, tok)
        fep = None
        if h.is_cuda and h.dtype in (torch.bfloat16, torch.float16) and _use_grouped_kernel(x_sorted, w_gu, w_dn) \
                and os.environ.get("REAL_EP_FUSED", "1") == "1":
            fep = EP.fused_ep_for(ctx, E, H, h.dtype, h.device)
        if fep is not None and x_sorted.shape[0] <= fep.max_rows:
            # device-driven peer-store dispatch / combine around the grouped GEMMs (csrc/ep.cu): no host sync in this layer
            y_sorted = EP.dispatch_compute_combine_fused(x_sorted, flat_e[order], E, w_gu, w_dn, c.activation_function, fep)
        else:
            y_sorted = EP.dispatch_compute_combine(x_sorted, flat_e[order], E, w_gu, w_dn, c.activation_function, ctx.tp_group)
        return torch.zeros(T, H, dtype=y_sorted.dtype, device=h.device).index_add_(0, tok, y_sorted * w_sorted.to(y_sorted.dtype).unsqueeze(-1))
    if ep and not torch.is_grad_enabled() and _use_grouped_kernel(h, w_gu, w_dn):
        # replicated tokens, sync-free (CUDA-graph capturable: the decode path): every rank sorts ALL assignments, the grouped GEMM
        # is pointed at the row range of MY experts through device-side offsets (rows of other experts are never touched), the
        # foreign rows are masked out and the partial outputs are all-reduced.  Inference only: the untouched rows hold
        # uninitialised memory, which `torch.where` discards in the forward pass but which would leak into the gradients
        # (0 * NaN in the routing-weight gradient, garbage dgrad rows scattered back into dh); training without SP takes the
        # compacting path below
        e_local = E // ctx.tp_size
        lo = ctx.tp_rank * e_local
        e_sorted = flat_e[order]
        offs = torch.zeros(E + 1, dtype=torch.int32, device=h.device)
        offs[1:] = _count_experts(flat_e, E).cumsum(0)
        from realhf_b200.ops import gemm as G
        x_sorted = h.index_select(0, tok)
        my_offs = offs[lo: lo + e_local + 1].contiguous()
        hid = G.grouped_linear(x_sorted, w_gu, my_offs)
        y_sorted = G.grouped_linear(OF.gated_act(hid, c.activation_function), w_dn, my_offs)
        mine = ((e_sorted >= lo) & (e_sorted < lo + e_local)).unsqueeze(-1)
        contrib = torch.where(mine, y_sorted * w_sorted.to(y_sorted.dtype).unsqueeze(-1), torch.zeros((), dtype=y_sorted.dtype, device=h.device))
        out = torch.zeros(T, H, dtype=h.dtype, device=h.device).index_add_(0, tok, contrib.to(h.dtype))
        return TP.reduce_from_tp(out, ctx)
    if ep:
        # replicated tokens: keep only the assignments of my experts, all-reduce the partial outputs
        e_local = E // ctx.tp_size
        lo = ctx.tp_rank * e_local
        e_sorted = flat_e[order]
        mine = (e_sorted >= lo) & (e_sorted < lo + e_local)
        sel = torch.nonzero(mine).squeeze(-1)
        x_sorted = h.index_select(0, tok[sel])
        counts = torch.bincount(e_sorted[sel] - lo, minlength=e_local).tolist()
        y_sorted = grouped_mlp(x_sorted, counts, w_gu, w_dn, c.activation_function)
        out = torch.zeros(T, H, dtype=h.dtype, device=h.device).index_add_(
            0, tok[sel], (y_sorted * w_sorted[sel].to(y_sorted.dtype).unsqueeze(-1)).to(h.dtype))
        return TP.reduce_from_tp(out, ctx)
    x_sorted = h.index_select(0, tok)
    if mcfg.use_grouped_gemm and _use_grouped_kernel(x_sorted, w_gu, w_dn):   # use_grouped_gemm=False: one GEMM pair per expert
        y_sorted = grouped_mlp_device(x_sorted, _count_experts(flat_e, E), w_gu, w_dn, c.activation_function)
    else:
        counts = torch.bincount(flat_e, minlength=E).tolist()
        y_sorted = grouped_mlp(x_sorted, counts, w_gu, w_dn, c.activation_function)
    out = torch.zeros(T, H, dtype=y_sorted.dtype, device=h.device).index_add_(0, tok, y_sorted * w_sorted.to(y_sorted.dtype).unsqueeze(-1))
    if ctx.tp_size > 1:  # experts are F-sharded over TP: partial sums
        out = TP.reduce_scatter_to_sp(out, ctx) if sp else TP.reduce_from_tp(out, ctx)
    return out
