"""Parameter specs, tensor-parallel sharding rules and pipeline partitioning.

One table drives everything that needs to know "which slice of which tensor lives where": model
construction (flat buffer layout), HF load/save (shard / merge), and parameter reallocation (segment
plans between two layouts).  Parity: `realhf/impl/model/nn/real_llm_parallel.py` (TP split rules
:13-26,129-174, shapes :203-276, PP partition :342-375) and `real_llm_base.py:394-482` (ordered keys).

Global parameter names are `"{layer_idx}.{name}"`: layer 0 is the embedding, 1..L the blocks, L+1 the
head.  Fused projections are single tensors made of *sections* along dim 0 (qkv = [q | k | v],
gate_up = [gate | up]); a column split slices every section, so a TP shard is again [q_r | k_r | v_r].
"""

from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from realhf_b200.api.model import ReaLModelConfig


@dataclasses.dataclass(frozen=True)
class ParamSpec:
    name: str                     # global key
    shape: Tuple[int, ...]        # unsharded shape
    split_dim: Optional[int]      # None = replicated over TP
    sections: Tuple[int, ...] = ()  # sizes along dim 0 (only for split_dim == 0 fused tensors)
    kv_sections: Tuple[int, ...] = ()  # indices of `sections` that hold KV heads (replicated when n_kv < tp)
    expert_dim: bool = False      # leading dim enumerates experts (expert parallel splits it)
    init: str = "normal"          # normal | ones | zeros
    sp_grad_sync: bool = False    # replicated param whose grad is partial under sequence parallel (norms)


def layer_param_specs(cfg: ReaLModelConfig, layer_idx: int, tied_head_copy: bool = False) -> List[ParamSpec]:
    """Ordered specs of one layer (unsharded).  `tied_head_copy`: the stage holds the head but not the embedding of a
    tied-embedding model (pp > 1), so it keeps its own copy of the matrix; the two copies' gradients are all-reduced over
    the embedding group before every optimizer step (reference: megatron.py:589-605)."""
    H, F, hd = cfg.hidden_dim, cfg.intermediate_dim, cfg.head_dim
    nq, nkv, L = cfg.n_q_heads, cfg.n_kv_heads, cfg.n_layers
    ln_bias = cfg.layer_norm_type is None
    norm_init = "zeros" if cfg.layer_norm_type == "gemma" else "ones"
    P = lambda n, *a, **k: ParamSpec(f"{layer_idx}.{n}", *a, **k)
    out: List[ParamSpec] = []
    if layer_idx == 0:
        out.append(P("wte.weight", (cfg.vocab_size, H), 0))
        if not cfg.apply_rotary:
            out.append(P("wpe.weight", (cfg.n_positions, H), None))
        return out
    if layer_idx == L + 1:
        if cfg.is_critic:
            out.append(P("head.weight", (1, H), None))
        elif not cfg.tied_embedding or tied_head_copy:
            out.append(P("head.weight", (cfg.vocab_size, H), 0))
        return out

    def norm(prefix):
        out.append(P(f"{prefix}.weight", (H,), None, init=norm_init, sp_grad_sync=True))
        if ln_bias:
            out.append(P(f"{prefix}.bias", (H,), None, init="zeros", sp_grad_sync=True))

    norm("attn.ln")
    qkv_sec = (nq * hd, nkv * hd, nkv * hd)
    out.append(P("attn.qkv.weight", (sum(qkv_sec), H), 0, qkv_sec, (1, 2)))
    if cfg.use_attention_bias:
        out.append(P("attn.qkv.bias", (sum(qkv_sec),), 0, qkv_sec, (1, 2), init="zeros"))
    out.append(P("attn.o.weight", (H, nq * hd), 1))
    if cfg.use_attn_proj_bias:
        out.append(P("attn.o.bias", (H,), None, init="zeros"))
    norm("mlp.ln")
    if cfg.mlp_type == "llama":
        out.append(P("mlp.gate_up.weight", (2 * F, H), 0, (F, F)))
        out.append(P("mlp.down.weight", (H, F), 1))
    elif cfg.mlp_type == "moe":
        E = cfg.moe.num_experts
        out.append(P("mlp.router.weight", (E, H), None))
        if cfg.moe.expert_parallel:  # whole experts per rank: split along the expert dim
            out.append(P("mlp.experts.gate_up.weight", (E, 2 * F, H), 0))
            out.append(P("mlp.experts.down.weight", (E, H, F), 0))
        else:                        # every rank holds all experts, FFN dim sliced (the reference's layout)
            out.append(P("mlp.experts.gate_up.weight", (E, 2 * F, H), 1, expert_dim=True))
            out.append(P("mlp.experts.down.weight", (E, H, F), 2, expert_dim=True))
    else:
        out.append(P("mlp.fc.weight", (F, H), 0))
        if cfg.use_mlp_bias:
            out.append(P("mlp.fc.bias", (F,), 0, init="zeros"))
        out.append(P("mlp.proj.weight", (H, F), 1))
        if cfg.use_mlp_bias:
            out.append(P("mlp.proj.bias", (H,), None, init="zeros"))
    if layer_idx == L:
        norm("ln_f")
    return out


def needs_tied_head_copy(cfg: ReaLModelConfig, layers: Sequence[int]) -> bool:
    layers = list(layers)
    return bool(cfg.tied_embedding) and (cfg.n_layers + 1) in layers and 0 not in layers


def model_param_specs(cfg: ReaLModelConfig, layers: Optional[Sequence[int]] = None) -> List[ParamSpec]:
    layers = list(range(cfg.n_layers + 2) if layers is None else layers)
    tied = needs_tied_head_copy(cfg, layers)
    return [s for i in layers for s in layer_param_specs(cfg, i, tied_head_copy=tied)]


# ------------------------------------------------------------------------------------------- TP shard geometry


def _section_slices(spec: ParamSpec, cfg: ReaLModelConfig, tp_rank: int, tp_size: int) -> List[Tuple[int, int]]:
    """Row ranges (in the unsharded dim 0) that TP rank `tp_rank` owns for a section-split tensor."""
    out, off = [], 0
    hd = cfg.head_dim
    for si, sec in enumerate(spec.sections):
        if si in spec.kv_sections and cfg.n_kv_heads < tp_size:
            # fewer KV heads than TP ranks: each rank holds ONE (replicated) KV head
            assert tp_size % cfg.n_kv_heads == 0, "tp must be a multiple of n_kv_heads when larger"
            head = tp_rank * cfg.n_kv_heads // tp_size
            out.append((off + head * hd, off + (head + 1) * hd))
        else:
            assert sec % tp_size == 0, f"{spec.name}: section {sec} not divisible by tp={tp_size}"
            n = sec // tp_size
            out.append((off + tp_rank * n, off + (tp_rank + 1) * n))
        off += sec
    return out


def shard_row_ranges(spec: ParamSpec, cfg: ReaLModelConfig, tp_rank: int, tp_size: int) -> Optional[List[Tuple[int, int]]]:
    """For dim-0-split tensors: list of [a,b) row ranges of the full tensor forming the shard, in order."""
    if spec.split_dim != 0 or tp_size == 1:
        return None
    if spec.sections:
        return _section_slices(spec, cfg, tp_rank, tp_size)
    n = spec.shape[0]
    if spec.name.endswith("wte.weight") or spec.name.endswith("head.weight"):
        assert n % tp_size == 0, f"vocab size {n} must be divisible by tp={tp_size}"
    per = n // tp_size
    return [(tp_rank * per, (tp_rank + 1) * per)]


def shard_shape(spec: ParamSpec, cfg: ReaLModelConfig, tp_size: int) -> Tuple[int, ...]:
    if spec.split_dim is None or tp_size == 1:
        return spec.shape
    shp = list(spec.shape)
    if spec.split_dim == 0:
        shp[0] = sum(b - a for a, b in shard_row_ranges(spec, cfg, 0, tp_size))
    elif spec.sections and spec.split_dim == 1 and spec.expert_dim:
        shp[1] //= tp_size
    else:
        assert shp[spec.split_dim] % tp_size == 0, (spec.name, shp, tp_size)
        shp[spec.split_dim] //= tp_size
    return tuple(shp)


def shard_tensor(spec: ParamSpec, cfg: ReaLModelConfig, full: torch.Tensor, tp_rank: int, tp_size: int) -> torch.Tensor:
    if spec.split_dim is None or tp_size == 1:
        return full
    if spec.split_dim == 0:
        return torch.cat([full[a:b] for a, b in shard_row_ranges(spec, cfg, tp_rank, tp_size)], dim=0)
    if spec.expert_dim and spec.split_dim == 1:  # [E, 2F, H]: gate|up sections along dim 1
        F = full.shape[1] // 2
        n = F // tp_size
        return torch.cat([full[:, tp_rank * n:(tp_rank + 1) * n], full[:, F + tp_rank * n:F + (tp_rank + 1) * n]], dim=1)
    n = full.shape[spec.split_dim] // tp_size
    return full.narrow(spec.split_dim, tp_rank * n, n)


def merge_shards(spec: ParamSpec, cfg: ReaLModelConfig, shards: List[torch.Tensor]) -> torch.Tensor:
    """Inverse of `shard_tensor` over all TP ranks (replicated KV heads are de-duplicated)."""
    t = len(shards)
    if spec.split_dim is None or t == 1:
        return shards[0]
    if spec.split_dim == 0:
        full = torch.empty(spec.shape, dtype=shards[0].dtype, device=shards[0].device)
        for r, sh in enumerate(shards):
            off = 0
            for a, b in shard_row_ranges(spec, cfg, r, t):
                full[a:b] = sh[off:off + (b - a)]
                off += b - a
        return full
    if spec.expert_dim and spec.split_dim == 1:
        n = shards[0].shape[1] // 2
        return torch.cat([s[:, :n] for s in shards] + [s[:, n:] for s in shards], dim=1)
    return torch.cat(shards, dim=spec.split_dim)


def shard_intervals(spec: ParamSpec, cfg: ReaLModelConfig, tp_rank: int, tp_size: int) -> List[Tuple[int, int]]:
    """Element intervals [a,b) of the *flattened full tensor* that make up the shard, in shard order.
    This is what parameter reallocation turns into copy segments."""
    numel = 1
    for d in spec.shape:
        numel *= d
    if spec.split_dim is None or tp_size == 1:
        return [(0, numel)]
    if spec.split_dim == 0:
        inner = numel // spec.shape[0]
        return [(a * inner, b * inner) for a, b in shard_row_ranges(spec, cfg, tp_rank, tp_size)]
    shp = spec.shape
    if len(shp) == 2:  # [rows, cols] split along cols: one interval per row
        n = shp[1] // tp_size
        return [(r * shp[1] + tp_rank * n, r * shp[1] + (tp_rank + 1) * n) for r in range(shp[0])]
    if spec.expert_dim and spec.split_dim == 1:  # [E, 2F, H]
        E, F2, H = shp
        F = F2 // 2
        n = F // tp_size
        out = []
        for e in range(E):
            base = e * F2 * H
            out.append((base + tp_rank * n * H, base + (tp_rank + 1) * n * H))
            out.append((base + (F + tp_rank * n) * H, base + (F + (tp_rank + 1) * n) * H))
        return out
    if spec.expert_dim and spec.split_dim == 2:  # [E, H, F]
        E, H, F = shp
        n = F // tp_size
        return [((e * H + h) * F + tp_rank * n, (e * H + h) * F + (tp_rank + 1) * n) for e in range(E) for h in range(H)]
    raise NotImplementedError(spec)


# ------------------------------------------------------------------------------------------- PP partition


def layer_param_count(cfg: ReaLModelConfig, layer_idx: int) -> int:
    n = 0
    for s in layer_param_specs(cfg, layer_idx):
        k = 1
        for d in s.shape:
            k *= d
        n += k
    return n


def partition_pipeline_layers(cfg: ReaLModelConfig, num_stages: int) -> Dict[int, Tuple[int, int]]:
    """stage -> [start, end) over layer indices 0..L+1, balanced by parameter count (embedding and head
    count as layers, like the reference `partition_pipeline_layers`)."""
    from realhf_b200.base.datapack import partition_balanced
    n_layers = cfg.n_layers + 2
    if num_stages == 1:
        return {0: (0, n_layers)}
    assert num_stages <= cfg.n_layers, "more pipeline stages than transformer blocks"
    counts = [max(1, layer_param_count(cfg, i)) for i in range(n_layers)]
    # the embedding must share a stage with >=1 block and so must the head: pin them to their neighbours
    w = counts[1:-1]
    w[0] += counts[0]
    w[-1] += counts[-1]
    b = partition_balanced(w, num_stages, 1)
    out = {}
    for s in range(num_stages):
        lo, hi = b[s] + 1, b[s + 1] + 1
        if s == 0:
            lo = 0
        if s == num_stages - 1:
            hi = n_layers
        out[s] = (lo, hi)
    return out
