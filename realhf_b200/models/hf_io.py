"""HuggingFace-format checkpoint load / save for any (pp, tp, dp) layout.

Parity: `realhf/impl/model/conversion/hf_registry.py` (load :62-199, save :201-365).  Format = a HF
directory: config.json + `model-XXXXX-of-YYYYY.safetensors` + index (the reference writes pickled `.bin`;
we write safetensors and read both).  Load: each rank opens only the files holding its pipeline stage's
layers, converts names, TP-slices.  Save: TP shards are gathered on tp-rank 0 of each stage (over NCCL
or gloo), merged, converted to HF names, split into size-bounded shards and written by the dp-rank-0
ranks; the first stage writes config and index.  Critic heads are saved as a `[1, hidden]` `lm_head.weight`
like the reference (such checkpoints load here, not in HF).
"""

from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from realhf_b200.api.model import SUPPORTED_HF_FAMILIES, HFFamilySpec, ReaLModelConfig
from realhf_b200.base.topology import ParallelContext
from realhf_b200.models import sharding
from realhf_b200.models.real_model import ReaLModel

MAX_SHARD_BYTES = int(os.environ.get("REAL_SAVE_MAX_SHARD_SIZE_BYTE", int(1e10)))


def family(name: str) -> HFFamilySpec:
    import realhf_b200.api.from_hf  # noqa: F401  (fills the registry)
    return SUPPORTED_HF_FAMILIES[name]


def load_hf_config(path: str):
    import transformers
    return transformers.AutoConfig.from_pretrained(path, trust_remote_code=True)


def _config_from_saved_json(fn: str, family_name: str, is_critic: bool) -> Optional[ReaLModelConfig]:
    """The `ReaLModelConfig` this framework wrote next to `config.json` when it saved the checkpoint, if it describes the same
    family and head type as requested; None otherwise (older files, foreign keys, critic-from-actor initialisation)."""
    import dataclasses

    from realhf_b200.api.model import ReaLMoEConfig
    try:
        with open(fn) as f:
            d = json.load(f)
        if d.pop("_family", None) != family_name or bool(d.get("is_critic", False)) != bool(is_critic):
            return None
        names = {f.name for f in dataclasses.fields(ReaLModelConfig)}
        if set(d) - names:
            return None
        if isinstance(d.get("moe"), dict):
            d["moe"] = ReaLMoEConfig(**d["moe"])
        return ReaLModelConfig(**d)
    except (OSError, ValueError, TypeError):
        return None


def config_from_hf_path(family_name: str, path: str, is_critic: bool = False) -> ReaLModelConfig:
    """Model config of a checkpoint directory.  Checkpoints written by this framework carry `real_model_config.json` (the exact
    config that produced the weights): reading it avoids instantiating the HuggingFace config class, whose first import costs
    every worker process several seconds (`transformers` pulls `torch._dynamo` in).  `REAL_TRUST_SAVED_MODEL_CONFIG=0` always goes
    through `config.json`; so do foreign checkpoints and a critic initialised from an actor checkpoint."""
    fast = os.path.join(path, "real_model_config.json")
    if os.environ.get("REAL_TRUST_SAVED_MODEL_CONFIG", "1") == "1" and os.path.exists(fast):
        cfg = _config_from_saved_json(fast, family_name, is_critic)
        if cfg is not None:
            return cfg
    return family(family_name).config_from_hf(load_hf_config(path), is_critic)


# ------------------------------------------------------------------------------------------- reading


def _weight_map(path: str) -> Dict[str, str]:
    for idx in ("model.safetensors.index.json", "pytorch_model.bin.index.json"):
        f = os.path.join(path, idx)
        if os.path.exists(f):
            with open(f) as fh:
                return json.load(fh)["weight_map"]
    for single in ("model.safetensors", "pytorch_model.bin"):
        if os.path.exists(os.path.join(path, single)):
            return {"*": single}
    raise FileNotFoundError(f"no HF weights under {path}")


from realhf_b200.base.saveload_utils import load_weight_file as _read_file  # noqa: E402
from realhf_b200.base.saveload_utils import split_state_dict_into_shards as _split_into_files  # noqa: E402


def load_hf_state_dict(path: str, names: Optional[List[str]] = None) -> Dict[str, torch.Tensor]:
    """HF tensors by name; with `names`, opens only the files that hold them."""
    wm = _weight_map(path)
    if "*" in wm:
        sd = _read_file(os.path.join(path, wm["*"]))
        return sd if names is None else {k: sd[k] for k in names if k in sd}
    files = sorted(set(wm.values()) if names is None else {wm[n] for n in names if n in wm})
    out: Dict[str, torch.Tensor] = {}
    for f in files:
        sd = _read_file(os.path.join(path, f))
        out.update(sd if names is None else {k: v for k, v in sd.items() if k in set(names)})
    return out


def load_from_hf(model: ReaLModel, family_name: str, path: str, init_critic_from_actor: bool = False):
    """Fill an instantiated (or empty) model shard from an HF directory."""
    fam = family(family_name)
    cfg = model.config
    names: List[str] = []
    for li in model.layers:
        if li == 0:
            names += fam.embedding_param_names(cfg)
        elif li <= cfg.n_layers:
            names += fam.tblock_param_names(cfg, li - 1)
        else:
            names += fam.head_param_names(cfg)
    if cfg.tied_embedding and model.is_last_stage and not model.is_first_stage:
        names += fam.embedding_param_names(cfg)
    hf_sd = load_hf_state_dict(path, names)
    if init_critic_from_actor and "lm_head.weight" in hf_sd:
        hf_sd.pop("lm_head.weight")
    sd = fam.sd_from_hf(hf_sd, cfg)
    if not model.instantiated:
        model.instantiate(init="empty")
    missing = []
    with torch.no_grad():
        tied_copy = sharding.needs_tied_head_copy(cfg, model.layers)
        for name, slot in model.slots.items():
            if name not in sd and tied_copy and name.endswith("head.weight") and "0.wte.weight" in sd:
                sd[name] = sd["0.wte.weight"]
            if name not in sd:
                missing.append(name)
                continue
            full = sd[name]
            assert tuple(full.shape) == tuple(slot.spec.shape), (name, full.shape, slot.spec.shape)
            sh = sharding.shard_tensor(slot.spec, cfg, full, model.ctx.tp_rank, model.ctx.tp_size)
            model.p[name].copy_(sh.to(model.dtype))
    head_name = f"{cfg.n_layers + 1}.head.weight"
    if missing:
        if missing == [head_name] and (init_critic_from_actor or cfg.is_critic):
            with torch.no_grad():
                model.p[head_name].normal_(0.0, 0.02)
        else:
            raise KeyError(f"checkpoint {path} lacks parameters {missing}")
    return model


# ------------------------------------------------------------------------------------------- writing


def gather_full_state_dict(model: ReaLModel) -> Optional[Dict[str, torch.Tensor]]:
    """Merge TP shards of this stage on tp-rank 0 (CPU tensors).  Other ranks return None."""
    ctx = model.ctx
    out: Dict[str, torch.Tensor] = {}
    for name, slot in model.slots.items():
        t = model.p[name].data
        if ctx.tp_size == 1 or slot.spec.split_dim is None:
            if ctx.tp_rank == 0:
                out[name] = t.detach().cpu()
            continue
        parts = [torch.empty_like(t) for _ in range(ctx.tp_size)]
        dist.all_gather(parts, t.contiguous(), group=ctx.tp_group)
        if ctx.tp_rank == 0:
            out[name] = sharding.merge_shards(slot.spec, model.config, [p.cpu() for p in parts])
    return out if ctx.tp_rank == 0 else None


def save_to_hf(model: ReaLModel, family_name: str, save_dir: str, tokenizer=None, max_shard_bytes: Optional[int] = None):
    """Collective over the model's ranks."""
    from safetensors.torch import save_file
    fam = family(family_name)
    cfg, ctx = model.config, model.ctx
    os.makedirs(save_dir, exist_ok=True)
    full = gather_full_state_dict(model)
    my_files: List[Tuple[str, List[str]]] = []
    n_mine = 0
    if full is not None and ctx.dp_rank == 0:
        hf_sd = fam.sd_to_hf(full, cfg)
        if cfg.tied_embedding and "lm_head.weight" in hf_sd:
            hf_sd.pop("lm_head.weight")
        chunks = _split_into_files(hf_sd, max_shard_bytes or MAX_SHARD_BYTES)
        n_mine = len(chunks)
    # every stage learns how many files the others write so names are globally consistent
    counts = [n_mine]
    if ctx.pp_size > 1:
        obj = [None] * ctx.pp_size
        dist.all_gather_object(obj, n_mine if (ctx.tp_rank == 0 and ctx.dp_rank == 0) else -1, group=ctx.pp_group)
        counts = obj
    if full is not None and ctx.dp_rank == 0:
        total = sum(c for c in counts if c > 0)
        base = sum(c for c in counts[: ctx.pp_rank] if c > 0)
        for j, chunk in enumerate(chunks):
            fn = f"model-{base + j + 1:05d}-of-{total:05d}.safetensors"
            save_file({k: v.contiguous() for k, v in chunk.items()}, os.path.join(save_dir, fn), metadata={"format": "pt"})
            my_files.append((fn, list(chunk.keys())))
    maps = [my_files]
    if ctx.pp_size > 1:
        obj = [None] * ctx.pp_size
        dist.all_gather_object(obj, my_files, group=ctx.pp_group)
        maps = obj
    if ctx.local_rank == ctx.topo.get_rank(pipe=0, data=0, model=0):
        weight_map = {k: fn for files in maps for fn, keys in files for k in keys}
        with open(os.path.join(save_dir, "model.safetensors.index.json"), "w") as f:
            json.dump({"metadata": {}, "weight_map": weight_map}, f, indent=1)
        hf_cfg = fam.config_to_hf(cfg)
        hf_cfg.architectures = [fam.hf_cls_name]
        hf_cfg.save_pretrained(save_dir)
        with open(os.path.join(save_dir, "real_model_config.json"), "w") as f:
            import dataclasses
            json.dump(dict(dataclasses.asdict(cfg), _family=family_name), f, indent=1)
        if tokenizer is not None:
            tokenizer.save_pretrained(save_dir)
    if dist.is_initialized() and ctx.model_group is not None:
        dist.barrier(group=ctx.model_group)


# ------------------------------------------------------------------------------------------- convenience (pp=tp=1)


def from_hf(family_name: str, path: str, is_critic: bool = False, init_critic_from_actor: bool = False, dtype=torch.bfloat16,
            device="cpu", ctx: Optional[ParallelContext] = None) -> ReaLModel:
    cfg = config_from_hf_path(family_name, path, is_critic)
    model = ReaLModel(cfg, ctx, dtype=dtype, device=device)
    return load_from_hf(model, family_name, path, init_critic_from_actor)


def to_hf_model(model: ReaLModel, family_name: str):
    """In-memory transformers model with the same weights (single-shard models)."""
    import transformers
    fam = family(family_name)
    hf_cfg = fam.config_to_hf(model.config)
    hf = getattr(transformers, fam.hf_cls_name)(hf_cfg)
    sd = fam.sd_to_hf({k: v.detach().float().cpu() for k, v in model.state_dict().items()}, model.config)
    hf_keys = set(hf.state_dict().keys())
    if any(k.endswith("mlp.experts.gate_up_proj") for k in hf_keys):  # transformers>=5 keeps MoE experts fused in memory
        sd = _fuse_moe_keys(sd, model.config)
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    bad = [k for k in missing if "rotary" not in k and "masked_bias" not in k and not k.endswith("attn.bias") and k != "lm_head.weight"]
    assert not bad and not unexpected, (bad, unexpected)
    return hf


def _fuse_moe_keys(sd: Dict[str, torch.Tensor], cfg: ReaLModelConfig) -> Dict[str, torch.Tensor]:
    """Classic Mixtral names (block_sparse_moe.experts.{e}.w{1,2,3}) -> fused in-memory names of newer transformers."""
    out = {k: v for k, v in sd.items() if "block_sparse_moe" not in k}
    E = cfg.moe.num_experts
    for i in range(cfg.n_layers):
        hp = f"model.layers.{i}.block_sparse_moe."
        if hp + "gate.weight" not in sd:
            continue
        np_ = f"model.layers.{i}.mlp."
        out[np_ + "gate.weight"] = sd[hp + "gate.weight"]
        out[np_ + "experts.gate_up_proj"] = torch.stack(
            [torch.cat([sd[hp + f"experts.{e}.w1.weight"], sd[hp + f"experts.{e}.w3.weight"]], 0) for e in range(E)])
        out[np_ + "experts.down_proj"] = torch.stack([sd[hp + f"experts.{e}.w2.weight"] for e in range(E)])
    return out
