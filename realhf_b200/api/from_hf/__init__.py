"""HuggingFace <-> ReaLModel conversion specs for llama / codellama / deepseek, qwen2, mistral, gemma, mixtral, gpt2.

Parity: `realhf/api/from_hf/{llama,gpt2,gemma,mistral,mixtral,qwen2}.py`.  Every family registers
config converters (both directions), state-dict converters (both directions; HF q/k/v and gate/up are
fused into this framework's `attn.qkv` and `mlp.gate_up` tensors) and per-layer HF parameter-name lists
(used to open only the checkpoint shards a pipeline stage needs).
"""

from __future__ import annotations

from typing import Dict, List

import torch

from realhf_b200.api.model import ReaLModelConfig, ReaLMoEConfig, register_hf_family

# ------------------------------------------------------------------------------------------- llama-like


def _llama_like_config_from_hf(hf, is_critic: bool, *, qkv_bias=False, norm_type="rms", act_default="silu",
                               normalize_embed=False, tied=None, moe=False, sliding_window=None) -> ReaLModelConfig:
    rs = getattr(hf, "rope_scaling", None) or {}
    rp = getattr(hf, "rope_parameters", None) or {}
    base = getattr(hf, "rope_theta", None) or rp.get("rope_theta", 10000.0)
    rtype = rs.get("type", rs.get("rope_type", rp.get("rope_type", None)))
    if rtype in ("default", None):
        rtype, factor = None, None
    else:
        factor = rs.get("factor", rp.get("factor", None))
    tied_emb = bool(getattr(hf, "tie_word_embeddings", False)) if tied is None else tied
    act = getattr(hf, "hidden_act", None) or getattr(hf, "hidden_activation", None) or act_default
    cfg = ReaLModelConfig(
        n_layers=hf.num_hidden_layers, n_kv_heads=getattr(hf, "num_key_value_heads", hf.num_attention_heads),
        n_q_heads=hf.num_attention_heads, hidden_dim=hf.hidden_size, intermediate_dim=hf.intermediate_size,
        vocab_size=hf.vocab_size, head_dim=getattr(hf, "head_dim", None) or hf.hidden_size // hf.num_attention_heads,
        n_positions=getattr(hf, "max_position_embeddings", None), embd_pdrop=0.0, resid_pdrop=0.0,
        attn_pdrop=getattr(hf, "attention_dropout", 0.0) or 0.0, layer_norm_epsilon=hf.rms_norm_eps,
        activation_function=act, scale_attn_by_inverse_layer_idx=False, scale_attn_weights=True,
        use_attention_bias=qkv_bias, use_attn_proj_bias=False, use_mlp_bias=False, layer_norm_type=norm_type,
        mlp_type="moe" if moe else "llama", apply_rotary=True, rotary_base=float(base), rotary_interleaved=False,
        rotary_scaling=factor, rotary_scaling_type=rtype, normalize_embed=normalize_embed,
        tied_embedding=False if is_critic else tied_emb, sliding_window=sliding_window, is_critic=is_critic)
    if moe:
        cfg.moe = ReaLMoEConfig(num_experts=hf.num_local_experts, top_k=hf.num_experts_per_tok, routing_type="aux_loss",
                                aux_loss_coeff=getattr(hf, "router_aux_loss_coef", 1e-3), capacity_factor=None,
                                input_jitter_eps=getattr(hf, "router_jitter_noise", 0.0) or 0.0)
    return cfg


def _llama_like_config_to_hf(cfg: ReaLModelConfig, hf_cls_name: str, **extra):
    import transformers
    cls = getattr(transformers, hf_cls_name)
    kw = dict(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_dim, intermediate_size=cfg.intermediate_dim,
              num_hidden_layers=cfg.n_layers, num_attention_heads=cfg.n_q_heads, num_key_value_heads=cfg.n_kv_heads,
              max_position_embeddings=cfg.n_positions or 4096, rms_norm_eps=cfg.layer_norm_epsilon,
              rope_theta=cfg.rotary_base, tie_word_embeddings=cfg.tied_embedding, attention_dropout=cfg.attn_pdrop)
    if cfg.rotary_scaling_type is not None:
        kw["rope_scaling"] = dict(type=cfg.rotary_scaling_type, rope_type=cfg.rotary_scaling_type, factor=cfg.rotary_scaling)
    kw.update(extra)
    return cls(**kw)


def _llama_like_sd_from_hf(sd: Dict[str, torch.Tensor], cfg: ReaLModelConfig, *, prefix="model.", moe=False) -> Dict[str, torch.Tensor]:
    """HF names -> this framework's names; tolerant of partial dicts (only the layers present are converted)."""
    out: Dict[str, torch.Tensor] = {}
    L = cfg.n_layers
    g = sd.get
    if f"{prefix}embed_tokens.weight" in sd:
        out["0.wte.weight"] = sd[f"{prefix}embed_tokens.weight"]
    for i in range(L):
        hp, rp = f"{prefix}layers.{i}.", f"{i + 1}."
        if f"{hp}self_attn.q_proj.weight" not in sd:
            continue
        out[rp + "attn.ln.weight"] = sd[hp + "input_layernorm.weight"]
        out[rp + "attn.qkv.weight"] = torch.cat([sd[hp + f"self_attn.{x}_proj.weight"] for x in "qkv"], dim=0)
        if cfg.use_attention_bias:
            out[rp + "attn.qkv.bias"] = torch.cat([sd[hp + f"self_attn.{x}_proj.bias"] for x in "qkv"], dim=0)
        out[rp + "attn.o.weight"] = sd[hp + "self_attn.o_proj.weight"]
        out[rp + "mlp.ln.weight"] = sd[hp + "post_attention_layernorm.weight"]
        if moe:
            E = cfg.moe.num_experts
            mp = hp + ("block_sparse_moe." if (hp + "block_sparse_moe.gate.weight") in sd else "mlp.")
            out[rp + "mlp.router.weight"] = sd[mp + "gate.weight"]
            if (mp + "experts.0.w1.weight") in sd:
                out[rp + "mlp.experts.gate_up.weight"] = torch.stack(
                    [torch.cat([sd[mp + f"experts.{e}.w1.weight"], sd[mp + f"experts.{e}.w3.weight"]], 0) for e in range(E)])
                out[rp + "mlp.experts.down.weight"] = torch.stack([sd[mp + f"experts.{e}.w2.weight"] for e in range(E)])
            else:  # fused expert tensors (newer transformers state dicts)
                out[rp + "mlp.experts.gate_up.weight"] = sd[mp + "experts.gate_up_proj"]
                out[rp + "mlp.experts.down.weight"] = sd[mp + "experts.down_proj"]
        else:
            out[rp + "mlp.gate_up.weight"] = torch.cat([sd[hp + "mlp.gate_proj.weight"], sd[hp + "mlp.up_proj.weight"]], dim=0)
            out[rp + "mlp.down.weight"] = sd[hp + "mlp.down_proj.weight"]
        if i == L - 1 and f"{prefix}norm.weight" in sd:
            out[rp + "ln_f.weight"] = sd[f"{prefix}norm.weight"]
    if "lm_head.weight" in sd and not cfg.tied_embedding:
        out[f"{L + 1}.head.weight"] = sd["lm_head.weight"]
    return out


def _llama_like_sd_to_hf(sd: Dict[str, torch.Tensor], cfg: ReaLModelConfig, *, prefix="model.", moe=False) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    L, hd = cfg.n_layers, cfg.head_dim
    nq, nkv, F = cfg.n_q_heads * hd, cfg.n_kv_heads * hd, cfg.intermediate_dim
    for k, v in sd.items():
        li, name = k.split(".", 1)
        li = int(li)
        if li == 0 and name == "wte.weight":
            out[f"{prefix}embed_tokens.weight"] = v
        elif li == L + 1:
            out["lm_head.weight"] = v
        else:
            hp = f"{prefix}layers.{li - 1}."
            if name == "attn.ln.weight":
                out[hp + "input_layernorm.weight"] = v
            elif name in ("attn.qkv.weight", "attn.qkv.bias"):
                kind = name.rsplit(".", 1)[1]
                q, kk, vv = torch.split(v, [nq, nkv, nkv], dim=0)
                out[hp + f"self_attn.q_proj.{kind}"], out[hp + f"self_attn.k_proj.{kind}"], out[hp + f"self_attn.v_proj.{kind}"] = q, kk, vv
            elif name == "attn.o.weight":
                out[hp + "self_attn.o_proj.weight"] = v
            elif name == "mlp.ln.weight":
                out[hp + "post_attention_layernorm.weight"] = v
            elif name == "mlp.gate_up.weight":
                out[hp + "mlp.gate_proj.weight"], out[hp + "mlp.up_proj.weight"] = torch.split(v, [F, F], dim=0)
            elif name == "mlp.down.weight":
                out[hp + "mlp.down_proj.weight"] = v
            elif name == "mlp.router.weight":
                out[hp + "block_sparse_moe.gate.weight"] = v
            elif name == "mlp.experts.gate_up.weight":
                for e in range(v.shape[0]):
                    out[hp + f"block_sparse_moe.experts.{e}.w1.weight"], out[hp + f"block_sparse_moe.experts.{e}.w3.weight"] = \
                        torch.split(v[e], [F, F], dim=0)
            elif name == "mlp.experts.down.weight":
                for e in range(v.shape[0]):
                    out[hp + f"block_sparse_moe.experts.{e}.w2.weight"] = v[e]
            elif name == "ln_f.weight":
                out[f"{prefix}norm.weight"] = v
            else:
                raise KeyError(f"unknown parameter {k}")
    return {k: v.contiguous() for k, v in out.items()}


def _llama_like_names(prefix="model.", qkv_bias=False, moe=False):
    def embedding(cfg) -> List[str]:
        return [f"{prefix}embed_tokens.weight"]

    def tblock(cfg, idx: int) -> List[str]:
        hp = f"{prefix}layers.{idx}."
        names = [hp + "input_layernorm.weight", hp + "post_attention_layernorm.weight", hp + "self_attn.o_proj.weight"]
        for x in "qkv":
            names.append(hp + f"self_attn.{x}_proj.weight")
            if qkv_bias:
                names.append(hp + f"self_attn.{x}_proj.bias")
        if moe:
            names.append(hp + "block_sparse_moe.gate.weight")
            for e in range(cfg.moe.num_experts):
                names += [hp + f"block_sparse_moe.experts.{e}.w{j}.weight" for j in (1, 2, 3)]
        else:
            names += [hp + f"mlp.{x}_proj.weight" for x in ("gate", "up", "down")]
        if idx == cfg.n_layers - 1:
            names.append(f"{prefix}norm.weight")
        return names

    def head(cfg) -> List[str]:
        return [] if cfg.tied_embedding else ["lm_head.weight"]

    return embedding, tblock, head


def _tiny_llama_like(**kw) -> ReaLModelConfig:
    base = dict(n_layers=4, n_kv_heads=4, n_q_heads=8, hidden_dim=64, intermediate_dim=128, vocab_size=128, head_dim=8,
                n_positions=512, embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, layer_norm_epsilon=1e-5,
                activation_function="silu", scale_attn_by_inverse_layer_idx=False, use_attention_bias=False,
                use_attn_proj_bias=False, use_mlp_bias=False, layer_norm_type="rms", mlp_type="llama", apply_rotary=True)
    base.update(kw)
    return ReaLModelConfig(**base)


def _register_llama_like(name, hf_cls_name, *, qkv_bias=False, norm_type="rms", normalize_embed=False, tied=None, moe=False,
                         act_default="silu", to_hf_extra=None, tiny_kw=None):
    emb, blk, head = _llama_like_names(qkv_bias=qkv_bias, moe=moe)
    extra_fn = to_hf_extra or (lambda cfg: {})
    register_hf_family(
        name=name, hf_cls_name=hf_cls_name,
        config_from_hf_converter=lambda hf, is_critic=False: _llama_like_config_from_hf(
            hf, is_critic, qkv_bias=qkv_bias, norm_type=norm_type, normalize_embed=normalize_embed, tied=tied, moe=moe,
            act_default=act_default, sliding_window=getattr(hf, "sliding_window", None)),
        config_to_hf_converter=lambda cfg: _llama_like_config_to_hf(cfg, hf_cls_name.replace("ForCausalLM", "Config"), **extra_fn(cfg)),
        sd_from_hf_converter=lambda sd, cfg: _llama_like_sd_from_hf(sd, cfg, moe=moe),
        sd_to_hf_converter=lambda sd, cfg: _llama_like_sd_to_hf(sd, cfg, moe=moe),
        embedding_param_names=emb, tblock_param_names=blk, head_param_names=head,
        make_test_config=lambda: _tiny_llama_like(**(tiny_kw or {})))


for _n in ("llama", "codellama", "deepseek"):
    _register_llama_like(_n, "LlamaForCausalLM")
_register_llama_like("qwen2", "Qwen2ForCausalLM", qkv_bias=True, tiny_kw=dict(use_attention_bias=True),
                     to_hf_extra=lambda cfg: dict(use_sliding_window=False))
_register_llama_like("mistral", "MistralForCausalLM", to_hf_extra=lambda cfg: dict(sliding_window=cfg.sliding_window))
_register_llama_like(
    "gemma", "GemmaForCausalLM", norm_type="gemma", normalize_embed=True, tied=True, act_default="gelu_pytorch_tanh",
    to_hf_extra=lambda cfg: dict(head_dim=cfg.head_dim, hidden_act="gelu_pytorch_tanh", hidden_activation="gelu_pytorch_tanh"),
    tiny_kw=dict(layer_norm_type="gemma", normalize_embed=True, tied_embedding=True, activation_function="gelu_pytorch_tanh",
                 layer_norm_epsilon=1e-6))
_register_llama_like(
    "mixtral", "MixtralForCausalLM", moe=True,
    to_hf_extra=lambda cfg: dict(num_local_experts=cfg.moe.num_experts, num_experts_per_tok=cfg.moe.top_k,
                                 router_aux_loss_coef=cfg.moe.aux_loss_coeff, sliding_window=cfg.sliding_window),
    tiny_kw=dict(mlp_type="moe", moe=ReaLMoEConfig(num_experts=4, top_k=2, capacity_factor=None)))

# ------------------------------------------------------------------------------------------- gpt2


def _gpt2_config_from_hf(hf, is_critic=False) -> ReaLModelConfig:
    return ReaLModelConfig(
        n_layers=hf.n_layer, n_kv_heads=hf.n_head, n_q_heads=hf.n_head, hidden_dim=hf.n_embd,
        intermediate_dim=hf.n_inner if hf.n_inner is not None else 4 * hf.n_embd, vocab_size=hf.vocab_size,
        n_positions=hf.n_positions, embd_pdrop=hf.embd_pdrop, resid_pdrop=hf.resid_pdrop, attn_pdrop=hf.attn_pdrop,
        layer_norm_epsilon=hf.layer_norm_epsilon, activation_function=hf.activation_function,
        scale_attn_by_inverse_layer_idx=bool(hf.scale_attn_by_inverse_layer_idx), scale_attn_weights=bool(hf.scale_attn_weights),
        use_attention_bias=True, use_attn_proj_bias=True, use_mlp_bias=True, layer_norm_type=None, mlp_type=None,
        apply_rotary=False, tied_embedding=not is_critic, is_critic=is_critic)


def _gpt2_config_to_hf(cfg: ReaLModelConfig):
    import transformers
    return transformers.GPT2Config(
        vocab_size=cfg.vocab_size, n_positions=cfg.n_positions, n_embd=cfg.hidden_dim, n_layer=cfg.n_layers,
        n_head=cfg.n_q_heads, n_inner=cfg.intermediate_dim, activation_function=cfg.activation_function,
        resid_pdrop=cfg.resid_pdrop, embd_pdrop=cfg.embd_pdrop, attn_pdrop=cfg.attn_pdrop,
        layer_norm_epsilon=cfg.layer_norm_epsilon, scale_attn_by_inverse_layer_idx=cfg.scale_attn_by_inverse_layer_idx,
        scale_attn_weights=cfg.scale_attn_weights, tie_word_embeddings=cfg.tied_embedding)


_GPT2_MAP = [("ln_1.weight", "attn.ln.weight", False), ("ln_1.bias", "attn.ln.bias", False),
             ("attn.c_attn.weight", "attn.qkv.weight", True), ("attn.c_attn.bias", "attn.qkv.bias", False),
             ("attn.c_proj.weight", "attn.o.weight", True), ("attn.c_proj.bias", "attn.o.bias", False),
             ("ln_2.weight", "mlp.ln.weight", False), ("ln_2.bias", "mlp.ln.bias", False),
             ("mlp.c_fc.weight", "mlp.fc.weight", True), ("mlp.c_fc.bias", "mlp.fc.bias", False),
             ("mlp.c_proj.weight", "mlp.proj.weight", True), ("mlp.c_proj.bias", "mlp.proj.bias", False)]


def _gpt2_sd_from_hf(sd, cfg):
    out = {}
    sd = {k[len("transformer."):] if k.startswith("transformer.") else k: v for k, v in sd.items()}
    if "wte.weight" in sd:
        out["0.wte.weight"] = sd["wte.weight"]
    if "wpe.weight" in sd:
        out["0.wpe.weight"] = sd["wpe.weight"]
    for i in range(cfg.n_layers):
        if f"h.{i}.ln_1.weight" not in sd:
            continue
        for hf_n, my_n, tr in _GPT2_MAP:  # HF GPT-2 uses Conv1D: weights are stored [in, out]
            v = sd[f"h.{i}.{hf_n}"]
            out[f"{i + 1}.{my_n}"] = v.t().contiguous() if tr else v
        if i == cfg.n_layers - 1 and "ln_f.weight" in sd:
            out[f"{i + 1}.ln_f.weight"], out[f"{i + 1}.ln_f.bias"] = sd["ln_f.weight"], sd["ln_f.bias"]
    if cfg.is_critic and "lm_head.weight" in sd:
        out[f"{cfg.n_layers + 1}.head.weight"] = sd["lm_head.weight"]
    return out


def _gpt2_sd_to_hf(sd, cfg):
    out = {}
    inv = {my: (hf, tr) for hf, my, tr in _GPT2_MAP}
    for k, v in sd.items():
        li, name = k.split(".", 1)
        li = int(li)
        if li == 0:
            out[f"transformer.{name}"] = v
        elif li == cfg.n_layers + 1:
            out["lm_head.weight"] = v
        elif name.startswith("ln_f"):
            out[f"transformer.{name}"] = v
        else:
            hf_n, tr = inv[name]
            out[f"transformer.h.{li - 1}.{hf_n}"] = v.t().contiguous() if tr else v
    if cfg.tied_embedding and "transformer.wte.weight" in out:
        out["lm_head.weight"] = out["transformer.wte.weight"]
    return out


register_hf_family(
    name="gpt2", hf_cls_name="GPT2LMHeadModel", config_from_hf_converter=_gpt2_config_from_hf,
    config_to_hf_converter=_gpt2_config_to_hf, sd_from_hf_converter=_gpt2_sd_from_hf, sd_to_hf_converter=_gpt2_sd_to_hf,
    embedding_param_names=lambda cfg: ["transformer.wte.weight", "transformer.wpe.weight"],
    tblock_param_names=lambda cfg, i: [f"transformer.h.{i}.{hf}" for hf, _, _ in _GPT2_MAP] +
    (["transformer.ln_f.weight", "transformer.ln_f.bias"] if i == cfg.n_layers - 1 else []),
    head_param_names=lambda cfg: ["lm_head.weight"] if cfg.is_critic else [],
    make_test_config=lambda: ReaLModelConfig(
        n_layers=4, n_kv_heads=8, n_q_heads=8, hidden_dim=64, intermediate_dim=256, vocab_size=128, n_positions=512,
        embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, activation_function="gelu_new",
        scale_attn_by_inverse_layer_idx=False, tied_embedding=True))
