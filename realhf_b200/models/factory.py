"""`real_model` factory (parity: `make_real_model`, nn/real_llm_api.py:857-894): builds the shard of a
ReaLModel that the enclosing `constants.model_scope` describes, from an HF directory or from scratch."""

from __future__ import annotations

import types
from typing import Optional

import torch

from realhf_b200.api import model as model_api
from realhf_b200.api.config import ModelName
from realhf_b200.base import constants
from realhf_b200.base.topology import ParallelContext
from realhf_b200.models import hf_io
from realhf_b200.models.real_model import ReaLModel

_DT = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32, None: torch.bfloat16}


def make_real_model(name: ModelName, device, model_path: str, is_critic: bool, init_from_scratch: bool = False,
                    init_critic_from_actor: bool = False, dtype: Optional[str] = None, hf_model_family: str = "llama",
                    config: Optional[model_api.ReaLModelConfig] = None, tokenizer=None, expert_parallel: bool = False) -> model_api.Model:
    scope = constants.current_scope() or {}
    ctx: ParallelContext = scope.get("ctx") or ParallelContext.single()
    instantiate = scope.get("instantiate", True)
    tdtype = _DT[dtype] if not isinstance(dtype, torch.dtype) else dtype
    if config is None:
        if model_path:
            config = hf_io.config_from_hf_path(hf_model_family, model_path, is_critic)
        else:
            config = hf_io.family(hf_model_family).make_test_config()
            config.is_critic = is_critic
            if is_critic:
                config.tied_embedding = False
    if expert_parallel and getattr(config, "moe", None) is not None and config.mlp_type == "moe":
        config.moe.expert_parallel = True
    m = ReaLModel(config, ctx, dtype=tdtype, device=device)
    m.hf_family = hf_model_family
    if instantiate:
        if init_from_scratch or not model_path:
            import os
            if os.environ.get("REAL_FAST_INIT", "0") == "1" and m.device.type == "cuda":
                m.init_random_fast(seed=1)  # benchmarks: device-side draw (NOT layout-invariant like `instantiate`)
            else:
                m.instantiate(seed=1)
        else:
            hf_io.load_from_hf(m, hf_model_family, model_path, init_critic_from_actor=init_critic_from_actor)
    if tokenizer is None:
        if model_path:
            try:
                from realhf_b200.api.data import load_hf_tokenizer
                tokenizer = load_hf_tokenizer(model_path)
            except Exception:
                tokenizer = None
        if tokenizer is None:
            tokenizer = types.SimpleNamespace(eos_token_id=1, pad_token_id=0, eos_token="</s>")
    model = model_api.Model(name, m, tokenizer, device, dtype=tdtype)
    model.module_config = config
    model.hf_family = hf_model_family
    return model


model_api.register_model("real_model", make_real_model)
