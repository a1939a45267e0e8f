"""Layer profiler: device-timed forward / backward / optimizer cost of embedding, block and head for given
(batch, seqlen) shapes, feeding the allocation search's cost table.

Parity: `realhf/search_engine/layers.py:56-277` + `apps/profile_layers.py`.  Timing uses CUDA events on the launching
stream after warm-up (CPU falls back to wall clock so the tool is testable without a GPU).
"""

from __future__ import annotations

import dataclasses
import json
import os
import time
from typing import Dict, List

import torch

from realhf_b200.api.model import ReaLModelConfig
from realhf_b200.base import constants
from realhf_b200.models.real_model import ReaLModel


def _time(fn, device, warmup: int = 2, iters: int = 5) -> float:
    for _ in range(warmup):
        fn()
    if device.type == "cuda":
        torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize(device)
        return a.elapsed_time(b) / iters * 1e3
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) / iters * 1e6


def profile_layers(cfg: ReaLModelConfig, batch_sizes: List[int], seqlens: List[int], device="cuda", dtype=torch.bfloat16,
                   n_blocks: int = 2) -> List[Dict]:
    """Returns rows {bs, seqlen, layer, op, time_us}; per-block numbers are averages over `n_blocks` blocks."""
    device = torch.device(device)
    small = dataclasses.replace(cfg, n_layers=n_blocks)
    m = ReaLModel(small, dtype=dtype, device=device)
    m.init_random_fast() if device.type == "cuda" else m.instantiate()
    rows = []
    for bs in batch_sizes:
        for sl in seqlens:
            T = bs * sl
            ids = torch.randint(0, cfg.vocab_size, (T,), device=device)
            cu = torch.arange(0, T + 1, sl, dtype=torch.int32, device=device)

            def fwd():
                with torch.no_grad():
                    return m(input_ids=ids, cu_seqlens=cu, max_seqlen=sl)

            def fwd_bwd():
                out = m(input_ids=ids, cu_seqlens=cu, max_seqlen=sl)
                out.hidden.float().sum().backward()
                for p in m.parameters():
                    p.grad = None

            t_f = _time(fwd, device)
            t_fb = _time(fwd_bwd, device)
            # embedding + head measured alone by running zero blocks
            m0 = ReaLModel(dataclasses.replace(cfg, n_layers=1), dtype=dtype, device=device, layer_range=(0, 1))
            m0.init_random_fast() if device.type == "cuda" else m0.instantiate()
            t_emb = _time(lambda: m0(input_ids=ids, cu_seqlens=cu, max_seqlen=sl), device)
            rows += [dict(bs=bs, seqlen=sl, layer="block", op="fwd", time_us=(t_f - t_emb) / n_blocks),
                     dict(bs=bs, seqlen=sl, layer="block", op="fwd_bwd", time_us=(t_fb - t_emb) / n_blocks),
                     dict(bs=bs, seqlen=sl, layer="embedding", op="fwd", time_us=t_emb)]
    return rows


def profile_head(cfg: ReaLModelConfig, batch_sizes: List[int], seqlens: List[int], device="cuda", dtype=torch.bfloat16, max_tokens: int = 8192) -> List[Dict]:
    """LM head rows: `fwd` / `fwd_bwd` of the fused LM-head + log-prob op over packed tokens (timed on at most `max_tokens` tokens,
    scaled linearly: the op is chunked over tokens anyway) and `decode` = full-vocabulary logits of one token per sequence."""
    from realhf_b200.ops import functional as OF
    device = torch.device(device)
    H, V = cfg.hidden_dim, cfg.vocab_size
    w = (torch.randn(V, H, device=device, dtype=torch.float32) * 0.02).to(dtype).requires_grad_(True)
    rows = []
    for bs in batch_sizes:
        for sl in seqlens:
            T = bs * sl
            t = min(T, max_tokens)
            x = (torch.randn(t, H, device=device, dtype=torch.float32)).to(dtype).requires_grad_(True)
            labels = torch.randint(0, V, (t,), device=device)

            def fwd():
                with torch.no_grad():
                    return OF.lm_head_logprobs(x, w, labels)

            def fwd_bwd():
                OF.lm_head_logprobs(x, w, labels).float().sum().backward()
                x.grad = None
                w.grad = None

            xd = x.detach()[:bs] if t >= bs else torch.randn(bs, H, device=device).to(dtype)

            def decode():
                with torch.no_grad():
                    return OF.linear(xd, w)

            rows += [dict(bs=bs, seqlen=sl, layer="head", op="fwd", time_us=_time(fwd, device) * T / t),
                     dict(bs=bs, seqlen=sl, layer="head", op="fwd_bwd", time_us=_time(fwd_bwd, device) * T / t),
                     dict(bs=bs, seqlen=sl, layer="head", op="decode", time_us=_time(decode, device))]
    return rows


def profile_decode(cfg: ReaLModelConfig, batch_sizes: List[int], ctx_lens: List[int], device="cuda", dtype=torch.bfloat16,
                   n_blocks: int = 2, n_tokens: int = 8) -> List[Dict]:
    """`decode` rows of one block: per-token time of `generate` (CUDA-graph decode on GPUs) for models of `n_blocks` and
    `2 * n_blocks` blocks; the difference divided by `n_blocks` is one block's share (embedding, head and sampling cancel)."""
    from realhf_b200.api.model import GenerationHyperparameters
    from realhf_b200.models import generation
    device = torch.device(device)

    def per_token(L: int, bs: int, ctx: int) -> float:
        m = ReaLModel(dataclasses.replace(cfg, n_layers=L), dtype=dtype, device=device)
        m.init_random_fast() if device.type == "cuda" else m.instantiate()
        ids = torch.randint(3, cfg.vocab_size, (bs * ctx,), device=device)
        cu = torch.arange(0, bs * ctx + 1, ctx, dtype=torch.int32, device=device)

        def run(n):
            g = GenerationHyperparameters(max_new_tokens=n, min_new_tokens=n, greedy=True, use_cuda_graph=device.type == "cuda")
            return lambda: generation.generate(m, ids, cu, g, eos_id=None, pad_id=0)

        t_short = _time(run(2), device, warmup=1, iters=2)
        t_long = _time(run(2 + n_tokens), device, warmup=1, iters=2)
        return max(t_long - t_short, 0.0) / n_tokens

    rows = []
    for bs in batch_sizes:
        for ctx in ctx_lens:
            a, b = per_token(n_blocks, bs, ctx), per_token(2 * n_blocks, bs, ctx)
            rows.append(dict(bs=bs, seqlen=ctx, layer="block", op="decode", time_us=max(b - a, 0.0) / n_blocks))
    return rows


def profile_optimizer(cfg: ReaLModelConfig, n_blocks: List[int] = (1, 2), device="cuda", dtype=torch.bfloat16, state_dtype: str = "fp32") -> List[Dict]:
    """`optimizer / step` rows: one step of the flat AdamW (grad-norm statistics + clip + fused update) over models of the given
    depths; `bs` holds the number of parameters in millions, which is what the cost model interpolates over."""
    from realhf_b200.engine.optim import FlatAdamW, OptimizerConfig
    device = torch.device(device)
    rows = []
    for L in n_blocks:
        m = ReaLModel(dataclasses.replace(cfg, n_layers=L), dtype=dtype, device=device)
        m.init_random_fast() if device.type == "cuda" else m.instantiate()
        on_gpu = device.type == "cuda"
        opt = FlatAdamW(m, OptimizerConfig(state_dtype=state_dtype if on_gpu else "fp32", use_master_weights=on_gpu and dtype != torch.float32,
                                           grad_dtype="bf16" if on_gpu else "fp32", lr_scheduler_type="constant", warmup_steps_proportion=0.0))
        opt.zero_grad()
        rows.append(dict(bs=m.flat_numel / 1e6, seqlen=0, layer="optimizer", op="step", time_us=_time(lambda: opt.step(), device)))
        del opt, m
    return rows


def dump_profile(rows: List[Dict], model_name: str) -> str:
    os.makedirs(constants.PROFILER_CACHE_PATH, exist_ok=True)
    p = os.path.join(constants.PROFILER_CACHE_PATH, f"layers_{model_name}.json")
    with open(p, "w") as f:
        json.dump(rows, f, indent=1)
    return p
