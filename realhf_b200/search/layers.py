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


def dump_profile(rows: List[Dict], model_name: str) -> str:
    os.makedirs(constants.PROFILER_CACHE_PATH, exist_ok=True)
    p = os.path.join(constants.PROFILER_CACHE_PATH, f"layers_{model_name}.json")
    with open(p, "w") as f:
        json.dump(rows, f, indent=1)
    return p
