"""Composable logits warpers (parity: `realhf/impl/model/utils/logits_warper.py`): temperature, top-k, top-p, epsilon.

The generation loop uses the fused sampling kernel (`ops/csrc/sampling.cu`) on CUDA; these classes are the reference
semantics (and the CPU path), usable from custom interfaces: `chained_logits_wraper([...])(input_ids, logits)`."""

from __future__ import annotations

from typing import List, Optional

import torch


class LogitsWarper:
    def __call__(self, input_ids: Optional[torch.Tensor], logits: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class TemperatureLogitsWarper(LogitsWarper):
    def __init__(self, temperature: float):
        assert temperature > 0
        self.temperature = temperature

    def __call__(self, input_ids, logits):
        return logits / self.temperature


class TopKLogitsWarper(LogitsWarper):
    def __init__(self, top_k: int, filter_value: float = -float("inf"), min_tokens_to_keep: int = 1):
        self.top_k, self.filter_value = max(top_k, min_tokens_to_keep), filter_value

    def __call__(self, input_ids, logits):
        k = min(self.top_k, logits.shape[-1])
        kth = torch.topk(logits, k)[0][..., -1, None]
        return logits.masked_fill(logits < kth, self.filter_value)


class TopPLogitsWarper(LogitsWarper):
    def __init__(self, top_p: float, filter_value: float = -float("inf"), min_tokens_to_keep: int = 1):
        assert 0 < top_p <= 1.0
        self.top_p, self.filter_value, self.min_keep = top_p, filter_value, min_tokens_to_keep

    def __call__(self, input_ids, logits):
        sorted_logits, sorted_idx = torch.sort(logits, descending=False)
        cum = sorted_logits.softmax(-1).cumsum(-1)
        remove = cum <= (1 - self.top_p)
        remove[..., -self.min_keep:] = False
        return logits.masked_fill(remove.scatter(-1, sorted_idx, remove), self.filter_value)


class EpsilonLogitsWarper(LogitsWarper):
    def __init__(self, epsilon: float, filter_value: float = -float("inf"), min_tokens_to_keep: int = 1):
        self.epsilon, self.filter_value, self.min_keep = epsilon, filter_value, min_tokens_to_keep

    def __call__(self, input_ids, logits):
        probs = logits.softmax(-1)
        remove = probs < self.epsilon
        k = min(self.min_keep, logits.shape[-1])
        remove = remove & (logits < torch.topk(logits, k)[0][..., -1, None])
        return logits.masked_fill(remove, self.filter_value)


def chained_logits_wraper(warpers: List[LogitsWarper]):
    def fn(input_ids, logits):
        for w in warpers:
            logits = w(input_ids, logits)
        return logits
    return fn


def top_k_top_p_logits(logits: torch.Tensor, top_k: int = 0, top_p: float = 1.0, inplace: bool = False) -> torch.Tensor:
    ws: List[LogitsWarper] = []
    if top_k and top_k > 0:
        ws.append(TopKLogitsWarper(top_k))
    if top_p < 1.0:
        ws.append(TopPLogitsWarper(top_p))
    out = chained_logits_wraper(ws)(None, logits)
    if inplace:
        logits.copy_(out)
        return logits
    return out
