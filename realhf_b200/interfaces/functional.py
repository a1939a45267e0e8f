"""Losses and packed-batch helpers shared by the algorithm interfaces.

Parity: `realhf/impl/model/utils/ppo_functional.py` (KL controllers :12-46, actor loss :49-123, critic loss
:135-202, packed rewards / GAE :291-309,566-586), `dpo_functional.py`, `functional.py` (shifted log-prob gather,
masked normalisation :226-293) and `modules/rms.py` (running mean/std value normalisers).
"""

from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

from realhf_b200.ops import functional as OF

# ------------------------------------------------------------------------------------------- packed index helpers


def shifted_rows_and_labels(seqlens: Sequence[int], ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """For next-token prediction over packed sequences: row indices of every token but the last of each sequence,
    and the ids of their successors.  Built from host-side lengths (no device sync)."""
    lens = np.asarray(seqlens, dtype=np.int64)
    ends = np.cumsum(lens) - 1
    keep = np.ones(int(lens.sum()), dtype=bool)
    keep[ends] = False
    rows_np = np.nonzero(keep)[0]
    rows = torch.from_numpy(rows_np).to(ids.device, non_blocking=True)
    return rows, ids.index_select(0, rows + 1)


def short1_cu_seqlens(seqlens: Sequence[int], device) -> torch.Tensor:
    lens = torch.tensor([l - 1 for l in seqlens], dtype=torch.int32)
    cu = torch.zeros(len(seqlens) + 1, dtype=torch.int32)
    cu[1:] = lens.cumsum(0)
    return cu.to(device, non_blocking=True)


def seq_end_indices(seqlens: Sequence[int], device) -> torch.Tensor:
    return (torch.tensor(seqlens, dtype=torch.int64).cumsum(0) - 1).to(device, non_blocking=True)


def masked_normalization(x: torch.Tensor, mask: Optional[torch.Tensor], group=None, eps: float = 1e-5,
                         unbiased: bool = False) -> torch.Tensor:
    """(x - mean) / std over masked entries, statistics all-reduced over `group` in ONE packed collective."""
    x64 = x.double()
    m = mask.double() if mask is not None else torch.ones_like(x64)
    stats = torch.stack([(x64 * m).sum(), (x64 * x64 * m).sum(), m.sum()])
    if group is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, group=group)
    s1, s2, n = stats[0], stats[1], stats[2].clamp(min=1)
    mean = s1 / n
    var = (s2 / n - mean * mean).clamp(min=0)
    if unbiased:
        var = var * n / (n - 1).clamp(min=1)
    out = ((x64 - mean) / (var.sqrt() + eps)).float()
    return out * mask if mask is not None else out


def dp_reduce_stats(stats: Dict[str, torch.Tensor], group, device) -> Dict[str, float]:
    """Sum every scalar in one packed all-reduce (the reference issues one collective per scalar)."""
    keys = sorted(stats)
    vec = torch.stack([torch.as_tensor(stats[k], dtype=torch.float64, device=device).reshape(()) for k in keys])
    if group is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, group=group)
    vals = vec.tolist()
    return dict(zip(keys, vals))


# ------------------------------------------------------------------------------------------- KL controllers / normalisers


class FixedKLController:
    def __init__(self, kl_coef: float):
        self.value = kl_coef

    def update(self, current_kl: float, n_steps: int):
        pass

    def state_dict(self):
        return {"value": self.value}

    def load_state_dict(self, sd):
        self.value = sd["value"]


class AdaptiveKLController(FixedKLController):
    """https://arxiv.org/abs/1909.08593 (section 2.2)."""

    def __init__(self, init_kl_coef: float, target: float, horizon: float):
        super().__init__(init_kl_coef)
        self.target, self.horizon = target, horizon

    def update(self, current_kl: float, n_steps: int):
        err = float(np.clip(current_kl / self.target - 1, -0.2, 0.2))
        self.value *= 1 + err * n_steps / self.horizon


class ExponentialRunningMeanStd:
    def __init__(self, beta: float = 0.999, epsilon: float = 1e-5, high_precision: bool = True):
        self.beta, self.eps = beta, epsilon
        self.dtype = torch.float64 if high_precision else torch.float32
        self.mean = self.mean_sq = self.debias = None

    def _init(self, device):
        if self.mean is None:
            self.mean = torch.zeros((), dtype=self.dtype, device=device)
            self.mean_sq = torch.zeros((), dtype=self.dtype, device=device)
            self.debias = torch.zeros((), dtype=self.dtype, device=device)
        elif self.mean.device != torch.device(device):  # restored from a checkpoint (host tensors)
            for k in ("mean", "mean_sq", "debias", "count"):
                if getattr(self, k, None) is not None:
                    setattr(self, k, getattr(self, k).to(device))

    @torch.no_grad()
    def update(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None, group=None):
        self._init(x.device)
        x = x.to(self.dtype)
        m = mask.to(self.dtype) if mask is not None else torch.ones_like(x)
        st = torch.stack([(x * m).sum(), (x * x * m).sum(), m.sum()])
        if group is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(st, group=group)
        n = st[2].clamp(min=1)
        self.mean.mul_(self.beta).add_(st[0] / n * (1 - self.beta))
        self.mean_sq.mul_(self.beta).add_(st[1] / n * (1 - self.beta))
        self.debias.mul_(self.beta).add_(1 - self.beta)

    def mean_std(self):
        d = self.debias.clamp(min=self.eps)
        mean = self.mean / d
        var = (self.mean_sq / d - mean ** 2).clamp(min=1e-2)
        return mean, var.sqrt()

    def normalize(self, x):
        self._init(x.device)
        mean, std = self.mean_std()
        return ((x.to(self.dtype) - mean) / std).float()

    def denormalize(self, x):
        self._init(x.device)
        mean, std = self.mean_std()
        return (x.to(self.dtype) * std + mean).float()

    def state_dict(self):
        return {k: getattr(self, k).cpu() for k in ("mean", "mean_sq", "debias", "count") if getattr(self, k, None) is not None}

    def load_state_dict(self, sd):
        for k, v in sd.items():
            setattr(self, k, None if v is None else v.to(self.dtype))


class MovingAverageRunningMeanStd(ExponentialRunningMeanStd):
    """Plain cumulative average (beta -> 1 limit)."""

    def __init__(self, epsilon: float = 1e-5, high_precision: bool = True):
        super().__init__(1.0, epsilon, high_precision)
        self.count = None

    @torch.no_grad()
    def update(self, x, mask=None, group=None):
        self._init(x.device)
        if self.count is None:
            self.count = torch.zeros((), dtype=self.dtype, device=x.device)
        x = x.to(self.dtype)
        m = mask.to(self.dtype) if mask is not None else torch.ones_like(x)
        st = torch.stack([(x * m).sum(), (x * x * m).sum(), m.sum()])
        if group is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(st, group=group)
        tot = self.count + st[2]
        self.mean = (self.mean * self.count + st[0]) / tot.clamp(min=1)
        self.mean_sq = (self.mean_sq * self.count + st[1]) / tot.clamp(min=1)
        self.count = tot
        self.debias = torch.ones_like(self.mean)


# ------------------------------------------------------------------------------------------- losses


def actor_loss_fn(logprobs, old_logprobs, advantages, eps_clip: float, loss_mask: Optional[torch.Tensor] = None):
    """Clipped-ratio PPO policy loss (mean over masked tokens) + stats (clip ratio, importance weight, approx KL)."""
    logprobs, old_logprobs, advantages = logprobs.float(), old_logprobs.float(), advantages.float()
    mask = loss_mask.bool() if loss_mask is not None else torch.ones_like(logprobs, dtype=torch.bool)
    n = mask.count_nonzero().clamp(min=1)
    diff = torch.where(mask, logprobs - old_logprobs, torch.zeros_like(logprobs))
    ratio = torch.where(mask, torch.exp(diff), torch.zeros_like(diff))
    l1 = -advantages * ratio
    l2 = -advantages * ratio.clamp(1.0 - eps_clip, 1.0 + eps_clip)
    loss = torch.where(mask, torch.max(l1, l2), torch.zeros_like(l1)).sum() / n
    with torch.no_grad():
        stat = dict(clip_ratio=((l1 < l2) & mask).count_nonzero() / n, importance_weight=ratio.sum() / n,
                    approx_kl=diff.sum() / n)
    return loss, stat


def critic_loss_fn(value, old_value, target_value, value_eps_clip: float, loss_mask: Optional[torch.Tensor] = None,
                   loss_fn_type: str = "mse"):
    value, old_value, target_value = value.float(), old_value.float(), target_value.float()
    if loss_fn_type == "huber":
        f = lambda x, y: F.huber_loss(x, y, reduction="none", delta=10.0)
    elif loss_fn_type == "mse":
        f = lambda x, y: 0.5 * (x - y) ** 2
    else:
        raise NotImplementedError(loss_fn_type)
    l_orig = f(value, target_value)
    clipped = old_value + (value - old_value).clamp(-value_eps_clip, value_eps_clip)
    l_clip = f(clipped, target_value)
    l = torch.max(l_orig, l_clip)
    mask = loss_mask.bool() if loss_mask is not None else torch.ones_like(l, dtype=torch.bool)
    n = mask.count_nonzero().clamp(min=1)
    with torch.no_grad():
        stat = dict(clip_ratio=((l_clip > l_orig) & mask).count_nonzero() / n)
    return torch.where(mask, l, torch.zeros_like(l)).sum() / n, stat


def dpo_loss(pi_logps, ref_logps, beta: float):
    """Inputs are per-sequence answer log-prob sums ordered [pos0, neg0, pos1, neg1, ...]."""
    pi_pos, pi_neg = pi_logps[0::2], pi_logps[1::2]
    ref_pos, ref_neg = ref_logps[0::2], ref_logps[1::2]
    logits = (pi_pos - pi_neg) - (ref_pos - ref_neg)
    loss = -F.logsigmoid(beta * logits).mean()
    with torch.no_grad():
        pos_score = beta * (pi_pos - ref_pos)
        neg_score = beta * (pi_neg - ref_neg)
        kl = -(pi_pos - ref_pos).sum() - (pi_neg - ref_neg).sum()
    return loss, pos_score.detach().sum(), neg_score.detach().sum(), kl.detach()


def packed_rewards_and_gae(old_logp, ref_logp, scores, values, seqlens: Sequence[int], no_eos, kl_ctl: float,
                           clip_reward: float, gamma: float, lam: float):
    """KL-shaped reward + terminal score + GAE in one fused kernel.  Lengths follow the PPO convention:
    old/ref log-probs [sum (L-1)], values [sum L].  Returns (advantages, returns, kl_rewards, rewards)."""
    cu = short1_cu_seqlens(seqlens, old_logp.device)
    return OF.ppo_rewards_gae(old_logp, ref_logp, scores, values, cu, no_eos, gamma, lam, kl_ctl, clip_reward)


def pair_generation_outputs(x, outs):
    """[(micro-batch of `x`, GenerationOutput)]: the engine may use more micro-batches than requested (a pipeline needs at least
    pp of them), so the contiguous partition of `x` is rebuilt from the number of sequences in every output."""
    from realhf_b200.api.data import SequenceSample
    items, off, pairs = x.unpack(), 0, []
    for o in outs:
        nb = int(o.tokens.shape[0])
        pairs.append((SequenceSample.gather(items[off: off + nb]), o))
        off += nb
    assert off == x.bs, f"generation returned {off} sequences for {x.bs} prompts"
    return pairs
