"""PPO actor / critic interfaces.

Parity: `realhf/impl/model/interface/ppo_interface.py` (PPOActorInterface :110, PPOCriticInterface :639).
Same data contract between MFCs (keys `packed_prompts -> packed_input_ids, packed_logprobs, prompt_mask,
seq_no_eos_mask, packed_logits_mask -> packed_ref_logprobs / rewards / values -> train`).  Differences by design:
  * log-probs come from the fused LM-head kernel (`ModelOutput.logprobs`): no [T, V] logits anywhere;
  * the logits mask travels bit-packed ([sum(L-1), V/8] uint8) and is aligned with the log-probs;
  * KL reward + terminal score + GAE are one kernel launch; all scalar statistics of a step are reduced in one
    packed all-reduce.
"""

from __future__ import annotations

import collections
import dataclasses
import functools
from typing import Dict, Optional

import torch

from realhf_b200.api.data import SequenceSample
from realhf_b200.api.model import GenerationHyperparameters, Model, ModelInterface, register_interface
from realhf_b200.interfaces import functional as IF
from realhf_b200.models import generation as gen
from realhf_b200.models.real_model import ModelOutput


def _engine_ctx(model: Model):
    return getattr(model.module, "ctx", None)


def _dp_group(model: Model):
    ctx = _engine_ctx(model)
    return None if ctx is None else ctx.dp_group


def _save_hf(model: Model, save_dir: str):
    from realhf_b200.models import hf_io
    eng = model.module
    m = getattr(eng, "module", eng)
    fam = getattr(model, "hf_family", None) or getattr(m, "hf_family", "llama")
    optim = getattr(eng, "optim", None)
    if optim is not None:
        optim.materialize()  # ZeRO-3 keeps only this rank's parameter shard between calls (collective over DP: all ranks save)
    try:
        hf_io.save_to_hf(m, fam, save_dir, tokenizer=model.tokenizer)
    finally:
        if optim is not None:
            optim.release()


def _actor_loss_from_output(out: ModelOutput, mb: SequenceSample, *, kl_adapter, eps_clip: float, temperature: float,
                            early_stop_imp_ratio: Optional[float], early_stop_kl: Optional[float]):
    seqlens = mb.flat_seqlens("packed_input_ids")
    rows, labels = IF.shifted_rows_and_labels(seqlens, mb.data["packed_input_ids"])
    mask_bits = mb.data.get("packed_logits_mask")
    logp = out.logprobs(labels, mask_bits, temperature, rows)
    loss_mask = mb.data["ppo_loss_mask"].bool()
    loss, stat = IF.actor_loss_fn(logp, mb.data["old_logp"], mb.data["advantages"], eps_clip, loss_mask)
    n = loss_mask.count_nonzero().clamp(min=1)
    stat = dict(ppo_approx_kl=stat["approx_kl"], actor_clip_ratio=stat["clip_ratio"],
                importance_weight=stat["importance_weight"], actor_loss=loss.detach())
    # early stopping: zero the loss on the device (no host sync) when the policy drifted too far
    if early_stop_imp_ratio is not None:
        loss = loss * (stat["importance_weight"] <= early_stop_imp_ratio).float()
    if early_stop_kl is not None:
        loss = loss * (stat["ppo_approx_kl"] <= early_stop_kl).float()
    return loss, stat


def _n_minibatches(wanted: int, local_bs: int, dp_group) -> int:
    """Every rank of a data-parallel group must take the same number of optimizer steps (each one contains the gradient
    collectives of the group).  A rank that received fewer sequences than `n_minibatches` cannot: fail with the reason instead of
    running fewer steps than its peers and leaving them inside a collective.  (The master gives every DP rank of a train MFC at
    least `n_minibatches` sequences whenever the batch has that many per rank.)"""
    if local_bs >= wanted:
        return wanted
    import torch.distributed as dist
    if dp_group is not None and dist.is_initialized() and dist.get_world_size(dp_group) > 1:
        raise ValueError(f"this data-parallel rank received {local_bs} sequences but ppo_n_minibatches={wanted}: every rank needs at least one "
                         f"sequence per minibatch -- raise dataset.train_bs_n_seqs, lower ppo.ppo_n_minibatches or use a smaller "
                         f"data-parallel degree for the training MFCs")
    return max(1, local_bs)


@dataclasses.dataclass
class PPOActorInterface(ModelInterface):
    n_minibatches: int = 4
    generation_config: Dict = dataclasses.field(default_factory=dict)
    kl_ctl: float = 0.1
    adv_norm: bool = True
    discount: float = 1.0
    gae_lambda: float = 1.0
    eps_clip: float = 0.2
    value_eps_clip: float = 0.2
    max_reward_clip: float = 5.0
    early_stop_kl: Optional[float] = None
    early_stop_imp_ratio: Optional[float] = None
    adaptive_kl_ctl: bool = False
    adaptive_kl_target: Optional[float] = 6
    adaptive_kl_horizon: Optional[float] = 10000
    enable_save: bool = True
    value_norm: bool = False
    value_norm_type: str = "exp"
    value_norm_beta: float = 0.99995
    value_norm_eps: float = 1e-5

    def __post_init__(self):
        self.kl_adapter = (IF.AdaptiveKLController(self.kl_ctl, self.adaptive_kl_target, self.adaptive_kl_horizon)
                           if self.adaptive_kl_ctl else IF.FixedKLController(self.kl_ctl))
        self.rms = None
        if self.value_norm:
            self.rms = (IF.ExponentialRunningMeanStd(self.value_norm_beta, self.value_norm_eps)
                        if self.value_norm_type == "exp" else IF.MovingAverageRunningMeanStd(self.value_norm_eps))
        g = self.generation_config
        self.gconfig = g if isinstance(g, GenerationHyperparameters) else GenerationHyperparameters(**g)

    def save(self, model: Model, save_dir: str):
        if self.enable_save:
            _save_hf(model, save_dir)

    def state_dict(self):
        return {"kl": self.kl_adapter.state_dict(), "rms": None if self.rms is None else self.rms.state_dict()}

    def load_state_dict(self, sd):
        self.kl_adapter.load_state_dict(sd["kl"])
        if self.rms is not None and sd.get("rms") is not None:
            self.rms.load_state_dict(sd["rms"])

    # ------------------------------------------------------------------ generate
    @torch.no_grad()
    def generate(self, model: Model, input_: SequenceSample, n_mbs=None) -> Optional[SequenceSample]:
        eng = model.module
        eng.eval()
        plens = input_.flat_seqlens("packed_prompts")
        x = SequenceSample.from_default(ids=input_.ids, seqlens=plens,
                                        data=dict(packed_input_ids=input_.data["packed_prompts"]))
        outs = eng.generate(x, tokenizer=model.tokenizer, gconfig=self.gconfig, num_micro_batches=n_mbs)
        if outs is None:
            return None
        dev = input_.data["packed_prompts"].device
        parts = []
        for mb, o in IF.pair_generation_outputs(x, outs):
            ids, cu, _ = _mb_prompt(mb, dev)
            packed, slens, lp, mask_bits, in_prompt = gen.concat_prompt_to_generation_output(ids, cu, o)
            parts.append((packed, slens, lp, mask_bits, in_prompt, o.no_eos))
        packed = torch.cat([p[0] for p in parts])
        slens = torch.cat([p[1] for p in parts]).tolist()  # the one host read of the call: sequence lengths are metadata
        data = dict(packed_input_ids=packed, packed_logprobs=torch.cat([p[2] for p in parts]),
                    prompt_mask=torch.cat([p[4] for p in parts]), seq_no_eos_mask=torch.cat([p[5] for p in parts]))
        if parts[0][3] is not None:
            data["packed_logits_mask"] = torch.cat([p[3] for p in parts])
        return SequenceSample.from_default(ids=input_.ids, seqlens=[int(s) for s in slens], data=data)

    # ------------------------------------------------------------------ inference (reference log-probs)
    @torch.no_grad()
    def inference(self, model: Model, input_: SequenceSample, n_mbs=None) -> Optional[SequenceSample]:
        eng = model.module
        eng.eval()
        temperature = self.gconfig.temperature

        def calc_logprobs(out: ModelOutput, mb: SequenceSample):
            rows, labels = IF.shifted_rows_and_labels(mb.flat_seqlens("packed_input_ids"), mb.data["packed_input_ids"])
            return out.logprobs(labels, mb.data.get("packed_logits_mask"), temperature, rows)

        logp = eng.forward(input_, num_micro_batches=n_mbs, post_hook=calc_logprobs)
        if logp is None:
            return None
        return SequenceSample.from_default(ids=input_.ids, seqlens=input_.flat_seqlens("packed_input_ids"),
                                           data=dict(packed_ref_logprobs=logp))

    # ------------------------------------------------------------------ train
    def train_step(self, model: Model, input_: SequenceSample, n_mbs=None) -> Dict:
        eng = model.module
        eng.eval()  # dropout would make the recomputed log-probs inconsistent with generation
        dev = input_.data["packed_input_ids"].device
        seqlens = input_.flat_seqlens("packed_input_ids")
        old_logp = input_.data["packed_logprobs"].float()
        ref_logp = input_.data["packed_ref_logprobs"].float()
        prompt_mask = input_.data["prompt_mask"].bool()
        scores = input_.data["rewards"].float()
        values = input_.data["values"].float()
        no_eos = input_.data["seq_no_eos_mask"].bool()
        group = _dp_group(model)

        if self.rms is not None:
            values = self.rms.denormalize(values)
        ends = IF.seq_end_indices(seqlens, dev)
        values = values.clone()
        values[ends] = torch.where(no_eos, values[ends], torch.zeros_like(values[ends]))  # V(EOS) = 0
        rows, _ = IF.shifted_rows_and_labels(seqlens, input_.data["packed_input_ids"])
        loss_mask = (~prompt_mask).index_select(0, rows + 1)
        old_logp, ref_logp = old_logp * loss_mask, ref_logp * loss_mask
        adv, ret, kl_rewards, _ = IF.packed_rewards_and_gae(old_logp, ref_logp, scores, values, seqlens, no_eos,
                                                            self.kl_adapter.value, self.max_reward_clip, self.discount,
                                                            self.gae_lambda)
        if self.rms is not None:
            self.rms.update(ret, mask=loss_mask, group=group)
        if self.adv_norm:
            adv = IF.masked_normalization(adv, loss_mask, group)
        data = dict(advantages=adv, old_logp=old_logp, ppo_loss_mask=loss_mask,
                    packed_input_ids=input_.data["packed_input_ids"], kl_rewards=kl_rewards)
        if input_.data.get("packed_logits_mask") is not None:
            data["packed_logits_mask"] = input_.data["packed_logits_mask"]
        batch = SequenceSample.from_default(ids=input_.ids, seqlens=seqlens, data=data)
        ctx = _engine_ctx(model)
        pp = ctx.pp_size if ctx is not None else 1
        n_mbs = n_mbs or 1
        minibatches = batch.split(_n_minibatches(self.n_minibatches, batch.bs, group), min_size=(pp * 2 * n_mbs if pp > 1 else n_mbs))

        sums = dict(n_seqs=float(len(seqlens)), task_reward=scores.sum(), n_tokens=loss_mask.count_nonzero(),
                    kl_reward=(kl_rewards * loss_mask).sum(), advantage=adv.sum(),
                    prompt_len=prompt_mask.count_nonzero(), seq_len=float(sum(seqlens)))
        loss_fn = functools.partial(_actor_loss_from_output, kl_adapter=self.kl_adapter, eps_clip=self.eps_clip,
                                    temperature=self.gconfig.temperature, early_stop_imp_ratio=self.early_stop_imp_ratio,
                                    early_stop_kl=self.early_stop_kl)
        train_stats: Dict[str, torch.Tensor] = collections.defaultdict(float)
        for mb in minibatches:
            st = eng.train_batch(mb, loss_fn, version_steps=model.version.global_step, num_micro_batches=n_mbs)
            for k, v in st.items():
                train_stats[k] = train_stats[k] + v
        model.inc_version()

        g = IF.dp_reduce_stats(sums, group, dev)
        n_tok, n_seq = max(g["n_tokens"], 1.0), max(g["n_seqs"], 1.0)
        out = dict(task_reward=g["task_reward"] / n_seq, kl_reward=g["kl_reward"] / n_tok, advantage=g["advantage"] / n_tok,
                   avg_seq_len=g["seq_len"] / n_seq, avg_prompt_len=g["prompt_len"] / n_seq, n_tokens=int(g["n_tokens"]),
                   n_seqs=int(g["n_seqs"]), kl_ctl=self.kl_adapter.value)
        ts = IF.dp_reduce_stats({k: torch.as_tensor(v, device=dev) for k, v in train_stats.items()}, group, dev)
        dp = ctx.dp_size if ctx is not None else 1
        for k, v in ts.items():
            out[k] = v / (dp * len(minibatches))
        # KL controller: mean KL between policy and reference over this batch (kl_reward = -kl_ctl * kl)
        if self.kl_adapter.value != 0:
            mean_ref_kl = -out["kl_reward"] / self.kl_adapter.value
            self.kl_adapter.update(mean_ref_kl, n_steps=int(g["n_seqs"]))
        return out

    # profiler hooks
    def _mock_inference(self, model: Model, data: SequenceSample):
        return data

    def _mock_train_step(self, model: Model, data: SequenceSample):
        return mock_ppo_train_inputs(data, model, with_values=True)


def _mb_prompt(mb: SequenceSample, device):
    lens = mb.flat_seqlens("packed_input_ids")
    cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
    cu[1:] = torch.tensor(lens, dtype=torch.int32).cumsum(0)
    return mb.data["packed_input_ids"], cu.to(device), max(lens)


def _critic_loss_from_output(out: ModelOutput, mb: SequenceSample, *, value_eps_clip: float, loss_fn_type: str, rms):
    seqlens = mb.flat_seqlens("packed_input_ids")
    rows, _ = IF.shifted_rows_and_labels(seqlens, mb.data["packed_input_ids"])
    new_values = out.values.index_select(0, rows)  # the last token of a sequence is not a state
    loss, stat = IF.critic_loss_fn(new_values, mb.data["values"], mb.data["returns"], value_eps_clip,
                                   mb.data["ppo_loss_mask"].bool(), loss_fn_type)
    return loss, dict(value_loss=loss.detach(), value_clip_ratio=stat["clip_ratio"])


@dataclasses.dataclass
class PPOCriticInterface(ModelInterface):
    n_minibatches: int = 4
    enable_save: bool = True
    kl_ctl: float = 0.1
    discount: float = 1.0
    gae_lambda: float = 0.95
    value_eps_clip: float = 0.2
    max_reward_clip: float = 5.0
    adaptive_kl_ctl: bool = False
    adaptive_kl_target: Optional[float] = 6
    adaptive_kl_horizon: Optional[float] = 10000
    value_loss_type: str = "mse"
    value_norm: bool = False
    value_norm_type: str = "exp"
    value_norm_beta: float = 0.99995
    value_norm_eps: float = 1e-5

    def __post_init__(self):
        self.kl_adapter = (IF.AdaptiveKLController(self.kl_ctl, self.adaptive_kl_target, self.adaptive_kl_horizon)
                           if self.adaptive_kl_ctl else IF.FixedKLController(self.kl_ctl))
        self.rms = None
        if self.value_norm:
            self.rms = (IF.ExponentialRunningMeanStd(self.value_norm_beta, self.value_norm_eps)
                        if self.value_norm_type == "exp" else IF.MovingAverageRunningMeanStd(self.value_norm_eps))

    def save(self, model: Model, save_dir: str):
        if self.enable_save:
            _save_hf(model, save_dir)

    def state_dict(self):
        return {"kl": self.kl_adapter.state_dict(), "rms": None if self.rms is None else self.rms.state_dict()}

    def load_state_dict(self, sd):
        self.kl_adapter.load_state_dict(sd["kl"])
        if self.rms is not None and sd.get("rms") is not None:
            self.rms.load_state_dict(sd["rms"])

    @torch.no_grad()
    def inference(self, model: Model, input_: SequenceSample, n_mbs=None) -> Optional[SequenceSample]:
        eng = model.module
        eng.eval()
        values = eng.forward(input_, num_micro_batches=n_mbs, post_hook=lambda out, mb: out.values)
        if values is None:
            return None
        return SequenceSample.from_default(ids=input_.ids, seqlens=input_.flat_seqlens("packed_input_ids"),
                                           data=dict(values=values.float()))

    def train_step(self, model: Model, input_: SequenceSample, n_mbs=None) -> Dict:
        eng = model.module
        eng.eval()
        dev = input_.data["packed_input_ids"].device
        seqlens = input_.flat_seqlens("packed_input_ids")
        old_logp = input_.data["packed_logprobs"].float()
        ref_logp = input_.data["packed_ref_logprobs"].float()
        prompt_mask = input_.data["prompt_mask"].bool()
        scores = input_.data["rewards"].float()
        values = input_.data["values"].float()
        no_eos = input_.data["seq_no_eos_mask"].bool()
        group = _dp_group(model)
        denorm = self.rms.denormalize(values) if self.rms is not None else values
        ends = IF.seq_end_indices(seqlens, dev)
        denorm, values = denorm.clone(), values.clone()
        keep = no_eos
        denorm[ends] = torch.where(keep, denorm[ends], torch.zeros_like(denorm[ends]))
        values[ends] = torch.where(keep, values[ends], torch.zeros_like(values[ends]))
        rows, _ = IF.shifted_rows_and_labels(seqlens, input_.data["packed_input_ids"])
        loss_mask = (~prompt_mask).index_select(0, rows + 1)
        old_logp, ref_logp = old_logp * loss_mask, ref_logp * loss_mask
        _, ret, kl_rewards, _ = IF.packed_rewards_and_gae(old_logp, ref_logp, scores, denorm, seqlens, no_eos,
                                                          self.kl_adapter.value, self.max_reward_clip, self.discount,
                                                          self.gae_lambda)
        if self.rms is not None:
            self.rms.update(ret, mask=loss_mask, group=group)
            norm_ret = self.rms.normalize(ret)
        else:
            norm_ret = ret
        with SequenceSample.disable_validation():  # `values` is already shifted to L-1 here (default rule says L)
            batch = SequenceSample.from_default(
                ids=input_.ids, seqlens=seqlens,
                data=dict(returns=norm_ret, values=values.index_select(0, rows), ppo_loss_mask=loss_mask,
                          packed_input_ids=input_.data["packed_input_ids"], kl_rewards=kl_rewards))
        batch.seqlens["values"] = [[l - 1] for l in seqlens]
        ctx = _engine_ctx(model)
        pp = ctx.pp_size if ctx is not None else 1
        n_mbs = n_mbs or 1
        minibatches = batch.split(_n_minibatches(self.n_minibatches, batch.bs, group), min_size=(pp * 2 * n_mbs if pp > 1 else n_mbs))
        loss_fn = functools.partial(_critic_loss_from_output, value_eps_clip=self.value_eps_clip,
                                    loss_fn_type=self.value_loss_type, rms=self.rms)
        train_stats: Dict[str, torch.Tensor] = collections.defaultdict(float)
        for mb in minibatches:
            st = eng.train_batch(mb, loss_fn, version_steps=model.version.global_step, num_micro_batches=n_mbs)
            for k, v in st.items():
                train_stats[k] = train_stats[k] + v
        model.inc_version()
        sums = dict(returns=(ret * loss_mask).sum(), n_tokens=loss_mask.count_nonzero(),
                    kl_reward=(kl_rewards * loss_mask).sum())
        g = IF.dp_reduce_stats(sums, group, dev)
        n_tok = max(g["n_tokens"], 1.0)
        out = dict(returns=g["returns"] / n_tok, n_tokens=int(g["n_tokens"]))
        ts = IF.dp_reduce_stats({k: torch.as_tensor(v, device=dev) for k, v in train_stats.items()}, group, dev)
        dp = ctx.dp_size if ctx is not None else 1
        for k, v in ts.items():
            out[k] = v / (dp * len(minibatches))
        if self.kl_adapter.value != 0:
            self.kl_adapter.update(-(g["kl_reward"] / n_tok) / self.kl_adapter.value, n_steps=len(seqlens))
        return out

    def _mock_train_step(self, model: Model, data: SequenceSample):
        return mock_ppo_train_inputs(data, model, with_values=True)


def mock_ppo_train_inputs(data: SequenceSample, model: Model, with_values: bool) -> SequenceSample:
    """Fabricate the inputs of a PPO train step from a batch of sequences (profiling / benchmarks)."""
    seqlens = data.flat_seqlens("packed_input_ids")
    dev = data.data["packed_input_ids"].device
    n, bs = sum(seqlens), len(seqlens)
    extra = dict(packed_logprobs=-torch.rand(n - bs, device=dev), packed_ref_logprobs=-torch.rand(n - bs, device=dev),
                 prompt_mask=torch.zeros(n, dtype=torch.bool, device=dev), rewards=torch.randn(bs, device=dev),
                 seq_no_eos_mask=torch.zeros(bs, dtype=torch.bool, device=dev))
    if with_values:
        extra["values"] = torch.randn(n, device=dev)
    data.update_(SequenceSample.from_default(ids=data.ids, seqlens=seqlens, data=extra))
    return data


register_interface("ppo_actor", PPOActorInterface)
register_interface("ppo_critic", PPOCriticInterface)
