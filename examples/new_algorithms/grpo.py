"""GRPO (group-relative policy optimisation) built only from public extension points.

Parity: reference `examples/new_algorithms/grpo/` — gen -> {rew_inf, ref_inf} -> actor_train, no critic.  Every prompt
is sampled `group_size` times; one data item holds the whole group (the inner list of `SequenceSample.seqlens`), and
the advantage of a response is its reward normalised inside its group.

    python examples/new_algorithms/grpo.py grpo experiment_name=grpo-demo trial_name=t0 actor.path=... rew.path=... dataset.path=...
"""

from __future__ import annotations

import dataclasses
import functools
from typing import Dict, Optional

import torch

from realhf_b200.api.config import DatasetAbstraction, ModelInterfaceAbstraction, ModelInterfaceType
from realhf_b200.api.data import SequenceSample
from realhf_b200.api.dfg import MFCDef
from realhf_b200.api.model import GenerationHyperparameters, Model, ModelInterface, register_interface
from realhf_b200.api.quickstart import MFCConfig, ModelTrainEvalConfig, PromptOnlyDatasetConfig, register_quickstart_exp
from realhf_b200.experiments.common import CommonExperimentConfig
from realhf_b200.interfaces import functional as IF
from realhf_b200.interfaces.ppo import _dp_group, _mb_prompt, _save_hf
from realhf_b200.models import generation as gen
from realhf_b200.models.real_model import ModelOutput


def _grpo_loss(out: ModelOutput, mb: SequenceSample, *, eps_clip: float, kl_coef: float, temperature: float):
    seqlens = mb.flat_seqlens("packed_input_ids")
    rows, labels = IF.shifted_rows_and_labels(seqlens, mb.data["packed_input_ids"])
    logp = out.logprobs(labels, None, temperature, rows)
    mask = mb.data["ppo_loss_mask"].bool()
    loss, st = IF.actor_loss_fn(logp, mb.data["old_logp"], mb.data["advantages"], eps_clip, mask)
    # k3 KL estimator against the reference policy
    d = (mb.data["ref_logp"] - logp) * mask
    kl = (torch.exp(d) - d - 1) * mask
    n = mask.count_nonzero().clamp(min=1)
    loss = loss + kl_coef * kl.sum() / n
    return loss, dict(actor_loss=loss.detach(), importance_weight=st["importance_weight"], kl=kl.sum().detach() / n)


@dataclasses.dataclass
class GRPOInterface(ModelInterface):
    group_size: int = 4
    n_minibatches: int = 2
    eps_clip: float = 0.2
    kl_coef: float = 0.04
    generation_config: Dict = dataclasses.field(default_factory=dict)
    enable_save: bool = True

    def __post_init__(self):
        self.gconfig = GenerationHyperparameters(**self.generation_config)

    @torch.no_grad()
    def generate(self, model: Model, input_: SequenceSample, n_mbs=None) -> Optional[SequenceSample]:
        eng = model.module
        plens = input_.flat_seqlens("packed_prompts")
        dev = input_.data["packed_prompts"].device
        # repeat every prompt group_size times
        cu = torch.tensor([0] + plens).cumsum(0).tolist()
        rep_ids, rep_lens = [], []
        for i, l in enumerate(plens):
            for g in range(self.group_size):
                rep_ids.append(input_.data["packed_prompts"][cu[i]:cu[i + 1]])
                rep_lens.append(l)
        x = SequenceSample.from_default(ids=[f"{i}-{g}" for i in input_.ids for g in range(self.group_size)], seqlens=rep_lens,
                                        data=dict(packed_input_ids=torch.cat(rep_ids)))
        outs = eng.generate(x, tokenizer=model.tokenizer, gconfig=self.gconfig, num_micro_batches=n_mbs)
        if outs is None:
            return None
        parts = []
        for mb, o in IF.pair_generation_outputs(x, outs):
            ids, c, _ = _mb_prompt(mb, dev)
            parts.append(gen.concat_prompt_to_generation_output(ids, c, o) + (o.no_eos,))
        packed = torch.cat([p[0] for p in parts])
        slens = [int(s) for s in torch.cat([p[1] for p in parts]).tolist()]
        G = self.group_size
        group = lambda lst: [lst[i * G:(i + 1) * G] for i in range(input_.bs)]
        keys = dict(packed_input_ids=(packed, slens), packed_logprobs=(torch.cat([p[2] for p in parts]), [l - 1 for l in slens]),
                    prompt_mask=(torch.cat([p[4] for p in parts]), slens), seq_no_eos_mask=(torch.cat([p[5] for p in parts]), [1] * len(slens)))
        with SequenceSample.disable_validation():
            return SequenceSample(keys=list(keys), ids=input_.ids, seqlens={k: group(v[1]) for k, v in keys.items()},
                                  trailing_shapes={k: () for k in keys}, dtypes={k: v[0].dtype for k, v in keys.items()},
                                  data={k: v[0] for k, v in keys.items()})

    @torch.no_grad()
    def inference(self, model: Model, input_: SequenceSample, n_mbs=None) -> Optional[SequenceSample]:
        eng = model.module

        def hook(out: ModelOutput, mb: SequenceSample):
            rows, labels = IF.shifted_rows_and_labels(mb.flat_seqlens("packed_input_ids"), mb.data["packed_input_ids"])
            return out.logprobs(labels, None, self.gconfig.temperature, rows)

        logp = eng.forward(input_, num_micro_batches=n_mbs, post_hook=hook)
        if logp is None:
            return None
        with SequenceSample.disable_validation():
            return SequenceSample(keys=["packed_ref_logprobs"], ids=input_.ids, trailing_shapes=dict(packed_ref_logprobs=()),
                                  dtypes=dict(packed_ref_logprobs=torch.float32), data=dict(packed_ref_logprobs=logp),
                                  seqlens=dict(packed_ref_logprobs=[[l - 1 for l in ls] for ls in input_.seqlens["packed_input_ids"]]))

    def train_step(self, model: Model, input_: SequenceSample, n_mbs=None) -> Dict:
        eng = model.module
        eng.eval()
        seqlens = input_.flat_seqlens("packed_input_ids")
        dev = input_.data["packed_input_ids"].device
        rewards = input_.data["rewards"].float().view(-1, self.group_size)
        adv_seq = ((rewards - rewards.mean(1, keepdim=True)) / (rewards.std(1, keepdim=True) + 1e-6)).view(-1)
        rows, _ = IF.shifted_rows_and_labels(seqlens, input_.data["packed_input_ids"])
        loss_mask = (~input_.data["prompt_mask"].bool()).index_select(0, rows + 1)
        adv = torch.repeat_interleave(adv_seq, torch.tensor([l - 1 for l in seqlens], device=dev)) * loss_mask
        flat_ids = [f"{i}-{g}" for i in input_.ids for g in range(self.group_size)]
        batch = SequenceSample.from_default(ids=flat_ids, seqlens=seqlens, data=dict(
            advantages=adv, old_logp=input_.data["packed_logprobs"].float() * loss_mask,
            ref_logp=input_.data["packed_ref_logprobs"].float() * loss_mask, ppo_loss_mask=loss_mask,
            packed_input_ids=input_.data["packed_input_ids"]))
        loss_fn = functools.partial(_grpo_loss, eps_clip=self.eps_clip, kl_coef=self.kl_coef, temperature=self.gconfig.temperature)
        stats: Dict[str, float] = {}
        from realhf_b200.interfaces.ppo import _dp_group, _n_minibatches
        mbs = batch.split(_n_minibatches(self.n_minibatches, batch.bs, _dp_group(model)))   # same number of steps on every DP rank
        for mb in mbs:
            st = eng.train_batch(mb, loss_fn, version_steps=model.version.global_step, num_micro_batches=n_mbs)
            for k, v in st.items():
                stats[k] = stats.get(k, 0.0) + float(v) / len(mbs)
        model.inc_version()
        stats["task_reward"] = float(rewards.mean())
        return stats

    def save(self, model: Model, save_dir: str):
        if self.enable_save:
            _save_hf(model, save_dir)


register_interface("grpo", GRPOInterface)


@dataclasses.dataclass
class GRPOConfig(CommonExperimentConfig):
    actor: ModelTrainEvalConfig = dataclasses.field(default_factory=ModelTrainEvalConfig)
    ref: ModelTrainEvalConfig = dataclasses.field(default_factory=ModelTrainEvalConfig)
    rew: ModelTrainEvalConfig = dataclasses.field(default_factory=ModelTrainEvalConfig)
    actor_train: MFCConfig = dataclasses.field(default_factory=MFCConfig)
    actor_gen: MFCConfig = dataclasses.field(default_factory=MFCConfig)
    ref_inf: MFCConfig = dataclasses.field(default_factory=MFCConfig)
    rew_inf: MFCConfig = dataclasses.field(default_factory=MFCConfig)
    dataset: PromptOnlyDatasetConfig = dataclasses.field(default_factory=PromptOnlyDatasetConfig)
    gen: GenerationHyperparameters = dataclasses.field(default_factory=GenerationHyperparameters)
    group_size: int = 4
    n_minibatches: int = 2

    def __post_init__(self):
        self.rew.type = dataclasses.replace(self.rew.type, is_critic=True)

    @property
    def models(self):
        return {"actor": self.actor, "ref": self.ref, "reward": self.rew}

    @property
    def rpcs(self):
        T = ModelInterfaceType
        itf = ModelInterfaceAbstraction("grpo", args=dict(group_size=self.group_size, n_minibatches=self.n_minibatches,
                                                          generation_config=dataclasses.asdict(self.gen)))
        rw = ModelInterfaceAbstraction("paired_rw", args=dict(enable_save=False))
        n = self.dataset.train_bs_n_seqs
        return {
            "actor_gen": MFCDef("actor_gen", n, T.GENERATE, itf, "actor", input_keys=("packed_prompts",),
                                output_keys=("packed_input_ids", "packed_logprobs", "prompt_mask", "seq_no_eos_mask"), n_mbs=self.actor_gen.n_mbs),
            "rew_inf": MFCDef("rew_inf", n, T.INFERENCE, rw, "reward", input_keys=("packed_input_ids",), output_keys=("rewards",), n_mbs=self.rew_inf.n_mbs),
            "ref_inf": MFCDef("ref_inf", n, T.INFERENCE, itf, "ref", input_keys=("packed_input_ids",), output_keys=("packed_ref_logprobs",), n_mbs=self.ref_inf.n_mbs),
            "actor_train": MFCDef("actor_train", n, T.TRAIN_STEP, itf, "actor",
                                  input_keys=("packed_input_ids", "packed_logprobs", "packed_ref_logprobs", "rewards", "prompt_mask", "seq_no_eos_mask"),
                                  log_return_value=True, n_mbs=self.actor_train.n_mbs),
        }

    @property
    def allocations(self):
        return {"actor_gen": self.actor_gen, "actor_train": self.actor_train, "ref_inf": self.ref_inf, "rew_inf": self.rew_inf}

    @property
    def datasets(self):
        return [DatasetAbstraction("prompt", args=dict(dataset_path=self.dataset.path, max_length=self.dataset.max_prompt_len))]

    @property
    def tokenizer_name_or_path(self):
        return self.actor.path

    @property
    def max_prompt_len(self):
        return self.dataset.max_prompt_len


register_quickstart_exp("grpo", GRPOConfig)

if __name__ == "__main__":
    import os
    import sys

    os.environ["REAL_USER_CODE"] = os.path.abspath(__file__)  # workers re-import this file to see the registrations
    from realhf_b200.apps.quickstart import main
    main(sys.argv[1:])
