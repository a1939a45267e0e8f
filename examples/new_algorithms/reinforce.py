"""ReMax / REINFORCE with a greedy-decoding baseline, using input/output key remaps to reuse interfaces.

Parity: reference `examples/new_algorithms/reinforce/` — sample_gen + greedy_gen, two reward inferences, one train MFC;
the second generation / reward MFCs reuse the same interface implementations through `output_key_remap` /
`input_key_remap` (dfg.py key remapping), baseline = reward of the greedy response.
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
from realhf_b200.interfaces.ppo import PPOActorInterface, _save_hf
from realhf_b200.models.real_model import ModelOutput


def _reinforce_loss(out: ModelOutput, mb: SequenceSample):
    seqlens = mb.flat_seqlens("packed_input_ids")
    rows, labels = IF.shifted_rows_and_labels(seqlens, mb.data["packed_input_ids"])
    logp = out.logprobs(labels, None, 1.0, rows)
    mask = mb.data["ppo_loss_mask"].float()
    loss = -(logp * mb.data["advantages"] * mask).sum() / mask.sum().clamp(min=1)
    return loss, dict(loss=loss.detach())


@dataclasses.dataclass
class ReinforceInterface(ModelInterface):
    generation_config: Dict = dataclasses.field(default_factory=dict)
    greedy: bool = False
    enable_save: bool = True

    def __post_init__(self):
        g = dict(self.generation_config)
        g["greedy"] = self.greedy or g.get("greedy", False)
        self._gen = PPOActorInterface(generation_config=g, enable_save=False)

    def generate(self, model: Model, input_: SequenceSample, n_mbs=None) -> Optional[SequenceSample]:
        res = self._gen.generate(model, input_, n_mbs)
        if res is not None:
            with SequenceSample.disable_validation():
                res = SequenceSample.gather([res], keys=["packed_input_ids", "prompt_mask"])
        return res

    def train_step(self, model: Model, input_: SequenceSample, n_mbs=None) -> Dict:
        eng = model.module
        eng.eval()
        seqlens = input_.flat_seqlens("packed_input_ids")
        dev = input_.data["packed_input_ids"].device
        adv_seq = input_.data["rewards"].float() - input_.data["greedy_rewards"].float()   # ReMax baseline
        rows, _ = IF.shifted_rows_and_labels(seqlens, input_.data["packed_input_ids"])
        mask = (~input_.data["prompt_mask"].bool()).index_select(0, rows + 1)
        adv = torch.repeat_interleave(adv_seq, torch.tensor([l - 1 for l in seqlens], device=dev))
        batch = SequenceSample.from_default(ids=input_.ids, seqlens=seqlens, data=dict(
            advantages=adv, ppo_loss_mask=mask, packed_input_ids=input_.data["packed_input_ids"]))
        st = eng.train_batch(batch, _reinforce_loss, version_steps=model.version.global_step, num_micro_batches=n_mbs)
        model.inc_version()
        return dict(loss=float(st["loss"]), reward=float(input_.data["rewards"].mean()), baseline=float(input_.data["greedy_rewards"].mean()))

    def save(self, model: Model, save_dir: str):
        if self.enable_save:
            _save_hf(model, save_dir)


register_interface("reinforce", ReinforceInterface)


@dataclasses.dataclass
class ReinforceConfig(CommonExperimentConfig):
    actor: ModelTrainEvalConfig = dataclasses.field(default_factory=ModelTrainEvalConfig)
    rew: ModelTrainEvalConfig = dataclasses.field(default_factory=ModelTrainEvalConfig)
    dataset: PromptOnlyDatasetConfig = dataclasses.field(default_factory=PromptOnlyDatasetConfig)
    gen: GenerationHyperparameters = dataclasses.field(default_factory=GenerationHyperparameters)
    allocation: MFCConfig = dataclasses.field(default_factory=MFCConfig)

    def __post_init__(self):
        self.rew.type = dataclasses.replace(self.rew.type, is_critic=True)

    @property
    def models(self):
        return {"actor": self.actor, "reward": self.rew}

    @property
    def rpcs(self):
        T = ModelInterfaceType
        g = dataclasses.asdict(self.gen)
        sample = ModelInterfaceAbstraction("reinforce", args=dict(generation_config=g))
        greedy = ModelInterfaceAbstraction("reinforce", args=dict(generation_config=g, greedy=True, enable_save=False))
        rw = ModelInterfaceAbstraction("paired_rw", args=dict(enable_save=False))
        n = self.dataset.train_bs_n_seqs
        return {
            "sample_gen": MFCDef("sample_gen", n, T.GENERATE, sample, "actor", input_keys=("packed_prompts",),
                                 output_keys=("packed_input_ids", "prompt_mask")),
            "greedy_gen": MFCDef("greedy_gen", n, T.GENERATE, greedy, "actor", input_keys=("packed_prompts",),
                                 output_keys=("greedy_packed_input_ids",),
                                 output_key_remap={"packed_input_ids": "greedy_packed_input_ids", "prompt_mask": "greedy_prompt_mask"}),
            "sample_rew": MFCDef("sample_rew", n, T.INFERENCE, rw, "reward", input_keys=("packed_input_ids",), output_keys=("rewards",)),
            "greedy_rew": MFCDef("greedy_rew", n, T.INFERENCE, rw, "reward", input_keys=("greedy_packed_input_ids",),
                                 input_key_remap={"greedy_packed_input_ids": "packed_input_ids"}, output_keys=("greedy_rewards",),
                                 output_key_remap={"rewards": "greedy_rewards"}),
            "actor_train": MFCDef("actor_train", n, T.TRAIN_STEP, sample, "actor",
                                  input_keys=("packed_input_ids", "rewards", "greedy_rewards", "prompt_mask"), log_return_value=True),
        }

    @property
    def allocations(self):
        return {k: self.allocation for k in self.rpcs}

    @property
    def datasets(self):
        return [DatasetAbstraction("prompt", args=dict(dataset_path=self.dataset.path, max_length=self.dataset.max_prompt_len))]

    @property
    def tokenizer_name_or_path(self):
        return self.actor.path

    @property
    def max_prompt_len(self):
        return self.dataset.max_prompt_len


register_quickstart_exp("reinforce", ReinforceConfig)

if __name__ == "__main__":
    import os
    import sys

    os.environ["REAL_USER_CODE"] = os.path.abspath(__file__)
    from realhf_b200.apps.quickstart import main
    main(sys.argv[1:])
