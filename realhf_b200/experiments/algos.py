"""Built-in experiments: sft, rw, dpo, ppo, gen (+ profile).

Parity: `realhf/experiments/common/{sft,rw,dpo,ppo,gen}_exp.py`.  Each is a dataclass whose fields are the CLI
options (`actor.path=...`, `ppo.gen.max_new_tokens=512`, `actor_train.parallel.model_parallel_size=2` ...).
"""

from __future__ import annotations

import copy
import dataclasses
from typing import Any, Dict, Optional

from realhf_b200.api.config import DatasetAbstraction, ModelInterfaceAbstraction, ModelInterfaceType
from realhf_b200.api.dfg import MFCDef
from realhf_b200.api.model import GenerationHyperparameters
from realhf_b200.api.quickstart import (MFCConfig, ModelTrainEvalConfig, PairedComparisonDatasetConfig, PromptAnswerDatasetConfig,
                                        PromptOnlyDatasetConfig, register_quickstart_exp)
from realhf_b200.experiments.common import CommonExperimentConfig

T = ModelInterfaceType


@dataclasses.dataclass
class SFTConfig(CommonExperimentConfig):
    model: ModelTrainEvalConfig = dataclasses.field(default_factory=ModelTrainEvalConfig)
    allocation: MFCConfig = dataclasses.field(default_factory=MFCConfig)
    dataset: PromptAnswerDatasetConfig = dataclasses.field(default_factory=PromptAnswerDatasetConfig)

    @property
    def models(self):
        return {"default": self.model}

    @property
    def rpcs(self):
        return {"trainDefault": MFCDef(name="trainDefault", n_seqs=self.dataset.train_bs_n_seqs, interface_type=T.TRAIN_STEP,
                                       interface_impl=ModelInterfaceAbstraction("sft"), model_name="default",
                                       input_keys=("packed_input_ids", "prompt_mask"), log_return_value=True,
                                       model_type=self.model.type, model_path=self.model.path, n_mbs=self.allocation.n_mbs)}

    @property
    def allocations(self):
        return {"trainDefault": self.allocation}

    @property
    def datasets(self):
        return [DatasetAbstraction("prompt_answer", args=dict(max_length=self.dataset.max_seqlen, dataset_path=self.dataset.train_path,
                                                              pad_to_max_length=self.dataset.pad_to_max_length))]

    @property
    def eval_datasets(self):
        if not self.dataset.valid_path:
            return None
        return [DatasetAbstraction("prompt_answer", args=dict(max_length=self.dataset.max_seqlen, dataset_path=self.dataset.valid_path))]

    @property
    def tokenizer_name_or_path(self):
        return self.model.path


@dataclasses.dataclass
class RWConfig(CommonExperimentConfig):
    is_sft_lora: bool = False
    model: ModelTrainEvalConfig = dataclasses.field(default_factory=ModelTrainEvalConfig)
    allocation: MFCConfig = dataclasses.field(default_factory=MFCConfig)
    dataset: PairedComparisonDatasetConfig = dataclasses.field(default_factory=PairedComparisonDatasetConfig)

    def __post_init__(self):
        self.model.type = dataclasses.replace(self.model.type, is_critic=True)
        self.model.init_critic_from_actor = True if not self.model.init_from_scratch else self.model.init_critic_from_actor

    @property
    def models(self):
        return {"default": self.model}

    @property
    def rpcs(self):
        return {"trainDefault": MFCDef(name="trainDefault", n_seqs=self.dataset.train_bs_n_seqs, interface_type=T.TRAIN_STEP,
                                       interface_impl=ModelInterfaceAbstraction("paired_rw"), model_name="default",
                                       input_keys=("packed_input_ids",), log_return_value=True, model_type=self.model.type,
                                       model_path=self.model.path, n_mbs=self.allocation.n_mbs)}

    @property
    def allocations(self):
        return {"trainDefault": self.allocation}

    @property
    def datasets(self):
        return [DatasetAbstraction("rw_pair", args=dict(max_length=self.dataset.max_seqlen, dataset_path=self.dataset.train_path,
                                                        max_pairs_per_prompt=self.dataset.max_pairs_per_prompt))]

    @property
    def eval_datasets(self):
        if not self.dataset.valid_path:
            return None
        return [DatasetAbstraction("rw_pair", args=dict(max_length=self.dataset.max_seqlen, dataset_path=self.dataset.valid_path,
                                                        max_pairs_per_prompt=self.dataset.max_pairs_per_prompt))]

    @property
    def tokenizer_name_or_path(self):
        return self.model.path


@dataclasses.dataclass
class DPOConfig(CommonExperimentConfig):
    is_sft_lora: bool = False
    actor: ModelTrainEvalConfig = dataclasses.field(default_factory=ModelTrainEvalConfig)
    ref: ModelTrainEvalConfig = dataclasses.field(default_factory=ModelTrainEvalConfig)
    actor_train: MFCConfig = dataclasses.field(default_factory=MFCConfig)
    ref_inf: MFCConfig = dataclasses.field(default_factory=MFCConfig)
    dataset: PairedComparisonDatasetConfig = dataclasses.field(default_factory=PairedComparisonDatasetConfig)
    beta: float = 0.1

    @property
    def models(self):
        return {"actor": self.actor, "ref": self.ref}

    @property
    def rpcs(self):
        itf = ModelInterfaceAbstraction("dpo", args=dict(beta=self.beta, enable_save=True))
        ref_itf = ModelInterfaceAbstraction("dpo", args=dict(beta=self.beta, enable_save=False))
        n = self.dataset.train_bs_n_seqs
        return {
            "ref_inf": MFCDef(name="ref_inf", n_seqs=n, interface_type=T.INFERENCE, interface_impl=ref_itf, model_name="ref",
                              input_keys=("packed_input_ids", "prompt_mask"), output_keys=("seqlogp",), model_type=self.ref.type,
                              model_path=self.ref.path, n_mbs=self.ref_inf.n_mbs),
            "actor_train": MFCDef(name="actor_train", n_seqs=n, interface_type=T.TRAIN_STEP, interface_impl=itf, model_name="actor",
                                  input_keys=("packed_input_ids", "seqlogp", "prompt_mask"), log_return_value=True,
                                  model_type=self.actor.type, model_path=self.actor.path, n_mbs=self.actor_train.n_mbs),
        }

    @property
    def allocations(self):
        return {"ref_inf": self.ref_inf, "actor_train": self.actor_train}

    @property
    def datasets(self):
        return [DatasetAbstraction("rw_pair", args=dict(max_length=self.dataset.max_seqlen, dataset_path=self.dataset.train_path,
                                                        max_pairs_per_prompt=self.dataset.max_pairs_per_prompt))]

    @property
    def tokenizer_name_or_path(self):
        return self.actor.path


@dataclasses.dataclass
class PPOHyperparameters:
    gen: GenerationHyperparameters = dataclasses.field(default_factory=GenerationHyperparameters)
    ppo_n_minibatches: int = 4
    kl_ctl: float = 0.1
    discount: float = 1.0
    gae_lambda: float = 1.0
    eps_clip: float = 0.2
    value_eps_clip: float = 0.2
    max_reward_clip: float = 20.0
    reward_output_scaling: float = 1.0
    reward_output_bias: float = 0.0
    early_stop_imp_ratio: float = 5.0
    use_adaptive_kl_ctl: bool = False
    adv_norm: bool = True
    value_norm: bool = True
    value_norm_type: str = "exp"
    value_norm_beta: float = 0.99995
    value_norm_eps: float = 1e-5


@dataclasses.dataclass
class PPOConfig(CommonExperimentConfig):
    is_sft_lora: bool = False
    is_rew_lora: bool = False
    actor: ModelTrainEvalConfig = dataclasses.field(default_factory=ModelTrainEvalConfig)
    critic: ModelTrainEvalConfig = dataclasses.field(default_factory=ModelTrainEvalConfig)
    ref: ModelTrainEvalConfig = dataclasses.field(default_factory=ModelTrainEvalConfig)
    rew: ModelTrainEvalConfig = dataclasses.field(default_factory=ModelTrainEvalConfig)
    actor_train: MFCConfig = dataclasses.field(default_factory=MFCConfig)
    critic_train: MFCConfig = dataclasses.field(default_factory=MFCConfig)
    actor_gen: MFCConfig = dataclasses.field(default_factory=MFCConfig)
    critic_inf: MFCConfig = dataclasses.field(default_factory=MFCConfig)
    rew_inf: MFCConfig = dataclasses.field(default_factory=MFCConfig)
    ref_inf: MFCConfig = dataclasses.field(default_factory=MFCConfig)
    dataset: PromptOnlyDatasetConfig = dataclasses.field(default_factory=PromptOnlyDatasetConfig)
    ppo: PPOHyperparameters = dataclasses.field(default_factory=PPOHyperparameters)

    def __post_init__(self):
        self.critic.type = dataclasses.replace(self.critic.type, is_critic=True)
        self.rew.type = dataclasses.replace(self.rew.type, is_critic=True)
        if self.actor.lora or self.critic.lora or self.ref.lora or self.rew.lora:
            raise NotImplementedError("LoRA is not supported (as in the reference, ppo_exp.py:199-202)")

    @property
    def models(self):
        return {"actor": self.actor, "critic": self.critic, "ref": self.ref, "reward": self.rew}

    @property
    def ppo_kwargs(self) -> Dict[str, Any]:
        p = self.ppo
        return dict(n_minibatches=p.ppo_n_minibatches, kl_ctl=p.kl_ctl, discount=p.discount, gae_lambda=p.gae_lambda,
                    eps_clip=p.eps_clip, value_eps_clip=p.value_eps_clip, max_reward_clip=p.max_reward_clip,
                    adaptive_kl_ctl=p.use_adaptive_kl_ctl, value_norm=p.value_norm, value_norm_type=p.value_norm_type,
                    value_norm_beta=p.value_norm_beta, value_norm_eps=p.value_norm_eps)

    @property
    def rpcs(self):
        p = self.ppo
        kw = self.ppo_kwargs
        actor_itf = ModelInterfaceAbstraction("ppo_actor", args={**copy.deepcopy(kw), "generation_config": dataclasses.asdict(p.gen),
                                                                 "early_stop_imp_ratio": p.early_stop_imp_ratio, "adv_norm": p.adv_norm})
        ref_itf = copy.deepcopy(actor_itf)
        ref_itf.args["enable_save"] = False
        critic_kw = {k: v for k, v in kw.items() if k != "eps_clip"}
        critic_itf = ModelInterfaceAbstraction("ppo_critic", args=copy.deepcopy(critic_kw))
        rw_itf = ModelInterfaceAbstraction("paired_rw", args=dict(enable_save=False, output_scaling=p.reward_output_scaling,
                                                                  output_bias=p.reward_output_bias))
        n = self.dataset.train_bs_n_seqs
        # greedy decoding filters nothing, so generation emits no keep-mask (and nobody must wait for one)
        mask_keys = () if (p.gen.force_no_logits_mask or p.gen.greedy or p.gen.temperature == 0.0) else ("packed_logits_mask",)
        return {
            "actor_gen": MFCDef(name="actor_gen", n_seqs=n, interface_type=T.GENERATE, interface_impl=actor_itf, model_name="actor",
                                input_keys=("packed_prompts",),
                                output_keys=("seq_no_eos_mask", "packed_input_ids", "packed_logprobs", "prompt_mask") + mask_keys,
                                balanced_dp=True, model_type=self.actor.type, model_path=self.actor.path, n_mbs=self.actor_gen.n_mbs),
            "rew_inf": MFCDef(name="rew_inf", n_seqs=n, interface_type=T.INFERENCE, interface_impl=rw_itf, model_name="reward",
                              input_keys=("packed_input_ids",), output_keys=("rewards",), model_type=self.rew.type,
                              model_path=self.rew.path, n_mbs=self.rew_inf.n_mbs),
            "ref_inf": MFCDef(name="ref_inf", n_seqs=n, interface_type=T.INFERENCE, interface_impl=ref_itf, model_name="ref",
                              input_keys=("packed_input_ids",) + mask_keys, output_keys=("packed_ref_logprobs",),
                              model_type=self.ref.type, model_path=self.ref.path, n_mbs=self.ref_inf.n_mbs),
            "critic_inf": MFCDef(name="critic_inf", n_seqs=n, interface_type=T.INFERENCE, interface_impl=critic_itf, model_name="critic",
                                 input_keys=("packed_input_ids", "seq_no_eos_mask"), output_keys=("values",), model_type=self.critic.type,
                                 model_path=self.critic.path, n_mbs=self.critic_inf.n_mbs),
            "actor_train": MFCDef(name="actor_train", n_seqs=n, interface_type=T.TRAIN_STEP, interface_impl=actor_itf, model_name="actor",
                                  input_keys=("packed_input_ids", "packed_logprobs", "packed_ref_logprobs", "rewards", "values",
                                              "prompt_mask", "seq_no_eos_mask") + mask_keys, log_return_value=True,
                                  model_type=self.actor.type, model_path=self.actor.path, n_mbs=self.actor_train.n_mbs),
            "critic_train": MFCDef(name="critic_train", n_seqs=n, interface_type=T.TRAIN_STEP, interface_impl=critic_itf, model_name="critic",
                                   input_keys=("packed_input_ids", "packed_logprobs", "packed_ref_logprobs", "rewards", "values",
                                               "prompt_mask", "seq_no_eos_mask"), log_return_value=True,
                                   model_type=self.critic.type, model_path=self.critic.path, n_mbs=self.critic_train.n_mbs),
        }

    @property
    def allocations(self):
        return {"actor_gen": self.actor_gen, "actor_train": self.actor_train, "critic_inf": self.critic_inf,
                "critic_train": self.critic_train, "ref_inf": self.ref_inf, "rew_inf": self.rew_inf}

    @property
    def datasets(self):
        return [DatasetAbstraction("prompt", args=dict(dataset_path=self.dataset.path, max_length=self.dataset.max_prompt_len,
                                                       pad_to_max_length=self.dataset.pad_to_max_length))]

    @property
    def tokenizer_name_or_path(self):
        return self.actor.path

    @property
    def max_prompt_len(self):
        return self.dataset.max_prompt_len

    @property
    def search_kwargs(self):
        return dict(num_gen_tokens=self.ppo.gen.max_new_tokens, n_ppo_minibatches=self.ppo.ppo_n_minibatches)


@dataclasses.dataclass
class GenerationConfig(CommonExperimentConfig):
    model: ModelTrainEvalConfig = dataclasses.field(default_factory=ModelTrainEvalConfig)
    gen: GenerationHyperparameters = dataclasses.field(default_factory=GenerationHyperparameters)
    dataset: PromptOnlyDatasetConfig = dataclasses.field(default_factory=PromptOnlyDatasetConfig)
    allocation: MFCConfig = dataclasses.field(default_factory=MFCConfig)
    output_file: Optional[str] = None

    @property
    def models(self):
        return {"default": self.model}

    @property
    def rpcs(self):
        itf = ModelInterfaceAbstraction("generation", args=dict(generation_config=dataclasses.asdict(self.gen), output_file=self.output_file))
        return {"default": MFCDef(name="default", n_seqs=self.dataset.train_bs_n_seqs, interface_type=T.GENERATE, interface_impl=itf,
                                  model_name="default", input_keys=("packed_prompts",), balanced_dp=True, log_return_value=True,
                                  model_type=self.model.type, model_path=self.model.path, n_mbs=self.allocation.n_mbs)}

    @property
    def allocations(self):
        return {"default": self.allocation}

    @property
    def datasets(self):
        return [DatasetAbstraction("prompt", args=dict(dataset_path=self.dataset.path, max_length=self.dataset.max_prompt_len))]

    @property
    def tokenizer_name_or_path(self):
        return self.model.path

    @property
    def max_prompt_len(self):
        return self.dataset.max_prompt_len


for _name, _cls in (("sft", SFTConfig), ("rw", RWConfig), ("dpo", DPOConfig), ("ppo", PPOConfig), ("gen", GenerationConfig)):
    register_quickstart_exp(_name, _cls)
