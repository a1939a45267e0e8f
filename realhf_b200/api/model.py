"""Model-side contracts: configs, Model / Engine / Backend / Interface ABCs and registries.

Parity: `realhf/api/core/model_api.py` (GenerationHyperparameters :25-94, ReaLMoEConfig :98-140,
ReaLModelConfig :144-264, ModelVersion, FinetuneSpec, PipelinableEngine :305-461, Model :465-510,
ModelBackend :513-545, ModelInterface :564-632, registries :635-738).
"""

from __future__ import annotations

import abc
import dataclasses
from typing import Any, Callable, Dict, List, Optional, Union

import torch

from realhf_b200.api.config import (ModelAbstraction, ModelBackendAbstraction, ModelInterfaceAbstraction, ModelName,
                                    ModelWrapperAbstraction)
from realhf_b200.api.data import SequenceSample


@dataclasses.dataclass
class GenerationHyperparameters:
    max_new_tokens: int = 256
    min_new_tokens: int = 256
    greedy: bool = False
    top_p: float = 0.9
    top_k: int = 200
    temperature: float = 1.0
    use_cuda_graph: bool = False
    force_cudagraph_recapture: bool = True
    force_no_logits_mask: bool = False
    # B200 addition (no reference counterpart): decode with e4m3 weights / activations on the fp8 tensor cores.  Sampled
    # tokens and the returned log-probs are those of the quantised policy (see docs/generation_fp8.md).
    fp8_weights: bool = False

    def __post_init__(self):
        if self.temperature == 0.0:
            self.greedy, self.temperature = True, 1.0
        if not (0.0 < self.top_p <= 1.0):
            raise ValueError("top_p must be in (0, 1]")
        if self.top_k <= 0:
            raise ValueError("top_k must be positive")
        if self.min_new_tokens > self.max_new_tokens:
            raise ValueError("min_new_tokens > max_new_tokens")


@dataclasses.dataclass
class ReaLMoEConfig:
    num_experts: int = 8
    top_k: int = 2
    routing_type: str = "aux_loss"  # aux_loss | sinkhorn | none
    aux_loss_coeff: float = 1e-3
    capacity_factor: Optional[float] = None
    pad_to_capacity: bool = False   # accepted for compatibility: dropped assignments get a zero routing weight and the grouped GEMM
                                    # takes device-side row offsets, so no expert batch needs padding to a static capacity
    token_drop_policy: str = "probs"  # probs | position
    z_loss_coeff: float = 0.0
    input_jitter_eps: float = 0.0
    use_grouped_gemm: bool = True
    expert_parallel: bool = False  # partition experts over the TP group (all-to-all dispatch); reference: not available


@dataclasses.dataclass
class ReaLModelConfig:
    n_layers: int
    n_kv_heads: int
    n_q_heads: int
    hidden_dim: int
    intermediate_dim: int
    vocab_size: int
    head_dim: Optional[int] = None
    n_positions: Optional[int] = None
    embd_pdrop: float = 0.1
    resid_pdrop: float = 0.1
    attn_pdrop: float = 0.1
    layer_norm_epsilon: float = 1e-5
    activation_function: str = "gelu"
    scale_attn_by_inverse_layer_idx: bool = True
    scale_attn_weights: bool = True
    use_attention_bias: bool = True
    use_attn_proj_bias: bool = True
    use_mlp_bias: bool = True
    layer_norm_type: Optional[str] = None  # None (LayerNorm) | "rms" | "gemma"
    mlp_type: Optional[str] = None  # None (plain MLP) | "llama" | "moe"
    apply_rotary: bool = False
    rotary_base: float = 10000.0
    rotary_interleaved: bool = False
    rotary_scaling: Optional[float] = None
    rotary_scaling_type: Optional[str] = None
    normalize_embed: bool = False
    abs_position_embedding_offset: int = 0
    do_layernorm_before: bool = True
    tied_embedding: bool = False
    sliding_window: Optional[int] = None
    moe: Optional[ReaLMoEConfig] = None
    is_critic: bool = False

    def __post_init__(self):
        if self.is_critic and self.tied_embedding:
            raise ValueError("a critic cannot tie its embedding and output weights")
        if self.head_dim is None:
            self.head_dim = self.hidden_dim // self.n_q_heads
        if isinstance(self.moe, dict):
            self.moe = ReaLMoEConfig(**self.moe)

    def n_params(self) -> int:
        """Parameter count (used by the PP partitioner, the FLOP model and the allocation search)."""
        h, f, v = self.hidden_dim, self.intermediate_dim, self.vocab_size
        qkv = h * self.head_dim * (self.n_q_heads + 2 * self.n_kv_heads)
        attn = qkv + self.n_q_heads * self.head_dim * h
        if self.mlp_type == "moe":
            mlp = self.moe.num_experts * 3 * h * f + h * self.moe.num_experts
        elif self.mlp_type == "llama":
            mlp = 3 * h * f
        else:
            mlp = 2 * h * f
        emb = v * h + (self.n_positions * h if (not self.apply_rotary and self.n_positions) else 0)
        head = h if self.is_critic else (0 if self.tied_embedding else v * h)
        return self.n_layers * (attn + mlp + 2 * h) + emb + head + h


@dataclasses.dataclass
class ModelVersion:
    epoch: int = 0
    epoch_step: int = 0
    global_step: int = 0


@dataclasses.dataclass
class FinetuneSpec:
    total_train_epochs: int
    total_train_steps: int
    steps_per_epoch: int


class PipelinableEngine(abc.ABC):
    """What an interface drives: the same four calls regardless of the (dp,tp,pp) layout behind them."""

    def train_batch(self, input_: SequenceSample, loss_fn: Callable, version_steps: int,
                    num_micro_batches: Optional[int] = None) -> Dict[str, Any]:
        raise NotImplementedError()

    def eval_batch(self, input_: SequenceSample, loss_fn: Callable, num_micro_batches: Optional[int] = None):
        raise NotImplementedError()

    def forward(self, input_: SequenceSample, num_micro_batches: Optional[int] = None,
                post_hook: Optional[Callable] = None, aggregate_fn: Callable = torch.cat):
        raise NotImplementedError()

    def generate(self, input_: SequenceSample, tokenizer, gconfig: GenerationHyperparameters = None,
                 num_micro_batches: Optional[int] = None):
        raise NotImplementedError()


@dataclasses.dataclass
class Model:
    """A model shard plus what travels with it (tokenizer, device, version counter, backend name)."""

    name: ModelName
    module: Any
    tokenizer: Any
    device: Union[str, torch.device]
    dtype: Optional[torch.dtype] = None
    version: ModelVersion = dataclasses.field(default_factory=ModelVersion)
    ft_spec: Optional[FinetuneSpec] = None
    backend_name: Optional[str] = None
    module_config: Optional[ReaLModelConfig] = None
    hf_family: Optional[str] = None

    def __post_init__(self):
        try:
            self.module = self.module.to(self.device)
        except (ValueError, AttributeError, NotImplementedError):
            pass

    def inc_version(self):
        self.version.global_step += 1
        self.version.epoch_step += 1
        if self.ft_spec and self.version.epoch_step >= self.ft_spec.steps_per_epoch:
            self.version.epoch += 1
            self.version.epoch_step = 0


class ModelBackend(abc.ABC):
    """Wraps `model.module` into a PipelinableEngine (optimizer, sharding, schedules)."""

    @abc.abstractmethod
    def _initialize(self, model: Model, spec: FinetuneSpec) -> Model:
        ...

    def initialize(self, model: Model, spec: FinetuneSpec) -> Model:
        model.ft_spec = spec
        return self._initialize(model, spec)

    def destroy(self, model: Model):
        pass

    def save(self, model: Model, save_dir: str):
        """Optimizer / scheduler state (the reference saves none; we do for `recover`)."""

    def load(self, model: Model, load_dir: str):
        pass


class NullBackend(ModelBackend):
    def _initialize(self, model, spec):
        return model


def null_model(name: ModelName, device) -> Model:
    return Model(name, torch.nn.Identity(), None, device)


def tokenizer_only_model(name: ModelName, device, tokenizer_path: str) -> Model:
    from realhf_b200.api.data import load_hf_tokenizer
    return Model(name, torch.nn.Identity(), load_hf_tokenizer(tokenizer_path), device)


class ModelInterface(abc.ABC):
    """An algorithm's view of a model: what `generate / inference / train_step` mean for it."""

    def save(self, model: Model, save_dir: str):
        pass

    def state_dict(self) -> Dict:
        """Algorithm state that lives outside the model (KL controller, value normaliser, ...): written with the recover
        states and restored on a recover run.  (The reference loses this state on recovery)"""
        return {}

    def load_state_dict(self, sd: Dict):
        pass

    def evaluate(self, model: Model, eval_dataloader) -> Dict:
        return {}

    def inference(self, model: Model, data: SequenceSample, n_mbs: Optional[int] = None) -> Optional[SequenceSample]:
        raise NotImplementedError()

    def generate(self, model: Model, data: SequenceSample, n_mbs: Optional[int] = None) -> Optional[SequenceSample]:
        raise NotImplementedError()

    def train_step(self, model: Model, data: SequenceSample, n_mbs: Optional[int] = None) -> Dict:
        raise NotImplementedError()

    # used by the profiler to fabricate inputs of the right shape
    def _mock_generate(self, model: Model, data: SequenceSample):
        return data

    def _mock_inference(self, model: Model, data: SequenceSample):
        return data

    def _mock_train_step(self, model: Model, data: SequenceSample):
        return data

    def mock(self, type_: str, model: Model, data: SequenceSample) -> SequenceSample:
        return getattr(self, f"_mock_{type_}")(model, data)


class NullInterface(ModelInterface):
    def inference(self, model, data, n_mbs=None):
        scores = torch.zeros(data.bs, dtype=torch.float32, device=model.device)
        return SequenceSample.from_default(seqlens=[1] * data.bs, ids=data.ids, data={"rewards": scores})

    def train_step(self, model, data, n_mbs=None):
        model.inc_version()
        return {}

    def save(self, model, save_dir):
        pass


ALL_MODEL_CLASSES: Dict[str, Callable] = {}
ALL_INTERFACE_CLASSES: Dict[str, Callable] = {}
ALL_BACKEND_CLASSES: Dict[str, Callable] = {}
ALL_WRAPPER_CLASSES: Dict[str, Callable] = {}
SUPPORTED_HF_FAMILIES: Dict[str, "HFFamilySpec"] = {}


def register_model(name: str, fn):
    if name in ALL_MODEL_CLASSES:
        raise KeyError(f"model `{name}` already registered")
    ALL_MODEL_CLASSES[name] = fn


def register_interface(name: str, cls):
    if name in ALL_INTERFACE_CLASSES:
        raise KeyError(f"interface `{name}` already registered")
    ALL_INTERFACE_CLASSES[name] = cls


def register_backend(name: str, cls):
    if name in ALL_BACKEND_CLASSES:
        raise KeyError(f"backend `{name}` already registered")
    ALL_BACKEND_CLASSES[name] = cls


def register_wrapper(name: str, cls):
    ALL_WRAPPER_CLASSES[name] = cls


def make_model_wrapper(cfg: ModelWrapperAbstraction) -> Callable[[Model], Model]:
    return ALL_WRAPPER_CLASSES[cfg.type_](**cfg.args)


def make_model(cfg: ModelAbstraction, name: ModelName, device) -> Model:
    model = ALL_MODEL_CLASSES[cfg.type_](**cfg.args, name=name, device=device)
    for w in cfg.wrappers:
        model = make_model_wrapper(w)(model)
    return model


def make_interface(cfg: ModelInterfaceAbstraction) -> ModelInterface:
    return ALL_INTERFACE_CLASSES[cfg.type_](**cfg.args)


def make_backend(cfg: ModelBackendAbstraction) -> ModelBackend:
    return ALL_BACKEND_CLASSES[cfg.type_](**cfg.args)


register_backend("null", NullBackend)
register_model("null", null_model)
register_model("tokenizer", tokenizer_only_model)
register_interface("null", NullInterface)


@dataclasses.dataclass
class HFFamilySpec:
    """How one HuggingFace family maps to ReaLModelConfig and parameter names (api/from_hf/*)."""

    name: str
    hf_cls_name: str
    config_from_hf: Callable[[Any, bool], ReaLModelConfig]
    config_to_hf: Callable[[ReaLModelConfig], Any]
    sd_from_hf: Callable[[Dict[str, torch.Tensor], ReaLModelConfig], Dict[str, torch.Tensor]]
    sd_to_hf: Callable[[Dict[str, torch.Tensor], ReaLModelConfig], Dict[str, torch.Tensor]]
    embedding_param_names: Callable[[ReaLModelConfig], List[str]]
    tblock_param_names: Callable[[ReaLModelConfig, int], List[str]]
    head_param_names: Callable[[ReaLModelConfig], List[str]]
    make_test_config: Optional[Callable[[], ReaLModelConfig]] = None
    real_config_maker: Optional[Callable] = None


def register_hf_family(name: str, hf_cls_name: str, config_from_hf_converter, config_to_hf_converter,
                       sd_from_hf_converter, sd_to_hf_converter, embedding_param_names, tblock_param_names,
                       head_param_names, real_config_maker=None, make_test_config=None):
    if name in SUPPORTED_HF_FAMILIES:
        raise KeyError(f"HF family `{name}` already registered")
    SUPPORTED_HF_FAMILIES[name] = HFFamilySpec(name, hf_cls_name, config_from_hf_converter, config_to_hf_converter,
                                               sd_from_hf_converter, sd_to_hf_converter, embedding_param_names,
                                               tblock_param_names, head_param_names, make_test_config,
                                               real_config_maker)
