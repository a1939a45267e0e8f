"""Engines: the `PipelinableEngine` implementations behind the `inference` and `train` backends.

Parity: `backend/inference.py` (PipelinableInferenceEngine), `backend/megatron.py`
(ReaLMegatronEngine.train_batch/forward/generate) and `backend/deepspeed.py`.  One class covers pp == 1 and
pp > 1 (delegating to `engine.pipe_runner`), data / tensor / sequence parallel layouts, micro-batching,
and the sharded optimizer in `engine.optim`.
"""

from __future__ import annotations

import collections
from typing import Any, Callable, Dict, List, Optional

import torch
import torch.distributed as dist

from realhf_b200.api.data import SequenceSample
from realhf_b200.api.model import (FinetuneSpec, GenerationHyperparameters, Model, ModelBackend, PipelinableEngine,
                                   register_backend)
from realhf_b200.engine.optim import FlatAdamW, OptimizerConfig
from realhf_b200.models import generation as gen
from realhf_b200.models.real_model import ModelOutput, ReaLModel

MAIN_KEY = "packed_input_ids"


def _mb_inputs(mb: SequenceSample, device, key: str = MAIN_KEY):
    lens = mb.flat_seqlens(key)
    cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
    cu[1:] = torch.tensor(lens, dtype=torch.int32).cumsum(0)
    return mb.data[key], cu.to(device, non_blocking=True), max(lens)


from realhf_b200.utils.padding import pad_sequence_parallel_input as pad_for_sp  # noqa: E402  (reference: nn/real_llm_api.py:392-433)


class ReaLEngine(PipelinableEngine):
    def __init__(self, model: ReaLModel, optimizer: Optional[FlatAdamW] = None):
        self.module = model
        self.optim = optimizer
        self.ctx = model.ctx
        self._pipe = None
        self._gen_state: Optional[gen.DecodeState] = None
        if self.ctx.pp_size > 1:
            from realhf_b200.engine.pipe_runner import PipelineRunner
            self._pipe = PipelineRunner(self)

    # convenience passthroughs used by interfaces
    def train(self, mode: bool = True):
        self.module.train(mode)
        return self

    def eval(self):
        self.module.eval()
        return self

    @property
    def config(self):
        return self.module.config

    def _check_ep(self):
        """The fused expert-parallel exchange sizes its receive buffers for `cap_factor` x the balanced load and reports an
        overflow through a device flag: read it once per engine call (the only host sync of the MoE path)."""
        ep = getattr(self.ctx, "_fused_ep", None)
        if ep is not None:
            ep.raise_if_overflow()

    # ------------------------------------------------------------------ single-stage forward
    def _forward_mb(self, mb: SequenceSample) -> ModelOutput:
        m = self.module
        ids, cu, mx = _mb_inputs(mb, m.device)
        n_pad = 0
        if m.sequence_parallel:
            ids, cu, mx, n_pad = pad_for_sp(ids, cu, mx, self.ctx.tp_size)
        out = m(input_ids=ids, cu_seqlens=cu, max_seqlen=mx)
        if n_pad:
            out.hidden = out.hidden[: out.hidden.shape[0] - n_pad]
        return out

    # ------------------------------------------------------------------ API
    def train_batch(self, input_: SequenceSample, loss_fn: Callable, version_steps: int,
                    num_micro_batches: Optional[int] = None) -> Dict[str, Any]:
        assert self.optim is not None, "train_batch needs the `train` backend"
        n_mbs = num_micro_batches or 1
        self.optim.begin_call(train=True)
        self.optim.zero_grad()
        if self._pipe is not None:
            stats = self._pipe.train_batch(input_, loss_fn, n_mbs)
        else:
            stats: Dict[str, Any] = collections.defaultdict(float)
            mbs = input_.split(min(n_mbs, input_.bs))
            for j, mb in enumerate(mbs):
                self.optim.arm(j == len(mbs) - 1)  # last micro-batch: gradient buckets are reduced from inside backward
                out = self._forward_mb(mb)
                loss, st = loss_fn(out, mb)
                self.optim.scale_loss(loss / len(mbs)).backward()
                for k, v in st.items():
                    stats[k] = stats[k] + (v.detach() if torch.is_tensor(v) else v) / len(mbs)
        ost = self.optim.step(version_steps)
        self.optim.end_call()
        self._check_ep()
        stats = dict(stats)
        stats.update(ost)
        return stats

    @torch.no_grad()
    def eval_batch(self, input_: SequenceSample, loss_fn: Callable, num_micro_batches: Optional[int] = None):
        n_mbs = num_micro_batches or 1
        if self.optim is not None:
            self.optim.begin_call()
        try:
            if self._pipe is not None:
                return self._pipe.eval_batch(input_, loss_fn, n_mbs)
            stats: Dict[str, Any] = collections.defaultdict(float)
            mbs = input_.split(min(n_mbs, input_.bs))
            for mb in mbs:
                _, st = loss_fn(self._forward_mb(mb), mb)
                for k, v in st.items():
                    stats[k] = stats[k] + (v.detach() if torch.is_tensor(v) else v) / len(mbs)
            return dict(stats)
        finally:
            if self.optim is not None:
                self.optim.end_call()

    @torch.no_grad()
    def forward(self, input_: SequenceSample, num_micro_batches: Optional[int] = None,
                post_hook: Optional[Callable[[ModelOutput, SequenceSample], Any]] = None,
                aggregate_fn: Callable = torch.cat):
        """Inference over micro-batches; `post_hook(output, mb)` reduces each output (e.g. to log-probs)
        before aggregation so full logits never pile up (reference: backend/inference.py:96-124)."""
        n_mbs = num_micro_batches or 1
        if self.optim is not None:
            self.optim.begin_call()
        try:
            if self._pipe is not None:
                return self._pipe.forward(input_, n_mbs, post_hook, aggregate_fn)
            outs = []
            for mb in input_.split(min(n_mbs, input_.bs)):
                out = self._forward_mb(mb)
                outs.append(post_hook(out, mb) if post_hook is not None else out.logits)
            self._check_ep()
            return aggregate_fn(outs) if len(outs) > 1 else outs[0]
        finally:
            if self.optim is not None:
                self.optim.end_call()

    @torch.no_grad()
    def generate(self, input_: SequenceSample, tokenizer, gconfig: GenerationHyperparameters = None,
                 num_micro_batches: Optional[int] = None) -> List[gen.GenerationOutput]:
        """One GenerationOutput per micro-batch (their widths may differ); see `interfaces.ppo` for the packing."""
        gconfig = gconfig or GenerationHyperparameters()
        n_mbs = num_micro_batches or 1
        eos = getattr(tokenizer, "eos_token_id", None)
        pad = getattr(tokenizer, "pad_token_id", None)
        pad = pad if pad is not None else (eos if eos is not None else 0)
        if self.optim is not None:
            self.optim.materialize()  # ZeRO-3: the caller's next train/forward call releases again
        if self._pipe is not None:
            return self._pipe.generate(input_, gconfig, eos, pad, n_mbs)
        outs = []
        for mb in input_.split(min(n_mbs, input_.bs)):
            ids, cu, _ = _mb_inputs(mb, self.module.device)
            o, self._gen_state = gen.generate(self.module, ids, cu, gconfig, eos, pad, state=self._gen_state)
            outs.append(o)
        if gconfig.force_cudagraph_recapture:
            self._gen_state = None  # releases the KV cache between calls, like the reference
        return outs


# ------------------------------------------------------------------------------------------- backends


class InferenceBackend(ModelBackend):
    """name `inference`: forward / generate only."""

    def _initialize(self, model: Model, spec: FinetuneSpec) -> Model:
        m: ReaLModel = model.module
        for p in m.parameters():
            p.requires_grad_(False)
        model.module = ReaLEngine(m)
        model.backend_name = "inference"
        return model


class TrainBackend(ModelBackend):
    """name `train` (aliases `megatron`, `deepspeed` for config compatibility): sharded AdamW + schedules.

    `zero_stage` 1/2 shard optimizer state (+ gradients at reduce time) over DP; `offload_optimizer`
    streams the state from pinned host memory; `zero_stage=3` additionally keeps parameters sharded
    between steps (`engine.zero3`)."""

    def __init__(self, optimizer: Optional[dict] = None, zero_stage: int = 1, offload_optimizer: bool = False,
                 offload_param: bool = False, enable_fp16: bool = False, enable_bf16: bool = True, **_ignored):
        cfg = optimizer if isinstance(optimizer, OptimizerConfig) else OptimizerConfig(**(optimizer or {}))
        cfg.offload = cfg.offload or offload_optimizer
        cfg.zero_stage = max(cfg.zero_stage, zero_stage)
        cfg.offload_param = cfg.offload_param or offload_param
        self.cfg = cfg
        self.zero_stage = zero_stage
        self.offload_param = offload_param

    def _initialize(self, model: Model, spec: FinetuneSpec) -> Model:
        m: ReaLModel = model.module
        total = spec.total_train_steps if spec is not None else 1000
        opt = FlatAdamW(m, self.cfg, total_steps=total)
        opt.release()  # no-op unless ZeRO-3
        model.module = ReaLEngine(m, opt)
        model.backend_name = "train"
        return model

    def save(self, model: Model, save_dir: str):
        import os
        eng: ReaLEngine = model.module
        os.makedirs(save_dir, exist_ok=True)
        c = eng.ctx
        torch.save(eng.optim.state_dict(), os.path.join(save_dir, f"optim_pp{c.pp_rank}_tp{c.tp_rank}_dp{c.dp_rank}.pt"))

    def load(self, model: Model, load_dir: str):
        import os
        eng: ReaLEngine = model.module
        c = eng.ctx
        f = os.path.join(load_dir, f"optim_pp{c.pp_rank}_tp{c.tp_rank}_dp{c.dp_rank}.pt")
        if os.path.exists(f):
            eng.optim.load_state_dict(torch.load(f, map_location="cpu", weights_only=False))
        else:
            # e.g. this rank's worker was killed outright and could not dump its shard: the weights are restored from the HF
            # checkpoint, the Adam moments of this shard restart from zero
            from realhf_b200.base import logging
            logging.getLogger("engine").warning(f"no optimizer state for this rank in {load_dir}: moments of this shard are reset")


register_backend("inference", InferenceBackend)
register_backend("train", TrainBackend)
register_backend("megatron", TrainBackend)
register_backend("deepspeed", TrainBackend)
