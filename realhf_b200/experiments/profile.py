"""`quickstart profile`: time MFC handles on mock data over a grid of batch shapes (parity: `experiments/benchmark/
profile_exp.py` + `examples/profiling/profile.sh`: cartesian product of batch size x sequence length x n_mbs x handle for one
model, fabricating inputs with the interfaces' `_mock_*` hooks).

The reference reconfigures its workers between grid points inside one launch; here the sweep runs in-process on the local
device (every handle is an ordinary `ModelInterface` call), which is what the allocation search's cost model consumes."""

from __future__ import annotations

import dataclasses
import json
import time
from typing import Any, Dict, List, Optional

import torch

from realhf_b200.api.config import ModelInterfaceAbstraction, ModelName
from realhf_b200.api.data import SequenceSample
from realhf_b200.api.model import FinetuneSpec, GenerationHyperparameters, make_backend, make_interface
from realhf_b200.api.quickstart import ModelTrainEvalConfig, register_quickstart_exp


@dataclasses.dataclass
class ProfileConfig:
    experiment_name: str = "profile"
    trial_name: str = "run"
    model: ModelTrainEvalConfig = dataclasses.field(default_factory=ModelTrainEvalConfig)
    interface: str = "ppo_actor"                     # registered interface name
    interface_kwargs: Dict[str, Any] = dataclasses.field(default_factory=dict)
    handles: List[str] = dataclasses.field(default_factory=lambda: ["inference", "train_step"])
    batch_sizes: List[int] = dataclasses.field(default_factory=lambda: [8])
    seqlens: List[int] = dataclasses.field(default_factory=lambda: [256])
    n_mbs: List[int] = dataclasses.field(default_factory=lambda: [1])
    gen: GenerationHyperparameters = dataclasses.field(default_factory=lambda: GenerationHyperparameters(max_new_tokens=32, min_new_tokens=32))
    repeats: int = 3
    device: str = "cuda"
    dtype: str = "bf16"
    output_file: Optional[str] = None

    def run_local(self) -> List[Dict[str, Any]]:
        import realhf_b200.engine.engine  # noqa: F401  (backends)
        import realhf_b200.interfaces.basic  # noqa: F401
        import realhf_b200.interfaces.ppo  # noqa: F401
        from realhf_b200.models.factory import make_real_model
        dev = torch.device(self.device if (self.device != "cuda" or torch.cuda.is_available()) else "cpu")
        dtype = self.dtype if dev.type == "cuda" else "fp32"
        model = make_real_model(ModelName("default", 0), dev, self.model.path, self.model.type.is_critic,
                                init_from_scratch=self.model.init_from_scratch or not self.model.path, dtype=dtype,
                                hf_model_family=self.model.type._class)
        train = "train_step" in self.handles
        opt = dataclasses.asdict(self.model.optimizer) if self.model.optimizer is not None else {}
        backend = make_backend(ModelInterfaceAbstraction("train", dict(optimizer=opt)) if train else ModelInterfaceAbstraction("inference", {}))
        model = backend.initialize(model, FinetuneSpec(1, 1000, 1000))
        kw = dict(self.interface_kwargs)
        if self.interface in ("ppo_actor", "generation") and "generation_config" not in kw:
            kw["generation_config"] = dataclasses.asdict(self.gen)
        itf = make_interface(ModelInterfaceAbstraction(self.interface, kw))
        vocab = model.module_config.vocab_size
        rows = []
        for bs in self.batch_sizes:
            for sl in self.seqlens:
                for mbs in self.n_mbs:
                    for h in self.handles:
                        times = []
                        for r in range(self.repeats + 1):
                            ids = torch.randint(2, vocab, (bs * sl,), device=dev)
                            key = "packed_prompts" if h == "generate" else "packed_input_ids"
                            data = SequenceSample.from_default(seqlens=[sl] * bs, ids=[f"{r}-{i}" for i in range(bs)], data={key: ids})
                            data = itf.mock(h, model, data)
                            if dev.type == "cuda":
                                torch.cuda.synchronize()
                            t0 = time.perf_counter()
                            getattr(itf, h)(model, data, n_mbs=mbs)
                            if dev.type == "cuda":
                                torch.cuda.synchronize()
                            if r > 0:  # first call warms up (allocator, graph capture, lazy init)
                                times.append(time.perf_counter() - t0)
                        row = dict(handle=h, interface=self.interface, bs=bs, seqlen=sl, n_mbs=mbs, secs=min(times),
                                   tokens_per_s=bs * sl / min(times))
                        rows.append(row)
                        print(json.dumps(row), flush=True)
        if self.output_file:
            with open(self.output_file, "w") as f:
                json.dump(rows, f, indent=1)
        return rows


register_quickstart_exp("profile", ProfileConfig)
