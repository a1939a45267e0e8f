"""`quickstart profile`: time MFC handles on mock data over a grid of batch shapes (parity: `experiments/benchmark/
profile_exp.py` + `examples/profiling/profile.sh`: cartesian product of batch size x sequence length x n_mbs x handle for one
model, fabricating inputs with the interfaces' `_mock_*` hooks).

Like the reference (`profile_exp.py:172-251`) ONE launch sweeps parallel layouts as well: `layouts=[d1m1p1,d2m1p1,d1m2p1,...]`
(the `d{dp}m{tp}p{pp}` syntax of `allocation_mode`).  A single-device layout runs in-process; every multi-device layout gets
its own process group (one process per device, NCCL on GPUs / gloo on CPU) that walks the same batch-shape grid, and a grid
point's time is the maximum over its ranks.  `bs` is the GLOBAL batch of a grid point (split over dp).  The rows feed the
allocation search's cost model (`search/engine.py`)."""

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
    layouts: List[str] = dataclasses.field(default_factory=lambda: ["d1m1p1"])   # d{dp}m{tp}p{pp}, swept inside one launch

    def run_local(self) -> List[Dict[str, Any]]:
        """All layouts x batch shapes x handles.  Returns the rows (also printed as JSON lines / written to `output_file`)."""
        import re
        rows: List[Dict[str, Any]] = []
        for lay in self.layouts:
            mt = re.fullmatch(r"d(\d+)m(\d+)p(\d+)", lay)
            if mt is None:
                raise ValueError(f"layout `{lay}` is not of the form d<dp>m<tp>p<pp>")
            dp, tp, pp = (int(x) for x in mt.groups())
            world = dp * tp * pp
            if world == 1:
                part = self._run_grid(None)
            else:
                from realhf_b200.base.testing import run_distributed
                cuda = self.device == "cuda" and torch.cuda.is_available()
                if cuda and torch.cuda.device_count() < world:
                    print(json.dumps(dict(layout=lay, skipped=f"needs {world} GPUs, {torch.cuda.device_count()} visible")), flush=True)
                    continue
                per_rank = run_distributed(_layout_worker, world, backend="nccl" if cuda else "gloo", timeout=3600,
                                           cfg=self, layout=(pp, dp, tp))
                part = []
                for i, row in enumerate(per_rank[0]):  # a point costs what its slowest rank needs
                    secs = max(r[i]["secs"] for r in per_rank)
                    part.append(dict(row, secs=secs, tokens_per_s=row["bs"] * row["seqlen"] / secs))
            for row in part:
                row["layout"] = lay
                print(json.dumps(row), flush=True)
            rows += part
        if self.output_file:
            with open(self.output_file, "w") as f:
                json.dump(rows, f, indent=1)
        return rows

    def _run_grid(self, ctx, quiet: bool = True) -> List[Dict[str, Any]]:
        import realhf_b200.engine.engine  # noqa: F401  (backends)
        import realhf_b200.interfaces.basic  # noqa: F401
        import realhf_b200.interfaces.ppo  # noqa: F401
        from realhf_b200.models.factory import make_real_model
        dev = torch.device(self.device if (not self.device.startswith("cuda") or torch.cuda.is_available()) else "cpu")
        dtype = self.dtype if dev.type == "cuda" else "fp32"
        from realhf_b200.base import constants
        dp = 1 if ctx is None else ctx.dp_size
        with constants.model_scope(ModelName("default", 0), ctx, True) if ctx is not None else _null():
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
                        lbs = max(1, bs // dp)  # this rank's share of the global batch
                        for r in range(self.repeats + 1):
                            ids = torch.randint(2, vocab, (lbs * sl,), device=dev)
                            key = "packed_prompts" if h == "generate" else "packed_input_ids"
                            data = SequenceSample.from_default(seqlens=[sl] * lbs, ids=[f"{r}-{i}" for i in range(lbs)], data={key: ids})
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
        return rows


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def _layout_worker(rank: int, world: int, cfg: "ProfileConfig", layout):
    """One rank of a multi-device layout of the sweep (spawned by `ProfileConfig.run_local`)."""
    import copy

    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    pc = copy.deepcopy(cfg)
    pp, dp, tp = layout
    cuda = pc.device == "cuda" and torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(rank)
        pc.device = f"cuda:{rank}"
    ctx = ParallelContext.build(ProcessTopology(pp, dp, tp), list(range(world)), rank, backend="nccl" if cuda else "gloo",
                                sequence_parallel=tp > 1 and "train_step" in pc.handles)
    return pc._run_grid(ctx)


register_quickstart_exp("profile", ProfileConfig)
