"""SPMD dataflow executor: every rank walks the MFC graph itself.

When every MFC of an experiment runs on the same device mesh (the data-parallel-everything allocation
that 180 GB B200s make optimal for <=13B models, and any single-GPU run), there is nothing for a
central master to arbitrate: each rank holds its DP shard of the batch on the device and executes the
MFCs in topological order.  This removes the ZMQ round trips, the metadata buffer and the broadcast-based
data transfer of the reference's master/worker runtime (`system/master_worker.py`, `model_worker.py`) from
the hot loop; asymmetric allocations (different meshes / layouts per MFC, parameter reallocation across
meshes) go through `realhf_b200.system.master_worker` instead.

Per-MFC device time is measured with CUDA events on the launching stream (max over ranks is taken by
the caller), matching the north-star metric definition.
"""

from __future__ import annotations

import dataclasses
import time
from typing import Any, Dict, List

import networkx as nx
import torch

from realhf_b200.api.data import SequenceSample
from realhf_b200.api.dfg import MFCDef, build_graph
from realhf_b200.api.model import Model, ModelInterface


def all_gather_sample(sample: SequenceSample, group, group_size: int, group_rank: int):
    """Concatenate the members' samples of a process group (in group-rank order) on every member.

    Used when an MFC runs on a tensor/pipeline-parallel layout inside an otherwise data-parallel pool: the TP (x PP)
    peers each hold a different DP shard of the batch, the model-parallel call needs the union.  Metadata travels by
    `all_gather_object`, tensors by one padded `all_gather` per key on the device.  Returns (gathered, first, count):
    my items are `gathered[first: first + count]`."""
    import torch.distributed as dist
    if group_size == 1:
        return sample, 0, sample.bs
    metas: List[Any] = [None] * group_size
    dist.all_gather_object(metas, sample.meta(), group=group)
    parts: List[Dict[str, torch.Tensor]] = [dict() for _ in range(group_size)]
    for k in sorted(sample.keys):
        mine = sample.data[k]
        if mine is None:
            for prt in parts:
                prt[k] = None
            continue
        lens = [m.total_len(k) for m in metas]
        mx = max(lens)
        buf = mine.new_zeros((mx, *mine.shape[1:]))
        buf[: mine.shape[0]] = mine
        out = [torch.empty_like(buf) for _ in range(group_size)]
        dist.all_gather(out, buf, group=group)
        for r in range(group_size):
            parts[r][k] = out[r][: lens[r]]
    samples = []
    for r, m in enumerate(metas):
        m.data = parts[r]
        samples.append(m)
    with SequenceSample.disable_validation():
        gathered = SequenceSample.gather(samples, keys=sample.keys)
    first = sum(m.bs for m in metas[:group_rank])
    return gathered, first, metas[group_rank].bs


@dataclasses.dataclass
class MFCRecord:
    name: str
    device_ms: float
    wall_ms: float
    result: Any


class SPMDExecutor:
    def __init__(self, rpcs: List[MFCDef], models: Dict[str, Model], interfaces: Dict[str, ModelInterface], device,
                 time_mfcs: bool = True):
        self.rpcs = rpcs
        self.G = build_graph(rpcs)
        self.order = [self.G.nodes[n]["object"] for n in nx.topological_sort(self.G)]
        self.models = models          # keyed by role (or str(ModelName))
        self.interfaces = interfaces  # keyed by MFC name
        self.device = torch.device(device)
        self.time_mfcs = time_mfcs and self.device.type == "cuda"
        self.hooks: Dict[str, List] = {}  # rpc name -> callables run before it (param realloc / offload reload)
        self.post_hooks: Dict[str, List] = {}
        # rpc name -> (group, size, rank): MFCs on a model-parallel layout gather their inputs over that group first
        self.regroup: Dict[str, Any] = {}

    def _model(self, rpc: MFCDef) -> Model:
        for k in (rpc.name, str(rpc.model_name), rpc.model_name.role):  # an MFC-specific replica (own layout) wins
            if k in self.models:
                return self.models[k]
        raise KeyError(f"no model for {rpc.model_name}")

    def add_layout_replica(self, rpc_name: str, src: Model, dst_ctx, src_topo, dst_topo, workers: List[int], my_worker: int,
                           fused_tp: bool = True) -> Model:
        """Run MFC `rpc_name` on its own parallel layout (the reference's parameter reallocation, realhf/impl/model/comm/
        param_realloc.py + system/model_worker.py:__param_realloc): builds an inference replica of `src` sharded per
        `dst_ctx`, a pre-hook that refreshes its weights from the source layout with the segment-copy realloc plan (all
        copies are local when the source is replicated over DP), and the input/output regrouping over the TP group."""
        from realhf_b200.api.config import ModelName
        from realhf_b200.engine.engine import InferenceBackend
        from realhf_b200.models.real_model import ReaLModel
        from realhf_b200.parallel import realloc
        assert dst_ctx.pp_size == 1, "SPMD replicas support tensor/data-parallel layouts; pipeline layouts use the master/worker runtime"
        eng = src.module
        real: ReaLModel = eng.module if hasattr(eng, "module") and not isinstance(eng, ReaLModel) else eng
        m = ReaLModel(real.config, dst_ctx, dtype=real.dtype, device=self.device)
        m.attach_flat(torch.zeros(m.flat_numel, dtype=real.dtype, device=self.device))
        for prm in m.parameters():
            prm.requires_grad_(False)
        if fused_tp and dst_ctx.tp_size > 1 and self.device.type == "cuda":
            from realhf_b200.parallel.fused_tp import FusedTP
            dst_ctx.symm = FusedTP(dst_ctx, max_tokens=256, max_features=real.config.hidden_dim, device=self.device)
        replica = InferenceBackend().initialize(Model(ModelName(src.name.role, src.name.replica_id + 1), m, src.tokenizer, self.device), None)
        plan = realloc.derive_plan(real.config, src_topo, workers, dst_topo, workers, for_worker=my_worker)
        exe = realloc.ReallocExecutor(plan, my_worker, torch.tensor([], dtype=real.dtype).element_size(), self.device)
        self.models[rpc_name] = replica
        self.hooks.setdefault(rpc_name, []).append(lambda: exe.run(real.flat_param.data, m.flat_param.data))
        if dst_ctx.tp_size > 1:
            self.regroup[rpc_name] = (dst_ctx.tp_group, dst_ctx.tp_size, dst_ctx.tp_rank)
        return replica

    def run_step(self, batch: SequenceSample) -> Dict[str, MFCRecord]:
        """`batch` carries this rank's shard of the dataset keys (tensors already on the device)."""
        pool = batch
        records: Dict[str, MFCRecord] = {}
        for rpc in self.order:
            for h in self.hooks.get(rpc.name, []):
                h()
            with SequenceSample.disable_validation():
                inp = SequenceSample.gather([pool], keys=rpc.input_keys)
            if rpc.input_key_remap:
                inp.remap_keys_(rpc.input_key_remap)
            itf, model = self.interfaces[rpc.name], self._model(rpc)
            t0 = time.perf_counter()
            if self.time_mfcs:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            rg = self.regroup.get(rpc.name)
            if rg is not None:
                inp, first, count = all_gather_sample(inp, *rg)
            res = getattr(itf, rpc.interface_type.value)(model, inp, n_mbs=rpc.n_mbs)
            if rg is not None and isinstance(res, SequenceSample):
                res = res.select(range(first, first + count))  # every peer computed the whole group's batch; keep my shard
            if self.time_mfcs:
                e1.record()
            if isinstance(res, SequenceSample):
                if rpc.output_key_remap:
                    res.remap_keys_(rpc.output_key_remap)
                with SequenceSample.disable_validation():
                    res = SequenceSample.gather([res], keys=[k for k in rpc.output_keys if k in res.keys] or None)
                pool.update_(res)
            for h in self.post_hooks.get(rpc.name, []):
                h()
            records[rpc.name] = MFCRecord(rpc.name, (e0, e1) if self.time_mfcs else 0.0, 0.0, res)
            records[rpc.name].wall_ms = (time.perf_counter() - t0) * 1e3
        if self.time_mfcs:
            torch.cuda.synchronize(self.device)
            for r in records.values():
                e0, e1 = r.device_ms
                r.device_ms = e0.elapsed_time(e1)
        self.last_pool = pool
        return records
