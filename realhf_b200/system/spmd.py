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
from typing import Any, Dict, List, Optional

import networkx as nx
import torch

from realhf_b200.api.data import SequenceSample
from realhf_b200.api.dfg import MFCDef, build_graph
from realhf_b200.api.model import Model, ModelInterface


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

    def _model(self, rpc: MFCDef) -> Model:
        for k in (str(rpc.model_name), rpc.model_name.role):
            if k in self.models:
                return self.models[k]
        raise KeyError(f"no model for {rpc.model_name}")

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
            res = getattr(itf, rpc.interface_type.value)(model, inp, n_mbs=rpc.n_mbs)
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
