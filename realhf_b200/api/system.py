"""System-level configuration: what the controller hands to each worker.

Parity: `realhf/api/core/system_api.py` — `Scheduling`, `ModelWorker`, `MasterWorker`,
`ExperimentSaveEvalControl`, `ExperimentConfig` (derives topologies, data-transfer pairs, param-realloc pairs,
which replica is instantiated, shard -> worker map) and the `Experiment` ABC + registry.
"""

from __future__ import annotations

import dataclasses
import os
from typing import Callable, Dict, List, Optional, Tuple, Union

from realhf_b200.api.config import (DataLoaderAbstraction, DatasetAbstraction, ModelName, ModelShardID,
                                    StandaloneModelShardAbstraction)
from realhf_b200.api.dfg import MFCDef, ParamReallocHook, build_graph
from realhf_b200.base.topology import ProcessTopology


@dataclasses.dataclass
class Scheduling:
    cpu: int = 4
    gpu: int = 0
    mem: int = 10000  # MB
    gpu_type: str = "tesla"
    node_type: Optional[str] = None
    nodelist: Optional[str] = None
    exclude: Optional[str] = None
    container_image: Optional[str] = None
    env_vars: Dict[str, str] = dataclasses.field(default_factory=dict)
    time_limit: Optional[str] = None
    begin: Optional[str] = None
    deadline: Optional[str] = None

    @staticmethod
    def master_worker_default(**kw):
        return Scheduling(**{"cpu": 8, "mem": 20000, **kw})

    @staticmethod
    def model_worker_default(**kw):
        return Scheduling(**{"cpu": 4, "gpu": 1, "mem": 60000, **kw})


@dataclasses.dataclass
class WorkerInformation:
    experiment_name: str = ""
    trial_name: str = ""
    worker_type: str = ""
    worker_index: int = -1
    worker_count: int = 0
    worker_tag: Optional[str] = None      # free-form label (reference schema; shows up in the pickled config only)
    host_key: Optional[str] = None        # reference schema: name-resolve key of the worker's host (workers publish their addresses
                                          # through the stream / process-group keys instead)
    watch_keys: Union[str, List[str], None] = None   # more keys whose disappearance ends the worker (besides the controller's lease)

    def system_setup(self, experiment_name, trial_name, worker_type, worker_index, worker_count):
        self.experiment_name, self.trial_name = experiment_name, trial_name
        self.worker_type, self.worker_index, self.worker_count = worker_type, worker_index, worker_count


@dataclasses.dataclass
class ExperimentSaveEvalControl:
    total_train_epochs: int = 1
    save_freq_epochs: Optional[int] = None
    save_freq_steps: Optional[int] = None
    save_freq_secs: Optional[int] = None
    eval_freq_epochs: Optional[int] = None
    eval_freq_steps: Optional[int] = None
    eval_freq_secs: Optional[int] = None
    benchmark_steps: Optional[int] = None
    # B200-native addition.  How many steps may be in flight in the master's dataflow walk: 2 lets MFCs of step s+1 whose inputs and
    # weights are ready (generation after `actor_train` of step s) run while the rest of step s (`critic_train`) is still busy on
    # other GPUs, like the reference's free-running request coroutines; 1 puts a barrier after every step.
    max_inflight_steps: int = 2


@dataclasses.dataclass
class ModelWorker:
    seed: int
    shards: List[StandaloneModelShardAbstraction]
    tokenizer_name_or_path: Optional[str] = None
    datasets: Optional[List[Union[str, DatasetAbstraction]]] = None
    dataloader: Union[str, DataLoaderAbstraction] = "packed"
    use_dataset_cache: bool = False
    cuda_cache_cleanliness: bool = True
    cuda_cache_clear_freq: int = 10
    backend: str = "nccl"  # process-group backend: nccl on GPUs, gloo for the CPU plumbing configuration
    device: str = "cuda"
    # filled by ExperimentConfig
    model_rpcs: Optional[List[MFCDef]] = None
    model_topos: Optional[Dict[ModelName, ProcessTopology]] = None
    msid2mwid: Optional[Dict[ModelShardID, int]] = None
    data_transfer_pairs: Optional[List[Tuple[ModelName, ModelName]]] = None
    sync_param_pairs: Optional[List[Tuple[ModelName, ModelName]]] = None
    profile_mode: bool = False   # inputs of every MFC are fabricated by `ModelInterface.mock` (profiling through the full runtime; the
                                 # `profile` quickstart experiment itself runs in-process, experiments/profile.py)
    worker_info: Optional[WorkerInformation] = None

    def __post_init__(self):
        names = [s.id.model_name for s in self.shards]
        if len(set(names)) != len(names):
            raise ValueError(f"a model worker cannot hold two shards of the same model: {names}")


@dataclasses.dataclass
class MasterWorker:
    exp_ctrl: ExperimentSaveEvalControl
    model_rpcs: List[MFCDef]
    n_model_workers: int
    model_topos: Dict[ModelName, ProcessTopology]
    msid2mwid: Optional[Dict[ModelShardID, int]] = None
    data_transfer_pairs: Optional[List[Tuple[ModelName, ModelName]]] = None
    sync_param_pairs: Optional[List[Tuple[ModelName, ModelName]]] = None
    worker_info: Optional[WorkerInformation] = None


@dataclasses.dataclass
class TasksGroup:
    count: int
    scheduling: Scheduling


@dataclasses.dataclass
class ExperimentScheduling:
    model_worker: TasksGroup
    master_worker: TasksGroup
    controller_image: Optional[str] = None   # reference schema (container of the controller job; the launcher runs where it is started)


@dataclasses.dataclass
class ExperimentConfig:
    exp_ctrl: ExperimentSaveEvalControl
    model_rpcs: List[MFCDef]
    model_worker: List[ModelWorker] = dataclasses.field(default_factory=list)
    master_worker: Optional[List[MasterWorker]] = None

    def __post_init__(self):
        G = build_graph(self.model_rpcs)
        names = sorted({s.id.model_name for w in self.model_worker for s in w.shards})
        rpc_names = {r.model_name for r in self.model_rpcs}
        if not rpc_names.issubset(names):
            raise ValueError(f"MFCs use models {rpc_names - set(names)} that no worker holds")
        for role in {n.role for n in names}:
            rids = sorted(n.replica_id for n in names if n.role == role)
            if rids != list(range(len(rids))):
                raise ValueError(f"replica ids of role `{role}` must be 0..k-1, got {rids}")
        topos = self._collect_topos(names)
        data_pairs = self._resolve_data_transfer_pairs()
        sync_pairs = self._resolve_param_realloc_pairs()
        inst = self._resolve_model_names_to_instantiate(names)
        msid2mwid: Dict[ModelShardID, int] = {}
        for i, mw in enumerate(self.model_worker):
            for s in mw.shards:
                s.should_instantiate = s.id.model_name in inst
                msid2mwid[s.id] = i
        for mw in self.model_worker:
            mw.model_topos, mw.msid2mwid = topos, msid2mwid
            mw.data_transfer_pairs, mw.sync_param_pairs = data_pairs, sync_pairs
            mw.model_rpcs = self.model_rpcs
        self.master_worker = [MasterWorker(exp_ctrl=self.exp_ctrl, model_rpcs=self.model_rpcs,
                                           n_model_workers=len(self.model_worker), model_topos=topos, msid2mwid=msid2mwid,
                                           data_transfer_pairs=data_pairs, sync_param_pairs=sync_pairs)]

    def _collect_topos(self, names) -> Dict[ModelName, ProcessTopology]:
        topos: Dict[ModelName, ProcessTopology] = {}
        for mw in self.model_worker:
            for s in mw.shards:
                t = topos.setdefault(s.id.model_name, s.id.topo)
                if t != s.id.topo:
                    raise ValueError(f"inconsistent topology for {s.id.model_name}")
        for n, t in topos.items():
            have = sum(1 for mw in self.model_worker for s in mw.shards if s.id.model_name == n)
            if have != t.world_size():
                raise ValueError(f"model {n}: {have} shards configured but its topology has {t.world_size()} ranks")
        return topos

    def _resolve_data_transfer_pairs(self) -> List[Tuple[ModelName, ModelName]]:
        """(producer model, consumer model) for every key flowing between MFCs; dataset keys come from the source MFC."""
        pairs = []
        G = self.model_rpcs[0]._G
        src = next(r for r in self.model_rpcs if r.is_src)
        for r in self.model_rpcs:
            for k in r.input_keys:
                prod = r.data_producers.get(k, src.model_name)
                if (prod, r.model_name) not in pairs:
                    pairs.append((prod, r.model_name))
        return pairs

    def _resolve_param_realloc_pairs(self) -> List[Tuple[ModelName, ModelName]]:
        pairs = []
        for r in self.model_rpcs:
            for h in r._pre_hooks + r._post_hooks:
                if isinstance(h, ParamReallocHook):
                    p = (h.source, r.model_name) if h.source is not None else (r.model_name, h.target)
                    if p not in pairs:
                        pairs.append(p)
        return pairs

    def _resolve_model_names_to_instantiate(self, names) -> List[ModelName]:
        """Of all replicas of a role only ONE owns memory at start: the trainable one (the rest receive weights
        by realloc).  Roles without a train MFC instantiate replica 0."""
        from realhf_b200.api.config import ModelInterfaceType
        out = []
        for role in sorted({n.role for n in names}):
            replicas = [n for n in names if n.role == role]
            train = [r.model_name for r in self.model_rpcs if r.role == role and r.interface_type == ModelInterfaceType.TRAIN_STEP]
            linked = any(role in (a.role, b.role) and a.role == b.role for a, b in self._resolve_param_realloc_pairs())
            if not linked:
                out += replicas  # independent replicas (no realloc between them): every one loads its own weights
            else:
                out.append(train[0] if train else ModelName(role, 0))
        return out

    def set_worker_information(self, experiment_name: str, trial_name: str):
        for i, mw in enumerate(self.model_worker):
            mw.worker_info = WorkerInformation(experiment_name, trial_name, "model_worker", i, len(self.model_worker))
        for i, m in enumerate(self.master_worker):
            m.worker_info = WorkerInformation(experiment_name, trial_name, "master_worker", i, 1)


class Experiment:
    """Produces the scheduling request and the ExperimentConfig of a run."""

    def scheduling_setup(self) -> ExperimentScheduling:
        raise NotImplementedError()

    def initial_setup(self) -> ExperimentConfig:
        raise NotImplementedError()


ALL_EXPERIMENT_CLASSES: Dict[str, Callable[[], Experiment]] = {}


def register_experiment(name: str, cls):
    ALL_EXPERIMENT_CLASSES[name] = cls


def make_experiment(name: str) -> Experiment:
    return ALL_EXPERIMENT_CLASSES[name]()
