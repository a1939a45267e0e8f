"""CommonExperimentConfig: from user-level dataclass options to per-worker system configs.

Parity: `realhf/experiments/common/common.py` (fields :158-179, allocation modes :319-402, worker configs
:404-487) and `experiments/common/utils.py` (replica ids :126-140, realloc / offload hooks :154-198).
Allocation modes: `manual`, `heuristic`, `pipe_data`, `pipe_model`, `search`, and the `d{a}m{b}p{c}` pattern.
The heuristic is re-derived for 180 GB B200s (see `heuristic_allocation`): models that fit replicate
data-parallel on every GPU instead of being cut by TP/PP.
"""

from __future__ import annotations

import dataclasses
import os
import re
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from realhf_b200.api.config import (DataLoaderAbstraction, DatasetAbstraction, ModelAbstraction, ModelBackendAbstraction,
                                    ModelInterfaceType, ModelName, ModelShardID, StandaloneModelShardAbstraction)
from realhf_b200.api.dfg import MFCDef, OffloadHook, ParamReallocHook
from realhf_b200.api.quickstart import (DeviceMesh, MFCConfig, ModelTrainEvalConfig, ParallelismConfig, RPCAllocation,
                                        find_parallel_strategies, make_device_mesh_from_name, parallelism_eq)
from realhf_b200.api.system import (Experiment, ExperimentConfig, ExperimentSaveEvalControl, ExperimentScheduling, ModelWorker,
                                    Scheduling, TasksGroup)
from realhf_b200.base.topology import PipeModelDataParallelTopology


@dataclasses.dataclass
class CommonExperimentConfig(Experiment):
    experiment_name: str = "default-exp"
    trial_name: str = "default-trial"
    mode: str = "local"  # local | slurm | ray
    debug: bool = True
    partition: Optional[str] = None  # Slurm partition; None: the cluster spec's, else "dev"
    wandb_mode: str = "disabled"  # disabled | online | offline (exported as WANDB_MODE to the master worker)
    tensorboard: bool = False     # TensorBoard event files under <log dir>/tensorboard
    image_name: Optional[str] = None
    recover_mode: str = "disabled"  # disabled | auto | save | resume
    recover_retries: int = 1
    ignore_worker_error: bool = False
    allocation_mode: str = "heuristic"
    allocation_use_cache: bool = False
    n_nodes: int = 1
    n_gpus_per_node: int = 8
    nodelist: Optional[str] = None
    seed: int = 1
    cache_clear_freq: Optional[int] = 10
    exp_ctrl: ExperimentSaveEvalControl = dataclasses.field(default_factory=ExperimentSaveEvalControl)
    device: str = "cuda"  # "cpu" runs the whole runtime on gloo (plumbing / CI configuration)
    dtype: str = "bf16"

    # ---- what concrete experiments provide
    @property
    def models(self) -> Dict[str, ModelTrainEvalConfig]:
        raise NotImplementedError()

    @property
    def rpcs(self) -> Dict[str, MFCDef]:
        raise NotImplementedError()

    @property
    def allocations(self) -> Dict[str, MFCConfig]:
        raise NotImplementedError()

    @property
    def datasets(self) -> List[DatasetAbstraction]:
        raise NotImplementedError()

    @property
    def eval_datasets(self) -> Optional[List[DatasetAbstraction]]:
        return None

    @property
    def eval_dataloader(self) -> DataLoaderAbstraction:
        return DataLoaderAbstraction("packed_eval", args=dict(batch_size=128))

    @property
    def tokenizer_name_or_path(self) -> str:
        raise NotImplementedError()

    @property
    def max_prompt_len(self) -> Optional[int]:
        return None

    @property
    def search_kwargs(self) -> Dict[str, Any]:
        return {}

    # ---- meshes
    @property
    def n_workers(self) -> int:
        return self.n_nodes * self.n_gpus_per_node

    @property
    def global_device_mesh(self) -> DeviceMesh:
        return DeviceMesh(self.n_nodes, self.n_gpus_per_node, np.ones((self.n_nodes, self.n_gpus_per_node), dtype=np.int32),
                          self.nodelist, self.nodelist)

    def scheduling_setup(self) -> ExperimentScheduling:
        return ExperimentScheduling(
            model_worker=TasksGroup(self.n_workers, Scheduling.model_worker_default(gpu=0 if self.device == "cpu" else 1,
                                                                                    nodelist=self.nodelist, container_image=self.image_name)),
            master_worker=TasksGroup(1, Scheduling.master_worker_default(container_image=self.image_name)))

    # ---- allocation
    def _heuristic_rpc_allocation(self) -> List[RPCAllocation]:
        return heuristic_allocation(self)

    def _search(self) -> List[RPCAllocation]:
        """`allocation_mode=search`.  With `allocation_use_cache` the result is stored under the profiler cache, keyed by everything
        the search depends on (cluster shape, MFCs and batch sizes, model families / sizes / paths, sequence lengths), and reused
        by later launches of the same problem (reference: `allocation_use_cache`, experiments/common/common.py:330-350)."""
        from realhf_b200.search.engine import search_rpc_allocations
        kw = dict(seq_len=self.max_prompt_len or 1024, cross_step_overlap=int(getattr(self.exp_ctrl, "max_inflight_steps", 2) or 1) > 1,
                  **self.search_kwargs)
        cache = None
        if self.allocation_use_cache:
            import hashlib
            import pickle
            from realhf_b200.base import constants
            key = repr((self.n_nodes, self.n_gpus_per_node, sorted((n, r.n_seqs, r.interface_type.value, r.role) for n, r in self.rpcs.items()),
                        sorted((role, m.type._class, m.type.size, m.type.is_critic, m.path) for role, m in self.models.items()),
                        sorted((k, repr(v)) for k, v in kw.items())))
            cache = os.path.join(constants.PROFILER_CACHE_PATH, "allocations", hashlib.sha1(key.encode()).hexdigest()[:16] + ".pkl")
            if os.path.exists(cache):
                with open(cache, "rb") as f:
                    saved = pickle.load(f)
                by_name = self.rpcs
                return [RPCAllocation(by_name[name], mesh, par) for name, mesh, par in saved]
        allocs = search_rpc_allocations(self.global_device_mesh, list(self.rpcs.values()), self.models, **kw)
        if cache is not None:
            os.makedirs(os.path.dirname(cache), exist_ok=True)
            with open(cache, "wb") as f:
                pickle.dump([(a.rpc.name, a.device_mesh, a.parallel) for a in allocs], f)
        return allocs

    def _get_rpc_allocations(self) -> List[RPCAllocation]:
        rpcs, mesh = self.rpcs, self.global_device_mesh
        mode = self.allocation_mode
        if mode == "manual":
            out = []
            for name, rpc in rpcs.items():
                a = self.allocations[name]
                dm = make_device_mesh_from_name(self.nodelist, a.device_mesh, self.n_nodes, self.n_gpus_per_node) if a.device_mesh else mesh
                if a.parallel.world_size != dm.n_gpus:
                    raise ValueError(f"MFC {name}: parallelism {a.parallel} does not fill device mesh {dm} ({dm.n_gpus} GPUs)")
                out.append(RPCAllocation(rpc, dm, a.parallel))
            return out
        if mode == "heuristic":
            return self._heuristic_rpc_allocation()
        if mode == "search":
            return self._search()
        n = mesh.n_gpus
        if mode == "pipe_data":
            pp = self.n_nodes
            par = ParallelismConfig(1, pp, n // pp)
        elif mode == "pipe_model":
            pp = self.n_nodes
            par = ParallelismConfig(min(self.n_gpus_per_node, n // pp), pp, n // pp // min(self.n_gpus_per_node, n // pp))
        else:
            m = re.match(r"^d(\d+)m(\d+)p(\d+)$", mode) or re.match(r"^d(\d+)p(\d+)m(\d+)$", mode)
            if not m:
                raise ValueError(f"unknown allocation mode `{mode}`")
            if mode.index("m") < mode.index("p"):
                dp, tp, pp = (int(x) for x in m.groups())
            else:
                dp, pp, tp = (int(x) for x in m.groups())
            par = ParallelismConfig(tp, pp, dp)
            if par.world_size != n:
                raise ValueError(f"allocation {mode} needs {par.world_size} GPUs, the experiment has {n}")
        out = []
        for name, rpc in rpcs.items():
            p = dataclasses.replace(par)
            # sequence parallel only helps training (and is unsupported by generation)
            p.use_sequence_parallel = (rpc.interface_type == ModelInterfaceType.TRAIN_STEP and p.model_parallel_size > 1)
            out.append(RPCAllocation(rpc, mesh, p))
        return out

    # ---- system config
    def initial_setup(self) -> ExperimentConfig:
        for role, m in self.models.items():
            if getattr(m, "lora", None) is not None or any(getattr(self, f, False) for f in ("is_sft_lora", "is_rew_lora")):
                # the option names exist for command-line compatibility; silently training full weights instead would be worse
                raise NotImplementedError(f"LoRA is not supported (model `{role}`; the reference rejects it too, ppo_exp.py:199-202)")
        rpc_allocs = self._get_rpc_allocations()
        for a in rpc_allocs:
            if isinstance(a.rpc, str):
                a.rpc = self.rpcs[a.rpc]
            cfgd = self.allocations.get(a.rpc.name) if self._has_allocations() else None
            if cfgd is not None and cfgd.n_mbs is not None and a.rpc.n_mbs is None:
                a.rpc.n_mbs = cfgd.n_mbs
        from realhf_b200.api.dfg import build_graph
        build_graph([a.rpc for a in rpc_allocs])
        resolve_replica_ids(rpc_allocs)
        resolve_rpc_hooks(rpc_allocs, self.models)
        self._check(rpc_allocs)
        workers = self._get_model_worker_configs(rpc_allocs)
        cfg = ExperimentConfig(exp_ctrl=self.exp_ctrl, model_rpcs=[a.rpc for a in rpc_allocs], model_worker=workers)
        cfg.set_worker_information(self.experiment_name, self.trial_name)
        self._rpc_allocs = rpc_allocs
        return cfg

    def _has_allocations(self) -> bool:
        try:
            self.allocations
            return True
        except NotImplementedError:
            return False

    def _check(self, rpc_allocs: List[RPCAllocation]):
        for a in rpc_allocs:
            role_cfg = self.models[a.rpc.role]
            if a.parallel.pipeline_parallel_size > 1 and a.rpc.interface_type == ModelInterfaceType.GENERATE and a.parallel.use_sequence_parallel:
                raise ValueError("sequence parallelism is not supported for generation")
            if a.rpc.n_seqs < a.parallel.data_parallel_size * a.parallel.pipeline_parallel_size:
                raise ValueError(f"MFC {a.rpc.name}: batch of {a.rpc.n_seqs} sequences is too small for dp x pp = "
                                 f"{a.parallel.data_parallel_size * a.parallel.pipeline_parallel_size}")
            if a.rpc.balanced_dp and a.rpc.n_seqs % a.parallel.data_parallel_size:
                # found at launch, not by the master in the middle of the first step
                raise ValueError(f"MFC {a.rpc.name} splits its batch into equal shares per data-parallel rank (balanced_dp): "
                                 f"{a.rpc.n_seqs} sequences are not divisible by dp = {a.parallel.data_parallel_size}; "
                                 f"change the batch size (dataset.train_bs_n_seqs) or the data-parallel degree of this MFC")

    def _get_model_worker_configs(self, rpc_allocs: List[RPCAllocation]) -> List[ModelWorker]:
        src_rpc = next(a for a in rpc_allocs if a.rpc.is_src)
        workers: List[ModelWorker] = []
        handled = set()
        shards_of: Dict[int, List[StandaloneModelShardAbstraction]] = {i: [] for i in range(self.n_workers)}
        data_owner_workers = set()
        for a in rpc_allocs:
            if a.rpc.model_name in handled:
                continue
            handled.add(a.rpc.model_name)
            mcfg = self.models[a.rpc.role]
            # MFCs with the same mesh and (tp, pp, dp) share one replica whatever their SP flag (SP does not change where weights
            # live): the replica runs token-sharded when ANY of them asks for it -- typically the training call; generation
            # switches SP off for its own duration (models/generation.py), inference works either way
            use_sp = any(b.parallel.use_sequence_parallel for b in rpc_allocs if b.rpc.model_name == a.rpc.model_name)
            topo = PipeModelDataParallelTopology(a.parallel.pipeline_parallel_size, a.parallel.model_parallel_size,
                                                 a.parallel.data_parallel_size, sequence_parallel=use_sp and a.parallel.model_parallel_size > 1,
                                                 gradient_checkpointing=mcfg.gradient_checkpointing, max_prompt_len=self.max_prompt_len)
            trainable = any(b.rpc.role == a.rpc.role and b.rpc.interface_type == ModelInterfaceType.TRAIN_STEP for b in rpc_allocs)
            model = ModelAbstraction("real_model", args=dict(
                model_path=mcfg.path, is_critic=mcfg.type.is_critic, init_from_scratch=mcfg.init_from_scratch,
                init_critic_from_actor=mcfg.init_critic_from_actor, dtype=self.dtype, hf_model_family=mcfg.type._class,
                expert_parallel=getattr(mcfg, "expert_parallel", False)))
            if trainable:
                backend = ModelBackendAbstraction("train", args=dict(optimizer=dataclasses.asdict(mcfg.optimizer),
                                                                      zero_stage=mcfg.zero_stage, offload_optimizer=mcfg.offload))
            else:
                backend = ModelBackendAbstraction("inference")
            ranks = a.device_mesh.global_ranks()
            for r in range(topo.world_size()):
                sid = ModelShardID.from_parallelism_rank(a.rpc.model_name, topo, r)
                shards_of[ranks[r]].append(StandaloneModelShardAbstraction(
                    id=sid, model=model, backend=backend,
                    eval_dataset=(self.eval_datasets[0] if self.eval_datasets else None),
                    eval_bs=int(getattr(getattr(self, "dataset", None), "valid_bs_n_seqs", 128) or 128)))
                if a is src_rpc and sid.tp_rank == 0 and sid.pp_rank == topo.get_dim("pipe") - 1:
                    data_owner_workers.add(ranks[r])
        for i in range(self.n_workers):
            workers.append(ModelWorker(
                seed=self.seed, shards=shards_of[i], tokenizer_name_or_path=self.tokenizer_name_or_path,
                datasets=self.datasets if i in data_owner_workers else None, cuda_cache_clear_freq=self.cache_clear_freq or 10,
                backend="gloo" if self.device == "cpu" else "nccl", device=self.device))
        return workers


# ------------------------------------------------------------------------------------------- replica / hook resolution


def resolve_replica_ids(rpc_allocs: List[RPCAllocation]):
    """MFCs of one role with the same (mesh, layout) share a replica; the trainable layout gets replica 0."""
    by_role: Dict[str, List[RPCAllocation]] = {}
    for a in rpc_allocs:
        by_role.setdefault(a.rpc.role, []).append(a)
    for role, allocs in by_role.items():
        allocs = sorted(allocs, key=lambda a: a.rpc.interface_type != ModelInterfaceType.TRAIN_STEP)
        seen: List[Tuple[DeviceMesh, ParallelismConfig, int]] = []
        for a in allocs:
            for mesh, par, rid in seen:
                if mesh == a.device_mesh and parallelism_eq(par, a.parallel):
                    a.rpc.model_name = ModelName(role, rid)
                    break
            else:
                rid = len(seen)
                seen.append((a.device_mesh, a.parallel, rid))
                a.rpc.model_name = ModelName(role, rid)


def resolve_rpc_hooks(rpc_allocs: List[RPCAllocation], model_configs: Dict[str, ModelTrainEvalConfig]):
    role_cnt: Dict[str, int] = {}
    for a in rpc_allocs:
        role_cnt[a.rpc.role] = role_cnt.get(a.rpc.role, 0) + 1
    for a in rpc_allocs:
        rpc = a.rpc
        if rpc.interface_type != ModelInterfaceType.TRAIN_STEP:
            continue
        for b in rpc_allocs:
            other = b.rpc
            if other.name == rpc.name or other.role != rpc.role or other.model_name == rpc.model_name:
                continue
            # weights go train-layout -> other layout before the call, and the replica is dropped after it
            other.add_pre_hook(ParamReallocHook(source=rpc.model_name))
            other.add_post_hook(ParamReallocHook(target=rpc.model_name))
    # non-trainable roles that share GPUs with others park their weights in host memory between calls
    for a in rpc_allocs:
        rpc = a.rpc
        trainable = any(b.rpc.role == rpc.role and b.rpc.interface_type == ModelInterfaceType.TRAIN_STEP for b in rpc_allocs)
        if trainable or not model_configs[rpc.role].offload:
            continue
        overlapped = any(b.rpc.role != rpc.role and b.device_mesh.overlap(a.device_mesh) for b in rpc_allocs)
        if overlapped and rpc.is_dst_of_model_role:
            rpc.add_post_hook(OffloadHook())


# ------------------------------------------------------------------------------------------- heuristic for B200


def model_bytes(cfg_or_family, n_params_b: Optional[float] = None) -> float:
    return (n_params_b if n_params_b is not None else float(cfg_or_family.size)) * 1e9 * 2


def heuristic_allocation(exp: CommonExperimentConfig) -> List[RPCAllocation]:
    """B200 heuristic.  Budget per GPU = 180 GB.  Let W = sum of bf16 weights of all roles and S = optimizer +
    gradient state of trainable roles under ZeRO-1 over all N GPUs.  If W + S/N (+ headroom for KV cache and
    activations) fits, every MFC runs data-parallel on the whole mesh (no TP/PP collectives at all, no parameter
    reallocation needed).  Otherwise roles are cut by the smallest TP (then PP across nodes) that makes them fit,
    generation keeps TP small and uses the remaining GPUs as DP, and training gets TP with sequence parallelism —
    the structure of the reference heuristic (ppo_exp.py:419-613) with the 80 GB constant replaced."""
    mesh = exp.global_device_mesh
    N, cap = mesh.n_gpus, 0.9 * 180e9
    models = exp.models
    rpcs = exp.rpcs
    trainable = {r.role for r in rpcs.values() if r.interface_type == ModelInterfaceType.TRAIN_STEP}
    W = sum(model_bytes(m.type) for m in models.values())
    S = sum(model_bytes(models[r].type) * (1 + 6) for r in trainable)  # bf16 grads + fp32 master/m/v
    headroom = 30e9
    out = []
    if W + S / N + headroom <= cap:
        for name, rpc in rpcs.items():
            out.append(RPCAllocation(rpc, mesh, ParallelismConfig(1, 1, N)))
        return out
    for name, rpc in rpcs.items():
        mb = model_bytes(models[rpc.role].type)
        is_train = rpc.interface_type == ModelInterfaceType.TRAIN_STEP
        need = mb * (8 if is_train else 1) / (N if is_train else 1) + mb + headroom
        tp = 1
        while tp < min(8, mesh.n_gpus_per_node) and (mb / tp + (mb * 7 / N if is_train else 0) + headroom) > cap:
            tp *= 2
        pp = 1
        while (mb / tp / pp + headroom) > cap and pp < exp.n_nodes:
            pp *= 2
        dp = max(1, N // tp // pp)
        out.append(RPCAllocation(rpc, mesh, ParallelismConfig(tp, pp, dp, use_sequence_parallel=is_train and tp > 1)))
    return out
