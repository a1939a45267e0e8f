"""Model worker: one process per GPU; holds model shards, datasets and the on-device data store.

Parity: `realhf/system/model_worker.py` (lazy setup :185-399, request handling :505-661, data transfer :781-814,
param realloc :422-463, offload :464-475, save / evaluate, recover).  Tensors never leave the worker: replies to
the master carry metadata only.  Runs with `nccl` on GPUs and with `gloo` on CPUs (the plumbing configuration).
"""

from __future__ import annotations

import os
import time
import traceback
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from realhf_b200.api import data as data_api
from realhf_b200.api import model as model_api
from realhf_b200.api import system as system_api
from realhf_b200.api.config import ModelInterfaceType, ModelName, ModelShardID
from realhf_b200.api.data import SequenceSample
from realhf_b200.base import constants, logging, monitor, name_resolve, seeding
from realhf_b200.base.topology import ParallelContext
from realhf_b200.parallel import realloc
from realhf_b200.system.stream import Payload, WorkerStream

logger = logging.getLogger("model_worker")


def _pg_key(exp, trial):
    return f"{exp}/{trial}/global_pg_addr"


class ModelWorker:
    def __init__(self, cfg: system_api.ModelWorker):
        self.cfg = cfg
        info = cfg.worker_info
        self.exp, self.trial, self.index, self.count = info.experiment_name, info.trial_name, info.worker_index, info.worker_count
        self.models: Dict[ModelName, model_api.Model] = {}
        self.ctxs: Dict[ModelName, ParallelContext] = {}
        self.backends: Dict[ModelName, model_api.ModelBackend] = {}
        self.interfaces: Dict[str, model_api.ModelInterface] = {}
        self.shard_ids: Dict[ModelName, ModelShardID] = {}
        self.data_storage: Dict[Any, SequenceSample] = {}
        self._n_calls: Dict[str, int] = {}
        self.dataloader = None
        self.data_iter = None
        self.epoch = 0
        self._realloc_cache: Dict[Tuple[ModelName, ModelName], realloc.ReallocExecutor] = {}
        self._host_copy_stale: set = set()  # models written by a realloc (e.g. reference EMA) since their last offload
        self._ipc_owned: Dict[ModelName, torch.Tensor] = {}
        self._ipc_peers: Dict[Tuple[ModelName, int], torch.Tensor] = {}
        self._exiting = False

    # ------------------------------------------------------------------ setup
    def setup(self):
        cfg = self.cfg
        seeding.set_random_seed(cfg.seed, offset=self.index)
        from realhf_b200.models import generation
        # the fused sampler's Philox seed: a function of the experiment seed, identical on every worker (TP ranks must agree)
        generation.seed_sampling(seeding.derive_seed("sampling"))
        if cfg.device == "cuda":
            torch.cuda.set_device(int(os.environ.get("REAL_LOCAL_GPU", self.index % max(torch.cuda.device_count(), 1))))
            self.device = torch.device("cuda", torch.cuda.current_device())
        else:
            self.device = torch.device("cpu")
        self.stream = WorkerStream(self.exp, self.trial, self.index)
        # global process group over all model workers
        if self.index == 0:
            from realhf_b200.system.stream import free_port, host_ip
            name_resolve.add(_pg_key(self.exp, self.trial), f"tcp://{host_ip()}:{free_port()}", replace=True)
        addr = name_resolve.wait(_pg_key(self.exp, self.trial), timeout=300)
        kw = dict(device_id=self.device) if cfg.device == "cuda" else {}
        dist.init_process_group(cfg.backend, init_method=addr, rank=self.index, world_size=self.count, **kw)
        # peer-memory paths (CUDA IPC / VMM handles: direct-store reallocation, fused TP kernels, NVLS optimizer) need every
        # participant on ONE host; across nodes the NCCL paths stay in place
        hosts = [None] * self.count
        dist.all_gather_object(hosts, os.uname().nodename)
        self.worker_hosts = hosts
        self.single_host = len(set(hosts)) == 1
        # every worker builds the groups of EVERY model in the same order (new_group is collective)
        worker_of: Dict[ModelName, List[int]] = {}
        for sid, wi in cfg.msid2mwid.items():
            worker_of.setdefault(sid.model_name, [None] * cfg.model_topos[sid.model_name].world_size())
            worker_of[sid.model_name][sid.parallelism_rank if sid.topo is not None else
                                      cfg.model_topos[sid.model_name].get_rank(pipe=sid.pp_rank, data=sid.dp_rank, model=sid.tp_rank)] = wi
        self.worker_of = worker_of
        for name in sorted(cfg.model_topos, key=str):
            topo = cfg.model_topos[name]
            ctx = ParallelContext.build(topo, worker_of[name], self.index, backend=cfg.backend,
                                        sequence_parallel=getattr(topo, "sequence_parallel", False),
                                        gradient_checkpointing=getattr(topo, "gradient_checkpointing", False))
            if ctx.is_member:
                self.ctxs[name] = ctx
        # fused tensor-parallel GEMM + collective kernels over peer memory (one node, GPUs mutually visible); construction is
        # collective over each TP group, so every member decides from the same config / env
        if cfg.device == "cuda" and self.single_host and os.environ.get("REAL_FUSED_TP", "1") != "0" and os.environ.get("REAL_ISOLATE_GPUS", "0") != "1":
            for name in sorted(self.ctxs, key=str):
                ctx = self.ctxs[name]
                if ctx.tp_size > 1 and getattr(ctx, "symm", None) is None:
                    try:
                        from realhf_b200.parallel.fused_tp import FusedTP
                        ctx.symm = FusedTP(ctx, max_tokens=int(os.environ.get("REAL_FUSED_TP_MAX_TOKENS", "32768")),
                                           max_features=int(os.environ.get("REAL_FUSED_TP_MAX_HIDDEN", "8192")), device=self.device)
                    except Exception as e:  # e.g. IPC not permitted in this container: plain GEMM + NCCL stays in place
                        logger.warning(f"fused TP kernels disabled for {name}: {e}")
        # models / backends / interfaces
        for shard in cfg.shards:
            name = shard.id.model_name
            self.shard_ids[name] = shard.id
            model_cfg = shard.model
            rdir = self._recover_dir(name)
            trainable = any(r.model_name == name and r.interface_type == ModelInterfaceType.TRAIN_STEP for r in cfg.model_rpcs)
            if rdir is None and trainable and shard.should_instantiate and os.environ.get("REAL_RECOVER_RUN", "0") == "1":
                # e.g. the rank that writes the weights was the one that died: the step counters will resume, the weights cannot
                logger.warning(f"recover run, but no saved weights for the trainable model {name}: starting it from {shard.model.args.get('model_path')}")
            if rdir is not None and shard.should_instantiate:
                # recover run: weights come from the states saved at the failure (reference: model_worker.py:308-313)
                import copy
                model_cfg = copy.deepcopy(shard.model)
                model_cfg.args.update(model_path=rdir, init_from_scratch=False, init_critic_from_actor=False)
                logger.info(f"recover run: loading {name} from {rdir}")
            with constants.model_scope(name, self.ctxs[name], instantiate=shard.should_instantiate):
                model = model_api.make_model(model_cfg, name=name, device=self.device)
            self.models[name] = model
            self.backends[name] = model_api.make_backend(shard.backend)
            self._eval_dataset_cfg = shard.eval_dataset
            self._eval_bs = int(getattr(shard, "eval_bs", 128) or 128)
        for rpc in cfg.model_rpcs:
            if rpc.model_name in self.models:
                self.interfaces[rpc.name] = model_api.make_interface(rpc.interface_impl)
                if os.environ.get("REAL_RECOVER_RUN", "0") == "1":
                    load_interface_state(rpc.name, self.interfaces[rpc.name], os.path.join(constants.RECOVER_ROOT, self.exp, self.trial, "ckpt"))
        # dataset (only on data-owner workers)
        if cfg.datasets:
            src = next(r for r in cfg.model_rpcs if r.is_src)
            ctx = self.ctxs[src.model_name]
            use_cache = cfg.use_dataset_cache or os.environ.get("REAL_DATASET_CACHE", "0") == "1"   # tokenised samples cached on disk
            cache_root = os.path.join(constants.run_dirs(self.exp, self.trial)["log"], "..", "..", "cache") if use_cache else None
            ds = [data_api.make_dataset(d, cfg.seed, ctx.dp_rank, ctx.dp_size, cfg.tokenizer_name_or_path, self.exp, self.trial, cache_root)
                  for d in cfg.datasets]
            dataset = ds[0] if len(ds) == 1 else torch.utils.data.ConcatDataset(ds)
            self.dataset = dataset
            self.dataset_size = len(dataset)
            g = torch.Generator()
            g.manual_seed(cfg.seed)
            self.dataloader = torch.utils.data.DataLoader(dataset, batch_size=max(1, src.n_seqs // ctx.dp_size), shuffle=True,
                                                          collate_fn=SequenceSample.gather, generator=g)
            self.data_iter = iter(self.dataloader)
        logger.info(f"model worker {self.index} ready: models {[str(n) for n in self.models]}")

    # ------------------------------------------------------------------ hooks
    def _data_transfer(self, plan: List[dict]):
        """plan entries: {key, ids, lens(list per id of list[int]), dtype, trailing, src, dsts} executed identically (same
        order) by every involved worker through one batch of isend/irecv."""
        ops, recvs, keep = [], [], []
        for e in plan:
            src, dsts = e["src"], e["dsts"]
            if self.index == src:
                items = [self.data_storage[i] for i in e["ids"]]
                t = torch.cat([it.data[e["key"]] for it in items], 0) if len(items) > 1 else items[0].data[e["key"]]
                t = t.contiguous()
                keep.append(t)
                for d in dsts:
                    if d != src:
                        ops.append(dist.P2POp(dist.isend, t, d))
            elif self.index in dsts:
                total = sum(sum(l) for l in e["lens"])
                buf = torch.empty((total, *e["trailing"]), dtype=e["dtype"], device=self.device)
                ops.append(dist.P2POp(dist.irecv, buf, src))
                recvs.append((e, buf))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for e, buf in recvs:
            off = 0
            for i, lens in zip(e["ids"], e["lens"]):
                n = sum(lens)
                piece = buf[off: off + n]
                off += n
                with SequenceSample.disable_validation():
                    s = SequenceSample(keys=[e["key"]], ids=[i], seqlens={e["key"]: [lens]}, trailing_shapes={e["key"]: e["trailing"]},
                                       dtypes={e["key"]: e["dtype"]}, data={e["key"]: piece})
                if i in self.data_storage:
                    self.data_storage[i].update_(s)
                else:
                    self.data_storage[i] = s

    _alias_logged: set = set()   # per process: (src, dst) pairs whose aliasing was reported

    def _param_realloc(self, spec: dict):
        """spec: {src: ModelName, dst: ModelName, eta}.  Both replicas' workers call this."""
        src_name, dst_name, eta = spec["src"], spec["dst"], spec.get("eta", 1.0)
        key = (src_name, dst_name)
        src_model, dst_model = self.models.get(src_name), self.models.get(dst_name)
        if spec.get("noop"):  # reverse direction of a plain realloc: the destination kept its weights, drop the copy
            if src_model is not None and spec.get("release_src"):
                _real(src_model).release_params()
            return
        mcfg = (src_model or dst_model).module_config
        if key not in self._realloc_cache:
            plan = realloc.derive_plan(mcfg, self.cfg.model_topos[src_name], self.worker_of[src_name],
                                       self.cfg.model_topos[dst_name], self.worker_of[dst_name], for_worker=self.index)
            es = torch.tensor([], dtype=(src_model or dst_model).dtype).element_size()
            self._realloc_cache[key] = realloc.ReallocExecutor(plan, self.index, es, self.device)
        ex = self._realloc_cache[key]
        # ZeRO-3 keeps only this rank's slice of the source between calls: gather the full buffer for the transfer (a collective
        # over the source's DP group -- every source worker runs this hook) and drop it again afterwards
        src_optim = getattr(src_model.module, "optim", None) if src_model is not None else None
        zero3_src = src_optim is not None and getattr(src_optim.cfg, "zero_stage", 1) >= 3 and not _real(src_model).instantiated
        if zero3_src:
            src_optim.materialize()
        src_flat = _real(src_model).flat_param.data if src_model is not None and _real(src_model).instantiated else None
        dst_flat = None
        # receive-only replicas (replica_id > 0, created by the allocation for another layout) live in IPC-shareable
        # memory on GPUs: senders store the destination layout straight into them over NVLink, one kernel per transfer
        direct = self._direct_realloc_ok(dst_name)
        # A receive-only replica whose shard on this GPU is exactly the source's shard (same tp / pp position: e.g. generation on
        # dp4 over half of the GPUs of a dp8 training layout) does not get a copy at all: it aliases the source's flat buffer.
        # Saves the copy and the replica's memory (13.5 GB per GPU for a 7B model); nothing is received from peers in that case.
        alias = (eta == 1.0 and not zero3_src and src_flat is not None and dst_model is not None and dst_name.replica_id > 0 and ex.whole_local_copy()
                 and _real(dst_model).flat_numel == src_flat.numel() and os.environ.get("REAL_REALLOC_ALIAS", "1") == "1")
        if alias:
            m = _real(dst_model)
            if not m.instantiated or m.flat_param.data_ptr() != src_flat.data_ptr():
                m.attach_flat(src_flat)
                for prm in m.parameters():
                    prm.requires_grad_(False)
                if key not in self._alias_logged:
                    self._alias_logged.add(key)
                    logger.info(f"realloc {src_name} -> {dst_name}: the replica's shard on this GPU is the source's shard; aliased, no copy")
            dst_flat = m.flat_param.data
        elif dst_model is not None:
            m = _real(dst_model)
            if not m.instantiated:
                if getattr(m, "_offloaded", None) is not None:
                    # the destination was parked in pinned host memory by an OffloadHook: bring its weights back before
                    # they are overwritten (eta == 1) or mixed into (EMA, eta != 1: a fresh zero buffer would turn
                    # ref <- eta*actor + (1-eta)*ref into eta*actor)
                    m.reload()
                    if self.device.type == "cuda":
                        torch.cuda.current_stream(self.device).synchronize()
                else:
                    self._alloc_recv_flat(dst_name, m)
            dst_flat = m.flat_param.data
        if dst_model is not None:
            self._host_copy_stale.add(dst_name)
        if direct:
            es = torch.tensor([], dtype=(src_model or dst_model).dtype).element_size()
            ptrs = {t.dst_worker: self._peer_flat_ptr(dst_name, t.dst_worker, ex.plan.dst_numel[t.dst_worker] * es) for t in ex.sends}
            ex.run(src_flat, dst_flat, eta=eta, peer_dst_ptrs=ptrs, notify=True, skip_local=alias)
        else:
            ex.run(src_flat, dst_flat, eta=eta, skip_local=alias)
        if zero3_src:
            if self.device.type == "cuda":
                torch.cuda.current_stream(self.device).synchronize()  # the transfer reads the gathered buffer asynchronously
            src_optim.release()
        # a non-trainable source replica is dropped after handing its weights back
        if spec.get("release_src") and src_model is not None:
            _real(src_model).release_params()

    # ------------------------------------------------------------------ direct (peer-store) reallocation plumbing
    def _direct_realloc_ok(self, dst_name: ModelName) -> bool:
        return (self.device.type == "cuda" and dst_name.replica_id > 0 and os.environ.get("REAL_REALLOC_DIRECT", "1") != "0"
                and os.environ.get("REAL_ISOLATE_GPUS", "0") != "1" and getattr(self, "single_host", True))

    def _ipc_key(self, name: ModelName, worker: int) -> str:
        return f"{self.exp}/{self.trial}/realloc_ipc/{name}/{worker}"

    def _alloc_recv_flat(self, name: ModelName, m):
        """Flat buffer of a replica that only ever receives weights by reallocation."""
        nbytes = m.flat_numel * torch.tensor([], dtype=m.dtype).element_size()
        if self._direct_realloc_ok(name):
            from realhf_b200.ops import lib
            buf = self._ipc_owned.get(name)
            if buf is None:  # allocated once: peers keep their mapping of it across release / re-attach cycles
                buf, handle = lib().symm_alloc(max(nbytes, 16), self.device.index)
                self._ipc_owned[name] = buf
                name_resolve.add(self._ipc_key(name, self.index), bytes(handle.tolist()).hex(), replace=True)
            flat = buf[:nbytes].view(m.dtype)
        else:
            flat = torch.zeros(m.flat_numel, dtype=m.dtype, device=self.device)
        m.attach_flat(flat)
        for p in m.parameters():
            p.requires_grad_(False)

    def _peer_flat_ptr(self, name: ModelName, worker: int, nbytes: int) -> int:
        """Device address (in this process) of `worker`'s flat buffer of replica `name`."""
        k = (name, worker)
        if k not in self._ipc_peers:
            if worker == self.index:
                self._ipc_peers[k] = self._ipc_owned[name]
            else:
                from realhf_b200.ops import lib
                hexs = name_resolve.wait(self._ipc_key(name, worker), timeout=300)
                handle = torch.tensor(list(bytes.fromhex(hexs)), dtype=torch.uint8)
                self._ipc_peers[k] = lib().symm_open(handle, max(nbytes, 16), self.device.index)
        return int(self._ipc_peers[k].data_ptr())

    # ------------------------------------------------------------------ request handling
    def _handle(self, req: Payload) -> Any:
        for h, d in zip(req.pre_hooks, req.pre_hook_data):
            self._run_hook(h, d)
        res = self._handle_core(req)
        for h, d in zip(req.post_hooks, req.post_hook_data):
            self._run_hook(h, d)
        return res

    def _run_hook(self, h: str, d: Any):
        if h == "data_transfer":
            with monitor.cuda_tmarked("data_transfer", monitor.CUDATimeMarkType.comm):
                self._data_transfer(d)
        elif h == "param_realloc":
            with monitor.cuda_tmarked("param_realloc", monitor.CUDATimeMarkType.mem_layout):
                self._param_realloc(d)
        elif h == "offload":
            m = self.models.get(d["model"])
            if m is not None:
                # weights that did not change since the last offload keep a valid pinned host copy: just free the device copy
                _real(m).offload(frozen=d["model"] not in self._host_copy_stale)
                self._host_copy_stale.discard(d["model"])
        else:
            raise NotImplementedError(h)

    # ------------------------------------------------------------------ dataset prefetch
    def _load_batch(self):
        """(next batch, whether the iterator wrapped into a new epoch to produce it).  Runs on the prefetch thread."""
        try:
            return next(self.data_iter), False
        except StopIteration:
            self.data_iter = iter(self.dataloader)
            return next(self.data_iter), True

    def _take_batch(self):
        """The batch for this `fetch`, and the start of the next one's load (tokenisation / collation on a side thread while
        the MFCs of this step run) -- the reference prefetches from the dataset at every poll (model_worker.py:401-416).
        The iterator is only ever touched by one thread at a time: the pending load is awaited before anything else."""
        import concurrent.futures
        if getattr(self, "_prefetch_pool", None) is None:
            self._prefetch_pool = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix="dataset-prefetch")
            self._prefetched = None
        fut, self._prefetched = self._prefetched, None
        out = fut.result() if fut is not None else self._load_batch()
        if os.environ.get("REAL_DATASET_PREFETCH", "1") == "1":
            self._prefetched = self._prefetch_pool.submit(self._load_batch)
        return out

    def _handle_core(self, req: Payload) -> Any:
        h = req.handle_name
        if h == "empty":
            return None
        if h == "spec":
            return dict(dataset_size=self.dataset_size, steps_per_epoch=len(self.dataloader))
        if h == "fetch":
            batch, wrapped = self._take_batch()
            final = False
            if wrapped:
                self.epoch += 1
            ignore = set(req.data.get("ignore_ids", [])) if isinstance(req.data, dict) else set()
            items = [x for x in batch.unpack() if x.ids[0] not in ignore]
            for it in items:
                self.data_storage[it.ids[0]] = it.to_device(self.device)
            meta = SequenceSample.gather(items).meta() if items else None
            return data_api.DataBatchMeta(dp_rank=self.ctxs[next(r for r in self.cfg.model_rpcs if r.is_src).model_name].dp_rank,
                                          meta_sample=meta, epoch=self.epoch, is_final_batch=final)
        if h == "clear_data_cache":
            for i in req.data:
                self.data_storage.pop(i, None)
            self._n_cache_clears = getattr(self, "_n_cache_clears", 0) + 1
            freq = self.cfg.cuda_cache_clear_freq
            if self.device.type == "cuda" and (self.cfg.cuda_cache_cleanliness or (freq and self._n_cache_clears % freq == 0)):
                # `cuda_cache_cleanliness`: after every step; otherwise every `cuda_cache_clear_freq` steps (reference: cache_clear_freq)
                torch.cuda.empty_cache()
                self._dirty_cache = False
            if monitor.TIME_MARK_DB:
                monitor.dump_tmark_db(os.path.join(constants.run_dirs(self.exp, self.trial)["log"], f"time_marks{self.index}.pkl"))
            return None
        name = req.model_name
        model = self.models[name]
        if h == "model_config":
            return model.module_config
        if h == "initialize":
            m = _real(model)
            if not m.instantiated:  # replica that only ever receives weights by realloc
                self._alloc_recv_flat(name, m)
            self.models[name] = self.backends[name].initialize(model, req.data)
            rdir = self._recover_dir(name)
            if rdir is not None and os.path.isdir(os.path.join(rdir, "optim")):
                self.backends[name].load(self.models[name], os.path.join(rdir, "optim"))  # optimizer moments + loss scale
            if rdir is not None:
                from realhf_b200.base import recover
                info = recover.load_recover_info(self.exp, self.trial)
                if info is not None:  # the version counter drives the LR schedule and the checkpoint names
                    v = self.models[name].version
                    v.epoch, v.epoch_step, v.global_step = info.recover_start.epoch, info.recover_start.epoch_step, info.recover_start.global_step
            return None
        if h == "save":
            rpc = next(r for r in self.cfg.model_rpcs if r.model_name == name)
            self.interfaces[rpc.name].save(model, req.data)
            self.backends[name].save(model, os.path.join(req.data, "optim"))
            return None
        if h == "evaluate":
            rpc = next(r for r in self.cfg.model_rpcs if r.model_name == name)
            if self._eval_dataset_cfg is None:
                return {}
            ctx = self.ctxs[name]
            ds = data_api.make_dataset(self._eval_dataset_cfg, self.cfg.seed, ctx.dp_rank, ctx.dp_size, self.cfg.tokenizer_name_or_path)
            # `dataset.valid_bs_n_seqs` sequences per evaluation batch over the whole DP group (reference: eval_bs of the shard)
            dl = data_api.make_dataloader("packed_eval", ds, batch_size=max(1, self._eval_bs // ctx.dp_size))
            return self.interfaces[rpc.name].evaluate(model, dl)
        if h in ("generate", "inference", "train_step"):
            self._maybe_inject_fault(h)
            rpc = next(r for r in self.cfg.model_rpcs if r.name == req.data["rpc_name"])
            ids = req.data["ids"]
            inp = SequenceSample.gather([self.data_storage[i] for i in ids], keys=rpc.input_keys)
            if rpc.input_key_remap:
                inp.remap_keys_(rpc.input_key_remap)
            if self.cfg.profile_mode:   # profiling experiments: the interface fabricates the keys this handle needs from the ids alone
                inp = self.interfaces[rpc.name].mock(h, model, inp)   # (reference: model_worker.py:740-741)
            t0 = time.perf_counter()
            if self.device.type == "cuda":
                if h == "generate" and self.cfg.cuda_cache_cleanliness and getattr(self, "_dirty_cache", False):
                    # With the master's look-ahead the generation of step s+1 can be queued before `clear_data_cache` of step s:
                    # give the KV cache (tens of GB in one piece) the blocks the training step left in the caching allocator
                    torch.cuda.empty_cache()
                    self._dirty_cache = False
                self._dirty_cache = getattr(self, "_dirty_cache", False) or h == "train_step"
                torch.cuda.reset_peak_memory_stats(self.device)
            self._n_calls[rpc.name] = call = self._n_calls.get(rpc.name, 0) + 1
            monitor.time_mark(f"{rpc.name}_start", f"model_worker/{self.index}", step=call - 1)  # REAL_TIME_MARK=1
            with self._mfc_profile(rpc.name), monitor.cuda_tmarked(rpc.name, _TMARK_OF[h], str(name)):
                res = getattr(self.interfaces[rpc.name], h)(model, inp, n_mbs=rpc.n_mbs)
            if self.device.type == "cuda":
                torch.cuda.synchronize()
            monitor.time_mark(f"{rpc.name}_end", f"model_worker/{self.index}", step=call - 1)
            dt = time.perf_counter() - t0
            ctx = self.ctxs[name]
            mem = self._memory_stats()
            if isinstance(res, SequenceSample):
                if rpc.output_key_remap:
                    res.remap_keys_(rpc.output_key_remap)
                if ctx.is_dp_head:
                    for it in res.unpack():
                        if it.ids[0] in self.data_storage:
                            self.data_storage[it.ids[0]].update_(it)
                        else:
                            self.data_storage[it.ids[0]] = it
                return dict(meta=res.meta() if ctx.is_dp_head else None, secs=dt, mem=mem)
            return dict(stats=res if ctx.is_dp_head else None, secs=dt, mem=mem)
        raise NotImplementedError(f"unknown request `{h}`")

    # ------------------------------------------------------------------ observability
    def _mfc_profile(self, rpc_name: str):
        """`REAL_DUMP_TRACE=1`: one torch.profiler chrome trace per MFC call under <log>/trace/; `REAL_DUMP_MEMORY=1`: allocator
        history snapshot per call (reference: model_worker.py:65-78, :663-721)."""
        import contextlib
        trace = os.environ.get("REAL_DUMP_TRACE", "0") == "1"
        memory = os.environ.get("REAL_DUMP_MEMORY", "0") == "1" and self.device.type == "cuda"
        if not trace and not memory:
            return contextlib.nullcontext()
        worker = self

        @contextlib.contextmanager
        def cm():
            out_dir = os.path.join(constants.run_dirs(worker.exp, worker.trial)["log"], "trace")
            os.makedirs(out_dir, exist_ok=True)
            worker._n_profiled = getattr(worker, "_n_profiled", 0) + 1
            tag = f"{rpc_name}_r{worker.index}_c{worker._n_profiled}"
            if memory:
                torch.cuda.memory._record_memory_history(max_entries=100000)
            prof = None
            if trace:
                acts = [torch.profiler.ProfilerActivity.CPU] + ([torch.profiler.ProfilerActivity.CUDA] if worker.device.type == "cuda" else [])
                prof = torch.profiler.profile(activities=acts, record_shapes=True, profile_memory=True, with_stack=False)
                prof.__enter__()
            try:
                yield
            finally:
                if prof is not None:
                    prof.__exit__(None, None, None)
                    prof.export_chrome_trace(os.path.join(out_dir, tag + ".json"))
                if memory:
                    torch.cuda.memory._dump_snapshot(os.path.join(out_dir, tag + ".mem.pkl"))
                    torch.cuda.memory._record_memory_history(enabled=None)
        return cm()

    def _memory_stats(self) -> Dict[str, float]:
        """Peak allocator numbers of the MFC that just ran; travels with the reply (the reference all-gathers these over the
        model group after every MFC, model_worker.py:1031-1040 -- here they ride on the control plane instead)."""
        if self.device.type != "cuda":
            return {}
        st = dict(peak_allocated_gb=torch.cuda.max_memory_allocated(self.device) / 2 ** 30,
                  peak_reserved_gb=torch.cuda.max_memory_reserved(self.device) / 2 ** 30, worker=self.index)
        try:
            free, total = torch.cuda.mem_get_info(self.device)
            st["device_used_gb"] = (total - free) / 2 ** 30
        except RuntimeError:  # driver query unavailable (e.g. inside some containers): the allocator numbers are enough
            pass
        return st

    # ------------------------------------------------------------------ main loop
    def run(self):
        self.setup()
        from realhf_b200.system.worker_control import WorkerServer, WorkerServerStatus
        # request / response control endpoint (ping / status / progress / exit), served on its own thread so that it answers
        # while an MFC is running
        self.ctl = WorkerServer(self.exp, self.trial, f"model_worker/{self.index}")
        self.ctl.set_status(WorkerServerStatus.RUNNING)
        self.ctl.register_handler("progress", lambda: dict(current=getattr(self, "_current_handle", None),
                                                           n_handled=getattr(self, "_n_handled", 0), memory=self._memory_stats()))
        while not self._exiting:
            if self.ctl.exit_requested.is_set():
                logger.info("exit requested through the control panel")
                break
            req = self.stream.poll(timeout_ms=50)
            if req is None:
                continue
            self._current_handle = req.handle_name
            self._n_handled = getattr(self, "_n_handled", 0) + 1
            if req.handle_name == "exit":
                self.stream.reply(req, None)
                break
            try:
                res = self._handle(req)
                self.stream.reply(req, res)
            except Exception as e:  # report and die: the controller / scheduler handles recovery
                err = f"{type(e).__name__}: {e}\n{traceback.format_exc()}"
                logger.error(err)
                self.stream.reply(req, None, error=err)
                break
        self.exit()

    def exit(self):
        if os.environ.get("REAL_SAVE_RECOVER_STATES", "0") == "1":
            self.save_recover_states()
        try:
            dist.destroy_process_group()
        except Exception as e:  # peers may already be gone at teardown
            logger.debug(f"destroy_process_group: {e!r}")
        self.stream.close()

    def _maybe_inject_fault(self, handle: str):
        """Fault injection for recovery tests (the reference has none): `REAL_FAULT_INJECT=<worker>:<handle>:<n>` makes that
        worker raise on its n-th call of `handle` -- only in the original run, never in a recover run."""
        spec = os.environ.get("REAL_FAULT_INJECT")
        if not spec or os.environ.get("REAL_RECOVER_RUN", "0") == "1":
            return
        w, hname, n = spec.split(":")
        if int(w) != self.index or hname != handle:
            return
        self._fault_calls = getattr(self, "_fault_calls", 0) + 1
        if self._fault_calls == int(n):
            raise RuntimeError(f"injected fault: worker {self.index}, call {n} of {handle}")

    def _recover_dir(self, name: ModelName) -> Optional[str]:
        if os.environ.get("REAL_RECOVER_RUN", "0") != "1":
            return None
        d = os.path.join(constants.RECOVER_ROOT, self.exp, self.trial, "ckpt", name.role)
        return d if os.path.exists(os.path.join(d, "config.json")) else None

    def save_recover_states(self):
        root = os.path.join(constants.RECOVER_ROOT, self.exp, self.trial, "ckpt")
        for name, model in self.models.items():
            if _real(model).instantiated:
                rpc = next((r for r in self.cfg.model_rpcs if r.model_name == name), None)
                if rpc is not None:
                    self.interfaces[rpc.name].save(model, os.path.join(root, name.role))
                    self.backends[name].save(model, os.path.join(root, name.role, "optim"))
        save_interface_states(self.interfaces, root, tag=str(self.index))


def save_interface_states(interfaces: Dict, root: str, tag: str = "0"):
    """KL controller / value-normaliser state of every interface (identical on every rank of a model: the statistics are
    all-reduced), one file per MFC, written atomically so concurrent ranks cannot leave a torn file."""
    for rpc_name, itf in interfaces.items():
        sd = itf.state_dict()
        if sd:
            os.makedirs(root, exist_ok=True)
            tmp = os.path.join(root, f".interface_{rpc_name}.{tag}.tmp")
            torch.save(sd, tmp)
            os.replace(tmp, os.path.join(root, f"interface_{rpc_name}.pt"))


def load_interface_state(rpc_name: str, itf, root: str) -> bool:
    f = os.path.join(root, f"interface_{rpc_name}.pt")
    if not os.path.exists(f):
        return False
    itf.load_state_dict(torch.load(f, weights_only=False))
    logger.info(f"recover run: restored the interface state of {rpc_name}")
    return True


_TMARK_OF = {"generate": monitor.CUDATimeMarkType.forward, "inference": monitor.CUDATimeMarkType.forward,
             "train_step": monitor.CUDATimeMarkType.backward}


def _real(model: model_api.Model):
    m = model.module
    return getattr(m, "module", m)
