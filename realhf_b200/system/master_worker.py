"""Master worker: walks the dataflow graph, holds only metadata, dispatches MFCs to model workers.

Parity: `realhf/system/master_worker.py` (lazy init :927-1271, request/reply coroutines :455-680, data loading
:683-781, save / eval / benchmark control :1307-1400, e2e logging :1407-1488, recover info :1541-1554).
One long-lived asyncio coroutine per MFC; an MFC fires as soon as `n_seqs` samples carry all of its input keys, so independent
MFCs (rew_inf / ref_inf / critic_inf of PPO) overlap when their meshes are disjoint -- also ACROSS steps: like the reference's
free-running request coroutines (master_worker.py:455-680: one semaphore per MFC plus "never run ahead of the train step of your
own role"), the generation of step s+1 starts as soon as `actor_train` of step s has replied, while `critic_train` of step s is
still running on its own GPUs.  `exp_ctrl.max_inflight_steps` bounds the look-ahead (1 = a barrier after every step).
"""

from __future__ import annotations

import asyncio
import collections
import os
import time
from typing import Dict, Hashable, List, Optional, Tuple

import numpy as np

from realhf_b200.base import asyncio_utils
from realhf_b200.api import system as system_api
from realhf_b200.api.config import ModelInterfaceType, ModelName
from realhf_b200.api.data import DataBatchMeta, SequenceSample
from realhf_b200.api.dfg import MFCDef, OffloadHook, ParamReallocHook
from realhf_b200.api.model import FinetuneSpec
from realhf_b200.base import constants, logging, monitor, recover, timeutil
from realhf_b200.system.buffer import AsyncIOSequenceBuffer
from realhf_b200.system.stream import MasterStream, Payload

logger = logging.getLogger("master_worker", "system")


class MasterWorker:
    def __init__(self, cfg: system_api.MasterWorker):
        self.cfg = cfg
        info = cfg.worker_info
        self.exp, self.trial = info.experiment_name, info.trial_name
        self.rpcs = cfg.model_rpcs
        self.topos = cfg.model_topos
        self.msid2mwid = cfg.msid2mwid
        self.workers_of: Dict[ModelName, List[int]] = {}
        for sid, w in self.msid2mwid.items():
            topo = self.topos[sid.model_name]
            self.workers_of.setdefault(sid.model_name, [None] * topo.world_size())
            self.workers_of[sid.model_name][topo.get_rank(pipe=sid.pp_rank, data=sid.dp_rank, model=sid.tp_rank)] = w
        self.src_rpc = next(r for r in self.rpcs if r.is_src)
        self.sink_rpcs = [r for r in self.rpcs if r.is_dst]
        self.data_owner: Dict[Tuple[Hashable, str], int] = {}
        self._pending: Dict[str, asyncio.Future] = {}
        self._ready_batch: List[Tuple[List[Payload], List[Payload], asyncio.Future]] = []
        self.step = self.epoch = self.epoch_step = 0
        self.rpc_secs: Dict[str, float] = collections.defaultdict(float)
        self.rpc_mem: Dict[str, dict] = {}
        self._rpc_batch_lens: Dict[str, List[int]] = {}
        self._rpc_batch_lens_by_step: Dict[int, Dict[str, List[int]]] = {}
        self._t_start = time.time()
        self.stats_log: List[Dict] = []
        self._metrics = None
        self._ids_by_step: Dict[int, List[Hashable]] = {}     # ids the source MFC consumed in each (global) step
        self._recovered_ignore: List[Hashable] = []
        self._epoch_first_step = 0
        self._rpc_secs_by_step: Dict[int, Dict[str, float]] = collections.defaultdict(lambda: collections.defaultdict(float))

    # ------------------------------------------------------------------ transport helpers
    async def _request(self, p: Payload) -> Payload:
        fut = asyncio.get_running_loop().create_future()
        self._pending[p.request_id] = fut
        self.stream.post(p)
        return await fut

    async def _dispatch(self, phase1: List[Payload], phase2: List[Payload]) -> List[Payload]:
        """Post the requests of one MFC.  MFCs that become ready within `REAL_MASTER_BATCH_MS` (2 ms) of each other -- in PPO
        `actor_train` and `critic_train` always do: the same reply makes both runnable -- are posted together: first the phase-1
        (data redistribution) requests of all of them, then their phase-2 (compute) requests.  Each needs tensors that live on the
        OTHER one's workers; posted one MFC after the other, the second one's source workers would be busy computing the first
        one before they get to send (model workers are serial), and two MFCs on disjoint GPUs would run one after the other.
        The reference reaches the same end on the worker side by running the pre-hooks of all queued requests first
        (model_worker.py:483-503), which needs its three-phase handshake; here the order stays a single global one."""
        loop = asyncio.get_running_loop()
        fut = loop.create_future()
        self._ready_batch.append((phase1, phase2, fut))
        if len(self._ready_batch) == 1:
            loop.call_later(float(os.environ.get("REAL_MASTER_BATCH_MS", "2")) / 1e3, self._flush_batch)
        return await fut

    def _flush_batch(self):
        batch, self._ready_batch = self._ready_batch, []
        if getattr(self, "_stop_launch", False):   # the walk is being torn down: nothing new goes to the workers
            for entry in batch:
                entry[2].cancel()
            return
        loop = asyncio.get_running_loop()
        futs: Dict[int, List[asyncio.Future]] = {i: [] for i in range(len(batch))}
        try:
            for phase in (0, 1):
                for i, entry in enumerate(batch):
                    for p in entry[phase]:
                        f = loop.create_future()
                        self._pending[p.request_id] = f
                        self.stream.post(p)
                        futs[i].append(f)
        except Exception as e:   # a timer callback has no caller to raise to: hand the failure to the MFC coroutines, which end the walk
            for entry in batch:
                if not entry[2].done():
                    entry[2].set_exception(RuntimeError(f"could not post the requests of an MFC: {e!r}"))
            return
        for i, entry in enumerate(batch):
            asyncio.ensure_future(self._collect(futs[i], entry[2]))

    @staticmethod
    async def _collect(futs: List[asyncio.Future], out: asyncio.Future):
        """Replies of one MFC (phase-1 requests first, then phase-2, in posting order) -> the future its coroutine awaits."""
        try:
            res = await asyncio.gather(*futs)
        except asyncio.CancelledError:
            out.cancel()
            raise
        except Exception as e:
            if not out.done():
                out.set_exception(e)
            return
        if not out.done():
            out.set_result(res)

    async def _pump(self):
        """Routes replies to the futures waiting for them."""
        while True:
            r = self.stream.poll(0)
            if r is None:
                await asyncio.sleep(0.001)
                continue
            fut = self._pending.pop(r.request_id, None)
            if fut is not None and not fut.done():
                if r.error:
                    fut.set_exception(RuntimeError(f"model worker {r.handler} failed in `{r.handle_name}`:\n{r.error}"))
                else:
                    fut.set_result(r)

    async def _group_request(self, workers: List[int], handle: str, data=None, model_name=None, **kw) -> List[Payload]:
        return await asyncio.gather(*[self._request(Payload(handler=w, handle_name=handle, data=data, model_name=model_name, **kw))
                                      for w in workers])

    # ------------------------------------------------------------------ init
    async def _lazy_init(self):
        cfg = self.cfg
        n = cfg.n_model_workers
        self.stream = MasterStream(self.exp, self.trial, n)
        self.stream.wait_workers()
        self._pump_task = asyncio.create_task(self._pump())
        # dataset size from the data owners (dp heads of the source MFC)
        src_workers = self._dp_heads(self.src_rpc.model_name)
        specs = await self._group_request(src_workers, "spec")
        self.dataset_size = sum(s.data["dataset_size"] for s in specs)
        if self.dataset_size < self.src_rpc.n_seqs:
            raise ValueError(f"the dataset has {self.dataset_size} usable samples but `{self.src_rpc.name}` consumes {self.src_rpc.n_seqs} "
                             f"sequences per step: lower the batch size (dataset.train_bs_n_seqs) or check the length filter "
                             f"(max_seqlen / max_prompt_len drops longer records)")
        steps_per_epoch = max(1, self.dataset_size // self.src_rpc.n_seqs)
        total_steps = steps_per_epoch * cfg.exp_ctrl.total_train_epochs
        self.ft_spec = FinetuneSpec(cfg.exp_ctrl.total_train_epochs, total_steps, steps_per_epoch)
        # backends: trainable replicas first so that weight-receiving replicas exist when the first realloc happens
        names = sorted(self.topos, key=lambda nm: (nm.role, nm.replica_id))
        for nm in names:
            await self._group_request(self.workers_of[nm], "initialize", data=self.ft_spec, model_name=nm)
        self.buffer = AsyncIOSequenceBuffer(self.rpcs)
        # model shapes for the analytic FLOP accounting of the step log (reference: master_worker.py:1407-1488)
        self.model_cfgs = {}
        for nm in names:
            try:
                r = await self._group_request(self.workers_of[nm][:1], "model_config", model_name=nm)
                self.model_cfgs[nm] = r[0].data
            except Exception as e:  # only the TFLOP/s column of the step log depends on it
                logger.warning(f"no model config for {nm} ({e}): FLOP accounting disabled for it")
        ec = cfg.exp_ctrl
        self.save_ctl = timeutil.EpochStepTimeFreqCtl(ec.save_freq_epochs, ec.save_freq_steps, ec.save_freq_secs)
        self.eval_ctl = timeutil.EpochStepTimeFreqCtl(ec.eval_freq_epochs, ec.eval_freq_steps, ec.eval_freq_secs)
        self.recover_info = recover.load_recover_info(self.exp, self.trial) if os.environ.get("REAL_RECOVER_RUN", "0") == "1" else None
        if self.recover_info is not None:
            self.step, self.epoch, self.epoch_step = (self.recover_info.recover_start.global_step,
                                                      self.recover_info.recover_start.epoch, self.recover_info.recover_start.epoch_step)
        logger.info(f"master ready: dataset {self.dataset_size} samples, {steps_per_epoch} steps/epoch, {total_steps} steps in total")

    def _dp_heads(self, name: ModelName) -> List[int]:
        topo = self.topos[name]
        pp, dp, tp = topo.dims
        return [self.workers_of[name][topo.get_rank(pipe=pp - 1, data=d, model=0)] for d in range(dp)]

    # ------------------------------------------------------------------ data loading
    def _epoch_of(self, step: int) -> int:
        """Epoch index of global step `step` (steps that are in flight ahead of `self.step` may belong to the next epoch)."""
        return self.epoch + (self.epoch_step + (step - self.step)) // self.ft_spec.steps_per_epoch

    async def _load_data(self, step: Optional[int] = None):
        src_workers = self._dp_heads(self.src_rpc.model_name)
        epoch = self.epoch if step is None else self._epoch_of(step)
        ignore = list(self.recover_info.hash_vals_to_ignore) if (self.recover_info and epoch == self.recover_info.recover_start.epoch) else []
        replies = await self._group_request(src_workers, "fetch", data=dict(ignore_ids=ignore))
        samples = []
        for w, r in zip(src_workers, replies):
            meta: DataBatchMeta = r.data
            if meta.meta_sample is None:
                continue
            for s in meta.meta_sample.unpack():
                if s.ids[0] in self.buffer:
                    # the last batch of an epoch was smaller than n_seqs, its samples are still waiting, and the next epoch
                    # brought the same prompt again: one live instance per id (the worker-side tensors are identical)
                    continue
                for k in s.keys:
                    self.data_owner[(s.ids[0], k)] = w
                samples.append(s)
        await self.buffer.put_batch(samples)

    # ------------------------------------------------------------------ one MFC
    def _transfer_plan(self, rpc: MFCDef, part: Dict[int, List[Hashable]], meta: SequenceSample) -> List[dict]:
        topo = self.topos[rpc.model_name]
        pp, dp, tp = topo.dims
        by_id = {i: s for i, s in zip(meta.ids, meta.unpack())}
        plan = []
        for d, ids in part.items():
            dsts = sorted({self.workers_of[rpc.model_name][topo.get_rank(pipe=p, data=d, model=t)] for p in range(pp) for t in range(tp)})
            for key in rpc.input_keys:
                groups: Dict[int, List[Hashable]] = collections.defaultdict(list)
                for i in ids:
                    groups[self.data_owner[(i, key)]].append(i)
                for src, gids in groups.items():
                    need = [w for w in dsts if w != src]
                    if not need:
                        continue
                    plan.append(dict(key=key, ids=gids, lens=[by_id[i].seqlens[key][0] for i in gids], dtype=meta.dtypes[key],
                                     trailing=tuple(meta.trailing_shapes[key] or ()), src=src, dsts=need))
        return plan

    async def _run_rpc_once(self, rpc: MFCDef, step: Optional[int] = None):
        step = self.step if step is None else step
        if rpc.is_src:
            # usually one fetch; after a recover run whole batches may be filtered out (ids consumed before the failure)
            attempts = 0
            while self.buffer.n_ready_for(rpc) < rpc.n_seqs:
                await self._load_data(step)
                attempts += 1
                if attempts > 2 * self.ft_spec.steps_per_epoch + 2:
                    raise RuntimeError(f"dataset cannot supply {rpc.n_seqs} fresh sequences "
                                       f"({self.buffer.n_ready_for(rpc)} ready after {attempts} fetches)")
        ids, meta = await self.buffer.get_batch_for_rpc(rpc)
        topo = self.topos[rpc.model_name]
        pp, dp, tp = topo.dims
        # partition over dp ranks: token-balanced contiguous split (or equal counts with balanced_dp)
        if rpc.balanced_dp:
            per = len(ids) // dp
            assert per * dp == len(ids), f"balanced_dp needs n_seqs % dp == 0 ({len(ids)} % {dp})"
            parts = [(k * per, (k + 1) * per) for k in range(dp)]
        else:
            need = max(1, (rpc.n_mbs or 1) * (2 * pp if pp > 1 else 1))
            if rpc.interface_type == ModelInterfaceType.TRAIN_STEP:
                # every DP rank runs the interface's minibatch loop (one optimizer step with the group's gradient collectives per
                # minibatch): it needs at least that many sequences, or the ranks would take different numbers of steps
                need = max(need, int((getattr(rpc.interface_impl, "args", None) or {}).get("n_minibatches", 1) or 1))
            parts = meta.get_split_spec(dp, min_size=need).partitions
        part = {d: ids[a:b] for d, (a, b) in enumerate(parts)}
        self._rpc_batch_lens_by_step.setdefault(step, {})[rpc.name] = self._batch_lens(meta)
        plan = self._transfer_plan(rpc, part, meta)
        involved = set(self.workers_of[rpc.model_name])
        for e in plan:
            involved.add(e["src"])
        pre_hooks, pre_data, post_hooks, post_data = [], [], [], []
        for h in rpc._pre_hooks:
            if isinstance(h, ParamReallocHook):
                src, dst = (h.source, rpc.model_name) if h.source is not None else (rpc.model_name, h.target)
                pre_hooks.append("param_realloc")
                pre_data.append(dict(src=src, dst=dst, eta=h.eta))
                involved |= set(self.workers_of[src]) | set(self.workers_of[dst])
        for h in rpc._post_hooks:
            if isinstance(h, ParamReallocHook):
                src, dst = (h.source, rpc.model_name) if h.source is not None else (rpc.model_name, h.target)
                trainable_dst = any(r.model_name == dst and r.interface_type == ModelInterfaceType.TRAIN_STEP for r in self.rpcs)
                if trainable_dst and h.eta == 1.0:
                    # the trainable replica never lost its weights: the reverse direction only drops the copy
                    post_hooks.append("param_realloc")
                    post_data.append(dict(src=src, dst=dst, eta=1.0, release_src=True, noop=True))
                else:
                    post_hooks.append("param_realloc")
                    post_data.append(dict(src=src, dst=dst, eta=h.eta))
                involved |= set(self.workers_of[src]) | set(self.workers_of[dst])
            elif isinstance(h, OffloadHook):
                post_hooks.append("offload")
                post_data.append(dict(model=rpc.model_name))
        # Phase 1: the redistribution of the input data, as hook-only requests to the workers that send or receive something.
        # Phase 2: the call itself (with the reallocation / offload hooks).  MFCs that become ready together are dispatched as
        # "all their phase-1 requests, then all their phase-2 requests" (`_dispatch`), still ONE global order on FIFO channels.
        movers = sorted({e["src"] for e in plan} | {d for e in plan for d in e["dsts"]})
        phase1 = [Payload(handler=w, handle_name="empty", model_name=rpc.model_name, pre_hooks=["data_transfer"], pre_hook_data=[plan])
                  for w in movers]
        phase2 = []
        member = {}
        for r in range(topo.world_size()):
            member[self.workers_of[rpc.model_name][r]] = topo.get_coord(r)
        hook_workers = set(member)
        for hd in pre_data + [d for h, d in zip(post_hooks, post_data) if h == "param_realloc"]:
            hook_workers |= set(self.workers_of[hd["src"]]) | set(self.workers_of[hd["dst"]])
        for w in sorted(hook_workers):
            if w in member:
                c = member[w]
                p = Payload(handler=w, handle_name=rpc.interface_type.value, model_name=rpc.model_name,
                            data=dict(rpc_name=rpc.name, ids=part[c.data]), pre_hooks=pre_hooks, pre_hook_data=pre_data,
                            post_hooks=post_hooks, post_hook_data=post_data)
            else:
                p = Payload(handler=w, handle_name="empty", model_name=rpc.model_name, pre_hooks=pre_hooks, pre_hook_data=pre_data,
                            post_hooks=[h for h in post_hooks if h == "param_realloc"],
                            post_hook_data=[d for h, d in zip(post_hooks, post_data) if h == "param_realloc"])
            phase2.append(p)
        t0 = time.perf_counter()
        replies = (await self._dispatch(phase1, phase2))[len(phase1):]
        self._rpc_secs_by_step[step][rpc.name] += time.perf_counter() - t0
        if step == self.step:
            self.rpc_secs[rpc.name] = self._rpc_secs_by_step[step][rpc.name]   # live view of the step being finished (control panel)
        heads = set(self._dp_heads(rpc.model_name))
        stats = []
        mems = [r.data["mem"] for r in replies if isinstance(r.data, dict) and r.data.get("mem")]
        if mems:  # max over the workers of this MFC, with the worker that holds it (reference: model_worker.py:999-1094)
            top = max(mems, key=lambda m: m.get("peak_allocated_gb", 0.0))
            self.rpc_mem[rpc.name] = top
        for r in replies:
            if r.handler not in heads or not isinstance(r.data, dict):
                continue
            if r.data.get("meta") is not None:
                m: SequenceSample = r.data["meta"]
                missing = set(rpc.output_keys) - set(m.keys)
                if missing:
                    # the consumers of these keys would wait forever: a hang with idle GPUs is the worst way to learn about it
                    raise RuntimeError(f"MFC `{rpc.name}` declares the output keys {sorted(rpc.output_keys)} but its interface returned "
                                       f"{sorted(m.keys)}: missing {sorted(missing)}")
                items = m.unpack()
                for it in items:
                    for k in it.keys:
                        self.data_owner[(it.ids[0], k)] = r.handler
                await self.buffer.amend_batch([it.ids[0] for it in items], items)
            if r.data.get("stats") is not None:
                stats.append(r.data["stats"])
        if stats and rpc.log_return_value:
            merged = {k: float(np.mean([s[k] for s in stats if k in s])) for k in stats[0] if isinstance(stats[0][k], (int, float))}
            logger.info(f"[{rpc.name}] step {step}: " + ", ".join(f"{k}={v:.4g}" for k, v in merged.items()))
            rec = {"rpc": rpc.name, "step": step, "epoch": self._epoch_of(step), "time": time.time(), **merged}
            self.stats_log.append(rec)
            self._write_stats(rec)
        if rpc.is_src:
            self._ids_by_step.setdefault(step, []).extend(ids)
        return ids

    # ------------------------------------------------------------------ main loop
    async def _mfc_loop(self, rpc: MFCDef, first: int, last: int):
        """All traversals [first, last) of one MFC.  Traversal s may start when
          * its own traversal s-1 has replied (this loop is sequential),
          * the train step of its own role has finished traversal s-1 (nobody computes with weights older than one update;
            reference: master_worker.py:502-509),
          * for a train step: step s-1 is finalised (its checkpoint / evaluation is not racing with the next update),
          * s is inside the look-ahead window: s < finalised steps + `max_inflight_steps`,
          * the trial is not paused,
        and, as always, when the buffer holds `n_seqs` samples with all of its input keys."""
        is_train = rpc.interface_type == ModelInterfaceType.TRAIN_STEP
        train = next((r for r in self.rpcs if r.role == rpc.role and r.interface_type == ModelInterfaceType.TRAIN_STEP), None)
        W = self._window

        def may_start(s: int) -> bool:
            if self._stop_launch:
                return True
            if self._hold or s >= self._finalized + W:
                return False
            if train is not None and train is not rpc and self._done[train.name] < s:
                return False
            return not is_train or self._finalized >= s

        for s in range(first, last):
            async with self._cv:
                await self._cv.wait_for(lambda: may_start(s))
                if self._stop_launch:
                    return
            await self._run_rpc_once(rpc, s)
            async with self._cv:
                self._done[rpc.name] = s + 1
                self._cv.notify_all()

    async def _finish_step(self, t0: float) -> float:
        """Book-keeping after every MFC has finished traversal `self.step`."""
        s = self.step
        done = self.buffer.pop_fully_consumed()
        if done:
            gone = set(done)
            for k in [k for k in self.data_owner if k[0] in gone]:
                del self.data_owner[k]
            await self._group_request(list(range(self.cfg.n_model_workers)), "clear_data_cache", data=done)
        secs = self._rpc_secs_by_step.pop(s, {})
        self._rpc_batch_lens = self._rpc_batch_lens_by_step.pop(s, {})
        self.rpc_secs.clear()
        self.rpc_secs.update(secs)
        self.step += 1
        self.epoch_step += 1
        if self.epoch_step >= self.ft_spec.steps_per_epoch:
            self.epoch += 1
            self.epoch_step = 0
            self._epoch_first_step = self.step
            self._recovered_ignore = []
            for k in [k for k in self._ids_by_step if k < self.step]:
                del self._ids_by_step[k]
        dt = time.perf_counter() - t0
        logger.info(f"step {self.step} (epoch {self.epoch}, {self.epoch_step}/{self.ft_spec.steps_per_epoch}) e2e {dt:.3f}s; "
                    + ", ".join(f"{k} {v:.2f}s" for k, v in secs.items()))
        self._log_throughput(dt)
        if self.rpc_mem:
            logger.info("peak memory: " + ", ".join(
                f"{k} {m.get('peak_allocated_gb', 0):.1f}/{m.get('peak_reserved_gb', 0):.1f} GB alloc/reserved @worker{m.get('worker')}"
                for k, m in self.rpc_mem.items()))
        self.rpc_secs.clear()
        self.rpc_secs.update(self._rpc_secs_by_step.get(self.step, {}))
        return dt

    def _ckpt_tag(self) -> str:
        """Name of a checkpoint written after the step that just finished: epoch index of that step, its 1-based position in
        the epoch, global step count (the reference's convention, master_worker.py:1330-1343 -- the last step of an epoch is
        `epoch{e}epochstep{steps_per_epoch}`, not `epoch{e+1}epochstep0`)."""
        if self.epoch_step == 0 and self.step > 0:
            return f"epoch{self.epoch - 1}epochstep{self.ft_spec.steps_per_epoch}globalstep{self.step}"
        return f"epoch{self.epoch}epochstep{self.epoch_step}globalstep{self.step}"

    async def _save(self):
        for nm in sorted(self.topos, key=str):
            if any(r.model_name == nm and r.interface_type == ModelInterfaceType.TRAIN_STEP for r in self.rpcs):
                d = os.path.join(constants.run_dirs(self.exp, self.trial)["save"], nm.role, self._ckpt_tag())
                await self._group_request(self.workers_of[nm], "save", data=d, model_name=nm)
                logger.info(f"saved {nm} to {d}")

    async def _eval(self):
        for nm in sorted(self.topos, key=str):
            if any(r.model_name == nm and r.interface_type == ModelInterfaceType.TRAIN_STEP for r in self.rpcs):
                res = await self._group_request(self.workers_of[nm], "evaluate", model_name=nm)
                # the statistics live on the last pipeline stage (already reduced over DP): report a DP head's reply
                heads = set(self._dp_heads(nm))
                stats = [r.data for r in res if r.handler in heads and r.data]
                logger.info(f"eval {nm}: {stats[:1]}")
                if stats:
                    self._write_stats({"rpc": f"eval/{nm.role}", "step": self.step, "epoch": self.epoch, "time": time.time(),
                                       **{k: v for k, v in stats[0].items() if isinstance(v, (int, float))}})

    async def _main(self):
        await self._lazy_init()
        ec = self.cfg.exp_ctrl
        times = []
        total = self.ft_spec.total_train_steps
        last = total if ec.benchmark_steps is None else min(total, max(self.step, ec.benchmark_steps))
        self._window = max(1, int(os.environ.get("REAL_MASTER_INFLIGHT_STEPS", getattr(ec, "max_inflight_steps", 2) or 1)))
        self._cv = asyncio.Condition()
        self._done = {r.name: self.step for r in self.rpcs}     # traversals every MFC has finished (global step numbering)
        self._stop_launch = self._hold = False
        self._finalized = self.step                               # steps whose MFCs AND post-step work (save / eval) are done
        self._epoch_first_step = self.step - self.epoch_step
        if self.recover_info is not None:   # a second failure in the same epoch must still skip what the first run consumed
            self._recovered_ignore = list(self.recover_info.hash_vals_to_ignore)
        loops = [asyncio.ensure_future(self._mfc_loop(r, self.step, last)) for r in self.rpcs]

        async def set_flags(**kw):
            async with self._cv:
                for k, v in kw.items():
                    setattr(self, k, v)
                self._cv.notify_all()

        try:
            while self.step < last:
                # controller commands act between steps; while paused no new MFC is launched (MFCs of the next step that are already
                # running finish on the workers)
                await set_flags(_hold=True)
                go_on = await self._check_control()
                await set_flags(_hold=False)
                if not go_on:
                    logger.info(f"stop requested by the controller at step {self.step}")
                    break
                t0 = time.perf_counter()
                s = self.step

                async def step_done():
                    async with self._cv:
                        await self._cv.wait_for(lambda: all(self._done[r.name] > s for r in self.rpcs))

                # the first failure of any MFC loop cancels the others (they would wait for its outputs forever)
                waiter = asyncio.ensure_future(step_done())
                try:
                    while not waiter.done():
                        await asyncio.wait([waiter] + [t for t in loops if not t.done()], return_when=asyncio.FIRST_COMPLETED)
                        asyncio_utils.raise_first_exception(loops)
                finally:
                    if not waiter.done():
                        waiter.cancel()
                times.append(await self._finish_step(t0))
                if self.save_ctl.check(epochs=int(self.epoch_step == 0), steps=1):
                    await self._save()
                if self.eval_ctl.check(epochs=int(self.epoch_step == 0), steps=1):
                    await self._eval()
                await set_flags(_finalized=self.step)   # wake the loops waiting for the window / the finalised step
                if ec.benchmark_steps is not None and self.step >= ec.benchmark_steps:
                    logger.info(f"benchmark finished: avg #e2e# time {np.mean(times):.3f}s over {len(times)} steps")
                    break
        finally:
            self._stop_launch = True
            await asyncio_utils.cancel_all(loops)
            if os.environ.get("REAL_SAVE_RECOVER_STATES", "0") == "1":
                self._dump_recover()
            try:
                await asyncio.wait_for(self._group_request(list(range(self.cfg.n_model_workers)), "exit"), timeout=30)
            except Exception as e:  # workers that already died cannot acknowledge; the launcher stops them anyway
                logger.warning(f"not every model worker acknowledged `exit`: {e!r}")
            self._pump_task.cancel()
            self.stream.close()
            if self._metrics is not None:
                self._metrics.close()  # flushes TensorBoard / finishes the wandb run
        return times

    @staticmethod
    def _batch_lens(meta: SequenceSample) -> List[int]:
        for k in ("packed_input_ids", "packed_prompts"):
            if k in meta.keys:
                return meta.flat_seqlens(k)
        return []

    def _log_throughput(self, step_secs: float):
        """Tokens per batch, analytic TFLOP/s (x3 for a training MFC, x4 with activation recomputation) and ETA."""
        try:
            flops, n_tokens = 0.0, 0
            for rpc in self.rpcs:
                lens = self._rpc_batch_lens.get(rpc.name) or []
                mc = self.model_cfgs.get(rpc.model_name)
                if not lens or mc is None:
                    continue
                args = (len(lens), lens, mc.n_layers, mc.hidden_dim, mc.intermediate_dim, mc.vocab_size)
                if rpc.interface_type == ModelInterfaceType.TRAIN_STEP:
                    flops += monitor.calculate_llama_train_flops(4 if getattr(self.topos[rpc.model_name], "gradient_checkpointing", False) else 3, *args)
                    n_tokens = max(n_tokens, sum(lens))
                elif rpc.interface_type == ModelInterfaceType.GENERATE:
                    g = (rpc.interface_impl.args or {}).get("generation_config", {}) or {}
                    flops += monitor.calculate_llama_gen_flops(len(lens), lens, int(g.get("max_new_tokens", 256)), *args[2:])
                else:
                    flops += monitor.calculate_llama_forward_flops(*args)
                    n_tokens = max(n_tokens, sum(lens))
            done = max(self.step, 1)
            eta = (self.ft_spec.total_train_steps - self.step) * (time.time() - self._t_start) / done
            n_gpus = max(1, self.cfg.n_model_workers)
            logger.info(f"throughput: {n_tokens} tokens in the batch, {flops / step_secs / 1e12:.2f} TFLOP/s total, "
                        f"{flops / step_secs / 1e12 / n_gpus:.2f} per worker; ETA {eta / 60:.1f} min")
        except Exception as e:  # accounting must never take the run down
            logger.debug(f"throughput accounting skipped: {e}")

    async def _check_control(self) -> bool:
        """Controller commands between steps: `pause` (publish PAUSED, wait for `resume`), `exit` (stop gracefully)."""
        from realhf_b200.apps.remote import control_key, status_key, status_ttl
        from realhf_b200.base import name_resolve
        ckey, skey = control_key(self.exp, self.trial, "master_worker", 0), status_key(self.exp, self.trial, "master_worker", 0)

        def cmd():
            try:
                return name_resolve.get(ckey)
            except name_resolve.NameEntryNotFoundError:
                return None
        ctl = self._control_server()
        paused = lambda c_: c_ == "pause" or ctl.paused.is_set()  # the control key, or a `pause` request on the RPC panel
        c = cmd()
        if paused(c) and not ctl.exit_requested.is_set():
            from realhf_b200.system.worker_control import WorkerServerStatus
            name_resolve.add(skey, "PAUSED", replace=True, keepalive_ttl=status_ttl())
            ctl.set_status(WorkerServerStatus.PAUSED)
            logger.info(f"paused by the controller at step {self.step}")
            while paused(c) and not ctl.exit_requested.is_set():
                await asyncio.sleep(0.2)
                c = cmd()
            name_resolve.add(skey, "RUNNING", replace=True, keepalive_ttl=status_ttl())
            ctl.set_status(WorkerServerStatus.RUNNING)
            logger.info("resumed")
        return c != "exit" and not ctl.exit_requested.is_set()

    def _control_server(self):
        """Request / response control endpoint of this worker (`system/worker_control.py`), created on first use."""
        ctl = getattr(self, "_ctl", None)
        if ctl is None:
            from realhf_b200.system.worker_control import WorkerServer, WorkerServerStatus
            ctl = self._ctl = WorkerServer(self.exp, self.trial, "master_worker/0")
            ctl.set_status(WorkerServerStatus.RUNNING)
            ctl.register_handler("progress", lambda: dict(step=self.step, epoch=self.epoch, epoch_step=self.epoch_step,
                                                          total_steps=self.ft_spec.total_train_steps if self.ft_spec else None,
                                                          rpc_secs=dict(self.rpc_secs)))
        return ctl

    def _write_stats(self, rec: Dict):
        """Per-step statistics -> stats.jsonl (+ TensorBoard / wandb when enabled), see `system/metrics.py`."""
        if self._metrics is None:
            from realhf_b200.system.metrics import MetricSinks
            self._metrics = MetricSinks(self.exp, self.trial, constants.run_dirs(self.exp, self.trial)["log"])
        self._metrics.log(rec)

    def _consumed_ids_this_epoch(self) -> List[Hashable]:
        """Ids consumed by the FINALISED steps of the current epoch (plus what a previous, recovered run of this epoch consumed).
        Prompts that a look-ahead generation of an unfinished step already took are not in the list: nothing was trained on them."""
        out = list(self._recovered_ignore)
        for st in range(self._epoch_first_step, self.step):
            out += self._ids_by_step.get(st, [])
        return out

    def _dump_recover(self):
        info = recover.RecoverInfo(recover_start=recover.StepInfo(self.epoch, self.epoch_step, self.step),
                                   last_step_info=recover.StepInfo(self.epoch, max(self.epoch_step - 1, 0), max(self.step - 1, 0)),
                                   hash_vals_to_ignore=self._consumed_ids_this_epoch())
        recover.dump_recover_info(info, self.exp, self.trial)

    def run(self):
        return asyncio.run(self._main())
