"""The master worker's dataflow walk against simulated model workers: which MFC instances overlap, which never do.

A `FakeStream` stands in for the ZMQ transport: every model worker is a serial FIFO that needs a configured time per MFC and answers
with the metadata a real worker would return.  The REAL `MasterWorker` code (buffer, transfer plans, hooks, MFC loops, step
finalisation) runs on top of it, for a PPO experiment whose actor lives on GPUs 0-1 and whose critic / reward live on GPUs 2-3.
Checked: with `max_inflight_steps=2` the generation of step s+1 runs while `critic_train` of step s is still busy (the overlap the
allocation search's simulator counts on, reference master_worker.py:455-680), never before `actor_train` of step s has replied; a train
step never starts before the previous step is finalised; `max_inflight_steps=1` is a barrier; a failing MFC stops the trial."""
import asyncio
import heapq
import time

import pytest
import torch

from realhf_b200.api.data import DataBatchMeta, SequenceSample
from realhf_b200.api.model import FinetuneSpec
from realhf_b200.base import timeutil
from realhf_b200.system.buffer import AsyncIOSequenceBuffer
from realhf_b200.system.master_worker import MasterWorker
from realhf_b200.system.stream import Payload

BS = 8
DUR = dict(actor_gen=0.10, rew_inf=0.02, ref_inf=0.02, critic_inf=0.02, actor_train=0.06, critic_train=0.30)
OUT_KEYS = dict(actor_gen=("seq_no_eos_mask", "packed_input_ids", "packed_logprobs", "prompt_mask", "packed_logits_mask"),
                rew_inf=("rewards",), ref_inf=("packed_ref_logprobs",), critic_inf=("values",))


def _meta(ids, keys, seqlen=4):
    keys = set(keys)
    return SequenceSample(keys=keys, ids=list(ids), seqlens={k: [[seqlen] for _ in ids] for k in keys},
                          trailing_shapes={k: () for k in keys}, dtypes={k: torch.float32 for k in keys},
                          data={k: torch.zeros(len(ids) * seqlen) for k in keys}, metadata={}).meta()


class FakeStream:
    def __init__(self, n_workers, fail_at=None):
        self.busy = [0.0] * n_workers
        self.heap = []
        self.events = []       # (rpc_name, traversal, worker, start, end)
        self.count = {}
        self.next_id = 0
        self.fail_at = fail_at
        self.seq = 0

    def post(self, p: Payload):
        now = time.perf_counter()
        w = p.handler
        name = p.data.get("rpc_name") if isinstance(p.data, dict) and "rpc_name" in p.data else None
        dur = DUR.get(name, 0.0) if p.handle_name in ("generate", "inference", "train_step") else 0.0
        start = max(now, self.busy[w])
        end = start + dur
        self.busy[w] = end
        data, err = None, None
        if p.handle_name == "fetch":
            ids = list(range(self.next_id, self.next_id + BS))
            self.next_id += BS
            data = DataBatchMeta(dp_rank=0, meta_sample=_meta(ids, ("packed_prompts",)), epoch=0, is_final_batch=False)
        elif name is not None:
            k = self.count.get((name, w), 0)
            self.count[(name, w)] = k + 1
            self.events.append((name, k, w, start, end))
            if self.fail_at == (name, k):
                err = "injected failure"
            elif p.handle_name == "train_step":
                data = dict(stats=dict(loss=1.0), secs=dur, mem=None)
            else:
                data = dict(meta=_meta(p.data["ids"], OUT_KEYS[name]), secs=dur, mem=None)
        self.seq += 1
        heapq.heappush(self.heap, (end, self.seq, Payload(handler=w, handle_name=p.handle_name, request_id=p.request_id, data=data,
                                                         model_name=p.model_name, is_reply=True, error=err)))
        return p.request_id

    def poll(self, timeout_ms=0):
        if self.heap and self.heap[0][0] <= time.perf_counter():
            return heapq.heappop(self.heap)[2]
        return None

    def close(self):
        pass


class SimMaster(MasterWorker):
    def __init__(self, cfg, n_steps, fail_at=None):
        super().__init__(cfg)
        self._n_steps, self._fail_at = n_steps, fail_at

    async def _lazy_init(self):
        self.stream = FakeStream(self.cfg.n_model_workers, self._fail_at)
        self._pump_task = asyncio.create_task(self._pump())
        self.dataset_size = BS * self._n_steps
        self.ft_spec = FinetuneSpec(1, self._n_steps, self._n_steps)
        self.buffer = AsyncIOSequenceBuffer(self.rpcs)
        self.model_cfgs = {}
        self.save_ctl = timeutil.EpochStepTimeFreqCtl(None, None, None)
        self.eval_ctl = timeutil.EpochStepTimeFreqCtl(None, None, None)
        self.recover_info = None

    async def _check_control(self) -> bool:
        return True


def _master_cfg(window):
    from realhf_b200.experiments.algos import PPOConfig
    cfg = PPOConfig(experiment_name="sched", trial_name="t", n_nodes=1, n_gpus_per_node=4, allocation_mode="manual")
    cfg.dataset.train_bs_n_seqs = BS
    for name in ("actor_gen", "actor_train", "ref_inf"):
        a = getattr(cfg, name)
        a.device_mesh, a.parallel.data_parallel_size = "NODE01:0,1", 2
    for name in ("critic_inf", "critic_train", "rew_inf"):
        a = getattr(cfg, name)
        a.device_mesh, a.parallel.data_parallel_size = "NODE01:2,3", 2
    cfg.exp_ctrl.max_inflight_steps = window
    return cfg.initial_setup().master_worker[0]


def _run(window, n_steps=4, fail_at=None):
    m = SimMaster(_master_cfg(window), n_steps, fail_at)
    t0 = time.perf_counter()
    times = m.run()
    return m, times, time.perf_counter() - t0


def _span(events, name, k):
    ev = [e for e in events if e[0] == name and e[1] == k]
    assert ev, (name, k)
    return min(e[3] for e in ev), max(e[4] for e in ev)


def test_generation_of_the_next_step_overlaps_the_critic_update_of_this_one():
    m, times, wall = _run(window=2)
    ev = m.stream.events
    assert len(times) == 4 and m.step == 4
    for k in range(3):
        gen_next, critic_train = _span(ev, "actor_gen", k + 1), _span(ev, "critic_train", k)
        actor_train = _span(ev, "actor_train", k)
        assert gen_next[0] >= actor_train[1] - 1e-3          # nobody generates with weights the update of step k has not produced yet
        assert gen_next[0] < critic_train[1] - 0.05          # ... but it does not wait for the critic's update on the other GPUs
        assert _span(ev, "critic_inf", k + 1)[0] >= critic_train[1] - 1e-3   # the critic's own users do wait for it
        # a train step never starts before the previous step was finalised (all of its MFCs replied)
        prev_end = max(_span(ev, n, k)[1] for n in DUR)
        assert _span(ev, "actor_train", k + 1)[0] >= prev_end - 1e-3
    # every traversal of every MFC ran exactly once per dp rank
    for name in DUR:
        assert sorted(e[1] for e in ev if e[0] == name) == sorted(list(range(4)) * 2)
    barrier_m, barrier_times, barrier_wall = _run(window=1)
    bev = barrier_m.stream.events
    for k in range(3):
        assert _span(bev, "actor_gen", k + 1)[0] >= _span(bev, "critic_train", k)[1] - 1e-3
    # steady state: with look-ahead a step costs max(actor chain, critic chain), with the barrier their serial sum
    assert wall < barrier_wall - 0.15, (wall, barrier_wall)


def test_look_ahead_never_exceeds_the_window():
    m, _, _ = _run(window=2, n_steps=5)
    ev = m.stream.events
    for k in range(2, 5):
        # traversal k of anything starts only after every MFC of traversal k-2 has replied
        done = max(_span(ev, n, k - 2)[1] for n in DUR)
        assert min(_span(ev, n, k)[0] for n in DUR) >= done - 1e-3


def test_a_failing_mfc_stops_the_walk_instead_of_hanging_it():
    m = SimMaster(_master_cfg(2), 4, fail_at=("critic_train", 1))
    with pytest.raises(RuntimeError, match="injected failure"):
        m.run()
    # recover bookkeeping: step 0 is finalised, step 1 failed; the generation of step 2 may already have taken its prompts
    # (look-ahead), but nothing was trained on them: only the prompts of finalised steps count as consumed
    assert m.step == 1
    assert sorted(m._consumed_ids_this_epoch()) == sorted(m._ids_by_step[0]) and len(m._ids_by_step[0]) == BS
    assert 1 in m._ids_by_step and set(m._ids_by_step[1]).isdisjoint(m._ids_by_step[0])


def test_the_allocation_search_simulator_predicts_the_masters_step_time():
    """The native list-scheduling simulator behind `allocation_mode=search` (mesh exclusivity, DFG edges, "users of a role wait for
    its train step" across iterations) and the real master walking the same allocation over simulated workers must agree on the
    steady-state step time -- otherwise the search optimises a runtime that does not exist."""
    from realhf_b200.ops import host
    h = host()
    if h is None:
        pytest.skip("host extension not built")
    names = list(DUR)
    role = dict(actor_gen=0, actor_train=0, ref_inf=1, rew_inf=2, critic_inf=3, critic_train=3)
    kind = dict(actor_gen=0, rew_inf=1, ref_inf=1, critic_inf=1, actor_train=2, critic_train=2)
    mesh = dict(actor_gen=0, actor_train=0, ref_inf=0, rew_inf=1, critic_inf=1, critic_train=1)
    idx = {n: i for i, n in enumerate(names)}
    edges = [(idx["actor_gen"], idx[n]) for n in names if n != "actor_gen"]
    edges += [(idx[a], idx[b]) for a in ("rew_inf", "ref_inf", "critic_inf") for b in ("actor_train", "critic_train")]
    prob = dict(n_gpus=4, mem_cap=1e18, link_bw=1e18, n_iters=3, role_bytes=[0.0] * 4, meshes=[[0, 1], [2, 3]], edges=edges,
                realloc_latency_us=0.0,
                rpcs=[dict(name=n, role=role[n], kind=kind[n], cands=[(mesh[n], 2, 1, 1, DUR[n] * 1e6, 0.0, 0.0)]) for n in names])
    predicted = h.simulate_allocation(prob, [0] * len(names))["time_us"] / 1e6
    m, times, _ = _run(window=2, n_steps=5)
    steady = sorted(times[1:])[len(times[1:]) // 2]
    # the real walk can only be slower than the model (dispatch batching, event-loop latency on a busy CI machine)
    assert 0.95 * predicted <= steady <= 1.3 * predicted, (predicted, times)
