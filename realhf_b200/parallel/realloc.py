"""Parameter reallocation: re-shard a role's weights between two (mesh, dp, tp, pp) layouts.

Parity: `realhf/impl/model/comm/param_realloc.py` (plan derivation :312-522, sender selection :82-138) and
`nn/real_llm_api.py:610-785` (execution, EMA patch).  The plan is pure index math over the sharding table
(`models/sharding.py`): for every destination shard and parameter, the element intervals it needs are
intersected with the intervals each source TP rank holds, giving (source offset, destination offset, length)
segments in the two flat buffers.

Execution differs by design.  The reference packs with `slice_intervals`, broadcasts over a dedicated NCCL
group per sender and unpacks with `set_intervals`.  Here a transfer is ONE launch of the segment-copy kernel
(`ops/csrc/segcopy.cu`) whose destination pointer is the peer GPU's flat buffer mapped through CUDA IPC
(`parallel/symm_mem.py`): the kernel writes the destination layout directly over NVLink — no pack buffer, no
unpack kernel, no per-pair communicator.  Without peer access (CPU/gloo, different nodes) it falls back to
pack -> batched isend/irecv -> unpack.  On NVSwitch every GPU reaches every peer at full bandwidth, so senders
are chosen only to spread load over source DP replicas (the reference's one-sender-per-node rule disappears).
"""

from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from realhf_b200.api.model import ReaLModelConfig
from realhf_b200.base.topology import ProcessTopology
from realhf_b200.models import sharding
from realhf_b200.models.real_model import build_layout
from realhf_b200.ops.functional import SegmentPlan


@dataclasses.dataclass
class Transfer:
    src_worker: int
    dst_worker: int
    src_off: Sequence[int]   # element offsets in the source flat buffer   } int64 numpy arrays when produced by `derive_plan`
    dst_off: Sequence[int]   # element offsets in the destination flat buffer } (a 70B reshard has tens of millions of row
    lens: Sequence[int]      # elements                                        } segments: python lists of them cost minutes)

    @property
    def numel(self) -> int:
        return int(np.asarray(self.lens, dtype=np.int64).sum())


@dataclasses.dataclass
class ReallocPlan:
    transfers: List[Transfer]
    dst_numel: Dict[int, int]  # dst worker -> flat numel of its destination shard


def _intersect(a: List[Tuple[int, int]], b: List[Tuple[int, int]]):
    """Sorted-by-start interval lists in the same coordinate system -> [(a_local_off, b_local_off, length)],
    where local offsets are positions inside the concatenation of each list's intervals (i.e. inside the shards)."""
    # shard order == list order; both lists are increasing in full-tensor coordinates by construction
    out = []
    i = j = 0
    a_base = b_base = 0
    while i < len(a) and j < len(b):
        a0, a1 = a[i]
        b0, b1 = b[j]
        lo, hi = max(a0, b0), min(a1, b1)
        if hi > lo:
            out.append((a_base + lo - a0, b_base + lo - b0, hi - lo))
        if a1 <= b1:
            a_base += a1 - a0
            i += 1
        else:
            b_base += b1 - b0
            j += 1
    return out


def _sorted_with_local(iv: List[Tuple[int, int]]):
    """Intervals may come in shard order that is not monotone (section splits are, row splits are): return them sorted
    by start together with each interval's local offset inside the shard."""
    loc, off = [], 0
    for a, b in iv:
        loc.append((a, b, off))
        off += b - a
    loc.sort()
    return loc


def _intersect_general(src_iv, dst_iv):
    s, d = _sorted_with_local(src_iv), _sorted_with_local(dst_iv)
    out = []
    i = j = 0
    while i < len(s) and j < len(d):
        a0, a1, ao = s[i]
        b0, b1, bo = d[j]
        lo, hi = max(a0, b0), min(a1, b1)
        if hi > lo:
            out.append((ao + lo - a0, bo + lo - b0, hi - lo))
        if a1 <= b1:
            i += 1
        else:
            j += 1
    return out


def _uncovered(segs, covered: List[Tuple[int, int]]):
    """Parts of `segs` ((src_off, dst_off, len), shard-local) whose destination range is not in `covered` yet; `covered`
    (sorted, disjoint destination ranges) is extended in place.  KV heads replicated over the source TP group (n_kv < tp) are
    held by several source ranks: the destination must receive each element exactly ONCE -- a second copy is harmless for a
    plain overwrite but applies an EMA merge (eta < 1) twice.

    One sweep over the segments (sorted by destination offset) against the sorted cover, then one merge of the cover with the
    new pieces: linear in len(segs) + len(covered).  (A column-split 4096-row weight contributes 4096 segments per source rank;
    the plan of a 7B model has ~2 M of them, so anything quadratic here costs minutes.)"""
    if not segs:
        return []
    segs = sorted(segs, key=lambda t: t[1])
    out, fresh = [], []
    ci, nc = 0, len(covered)
    for so, do, ln in segs:
        p0, p1 = do, do + ln
        while ci < nc and covered[ci][1] <= p0:
            ci += 1
        k = ci
        while p0 < p1:
            if k >= nc or covered[k][0] >= p1:      # nothing (more) of the cover inside [p0, p1)
                out.append((so + (p0 - do), p0, p1 - p0))
                fresh.append((p0, p1))
                break
            c0, c1 = covered[k]
            if c0 > p0:
                out.append((so + (p0 - do), p0, c0 - p0))
                fresh.append((p0, c0))
            p0 = max(p0, c1)
            k += 1
    if fresh:
        merged, i, j = [], 0, 0
        while i < nc or j < len(fresh):
            if j >= len(fresh) or (i < nc and covered[i][0] <= fresh[j][0]):
                r = covered[i]; i += 1
            else:
                r = fresh[j]; j += 1
            if merged and merged[-1][1] >= r[0]:
                if r[1] > merged[-1][1]:
                    merged[-1] = (merged[-1][0], r[1])
            else:
                merged.append(r)
        covered[:] = merged
    return out


def _spec_sources(spec, cfg: ReaLModelConfig, s_tp: int, d_tp: int, dtp: int, cache: dict):
    """[(source tp rank, src_off[], dst_off[], len[])] (numpy int64, shard-local element offsets): which source TP ranks supply
    destination TP rank `dtp` with which pieces of one parameter.  Depends only on the parameter's geometry, so every layer
    of the model shares the result through `cache`."""
    key = (spec.shape, spec.split_dim, spec.sections, spec.kv_sections, spec.expert_dim, s_tp, d_tp, dtp)
    hit = cache.get(key)
    if hit is not None:
        return hit
    numel = 1
    for d in sharding.shard_shape(spec, cfg, d_tp):
        numel *= d
    whole = (np.zeros(1, np.int64), np.zeros(1, np.int64), np.array([numel], np.int64))
    if s_tp == d_tp:
        out = [(dtp,) + whole]                         # same TP degree: the peer shard as a whole (also for replicated KV heads)
    elif spec.split_dim is None:
        out = [(dtp % s_tp,) + whole]                  # replicated tensor: one source copy is enough
    else:
        d_iv = sharding.shard_intervals(spec, cfg, dtp, d_tp)
        out, covered = [], []
        for stp in range(s_tp):
            segs = _uncovered(_intersect_general(sharding.shard_intervals(spec, cfg, stp, s_tp), d_iv), covered)
            if segs:
                arr = np.asarray(segs, dtype=np.int64).reshape(-1, 3)
                out.append((stp, np.ascontiguousarray(arr[:, 0]), np.ascontiguousarray(arr[:, 1]), np.ascontiguousarray(arr[:, 2])))
    cache[key] = out
    return out


def derive_plan(cfg: ReaLModelConfig, src_topo: ProcessTopology, src_workers: Sequence[int], dst_topo: ProcessTopology,
                dst_workers: Sequence[int], volumes_only: bool = False, for_worker: Optional[int] = None):
    """`*_workers[r]` = worker (GPU) index of layout-local rank r.  Critic / actor pairs must share the architecture.
    `volumes_only`: return {(source worker, destination worker): elements} without materialising the segments (what the
    allocation search's cost model needs; a 70B plan over 64 GPUs has tens of millions of segments).
    `for_worker`: materialise only the transfers this worker sends or receives (what its `ReallocExecutor` runs); `dst_numel`
    still covers every destination.

    Cost: the interval intersection runs once per distinct parameter geometry and (source tp, destination tp rank) pair; the
    per-layer work is numpy offset arithmetic.  LLaMA-7B, dp8 -> dp4*tp2 (2.1 M row segments before coalescing): ~1 s."""
    s_pp, s_dp, s_tp = src_topo.dims
    d_pp, d_dp, d_tp = dst_topo.dims
    src_stage = sharding.partition_pipeline_layers(cfg, s_pp)
    dst_stage = sharding.partition_pipeline_layers(cfg, d_pp)
    layer_to_src_pp = {l: p for p, (a, b) in src_stage.items() for l in range(a, b)}
    src_layouts = {p: build_layout(cfg, range(*src_stage[p]), s_tp)[0] for p in range(s_pp)}
    chunks: Dict[Tuple[int, int], List[Tuple[np.ndarray, np.ndarray, np.ndarray]]] = {}
    volumes: Dict[Tuple[int, int], int] = {}
    ln_sum: Dict[int, int] = {}      # id(len array of a cached geometry entry) -> its sum
    dst_numel: Dict[int, int] = {}
    cache: dict = {}
    for dpp in range(d_pp):
        d_slots, d_total = build_layout(cfg, range(*dst_stage[dpp]), d_tp)
        for ddp in range(d_dp):
            for dtp in range(d_tp):
                dw = dst_workers[dst_topo.get_rank(pipe=dpp, data=ddp, model=dtp)]
                dst_numel[dw] = d_total
                for name, dslot in d_slots.items():
                    li = int(name.split(".", 1)[0])
                    spp = layer_to_src_pp[li]
                    if name not in src_layouts[spp]:  # destination keeps a copy of the tied embedding as its head
                        assert cfg.tied_embedding and name.endswith("head.weight"), name
                        spp = layer_to_src_pp[0]
                        sslot = src_layouts[spp]["0.wte.weight"]
                    else:
                        sslot = src_layouts[spp][name]
                    for stp, so, do, ln in _spec_sources(dslot.spec, cfg, s_tp, d_tp, dtp, cache):
                        # pick the source DP replica: same GPU if possible, else spread by destination dp rank
                        cands = [src_workers[src_topo.get_rank(pipe=spp, data=k, model=stp)] for k in range(s_dp)]
                        sw = dw if dw in cands else cands[(ddp * d_tp + dtp) % s_dp]
                        if for_worker is not None and sw != for_worker and dw != for_worker:
                            continue
                        if volumes_only:
                            n = ln_sum.get(id(ln))
                            if n is None:
                                n = ln_sum[id(ln)] = int(ln.sum())
                            volumes[(sw, dw)] = volumes.get((sw, dw), 0) + n
                        else:
                            chunks.setdefault((sw, dw), []).append((so + sslot.offset, do + dslot.offset, ln))
    if volumes_only:
        return volumes
    transfers = []
    for (sw, dw), parts in sorted(chunks.items()):
        so, do, ln = (np.concatenate([p[i] for p in parts]) for i in range(3))
        so, do, ln = _coalesce_arrays(so, do, ln)
        transfers.append(Transfer(sw, dw, so, do, ln))
    return ReallocPlan(transfers, dst_numel)


def _coalesce_arrays(so: np.ndarray, do: np.ndarray, ln: np.ndarray):
    """Merge segments that are adjacent on both sides (whole parameters, consecutive rows of equal pitch...)."""
    if len(ln) <= 1:
        return so, do, ln
    joined = (so[1:] == so[:-1] + ln[:-1]) & (do[1:] == do[:-1] + ln[:-1])
    starts = np.flatnonzero(np.concatenate([[True], ~joined]))
    return so[starts], do[starts], np.add.reduceat(ln, starts)


class ReallocExecutor:
    """Runs a plan for the process of `my_worker`.  `src_flat` / `dst_flat` are this worker's flat buffers (or None)."""

    def __init__(self, plan: ReallocPlan, my_worker: int, elem_size: int, device, worker_to_rank=None):
        self.plan, self.me, self.es, self.device = plan, my_worker, elem_size, torch.device(device)
        self.w2r = worker_to_rank or (lambda w: w)
        self.local = [t for t in plan.transfers if t.src_worker == my_worker and t.dst_worker == my_worker]
        self.sends = [t for t in plan.transfers if t.src_worker == my_worker and t.dst_worker != my_worker]
        self.recvs = [t for t in plan.transfers if t.dst_worker == my_worker and t.src_worker != my_worker]
        es = elem_size
        arr = lambda x: np.asarray(x, dtype=np.int64) * es
        mk = lambda so, do, ln: SegmentPlan(arr(so), arr(do), arr(ln), self.device)
        self.local_plans = [mk(t.src_off, t.dst_off, t.lens) for t in self.local]
        # direct peer-store plans (source offsets -> destination offsets in the PEER's flat buffer)
        self.direct_plans = [mk(t.src_off, t.dst_off, t.lens) for t in self.sends]
        # pack / unpack plans for the NCCL / gloo fallback
        self.pack_plans = [mk(t.src_off, _prefix(t.lens), t.lens) for t in self.sends]
        self.unpack_plans = [mk(_prefix(t.lens), t.dst_off, t.lens) for t in self.recvs]

    def dst_numel(self) -> Optional[int]:
        return self.plan.dst_numel.get(self.me)

    def whole_local_copy(self) -> bool:
        """This worker's part of the plan is ONE local copy of its entire source buffer onto its entire destination buffer, and
        nothing arrives from peers: the destination shard IS the source shard (same tp / pp position, e.g. a dp4 generation
        replica on half of the GPUs of a dp8 training layout).  The caller may then alias the buffers instead of copying."""
        if self.recvs or len(self.local) != 1:
            return False
        t = self.local[0]
        n = self.plan.dst_numel.get(self.me)
        return len(t.lens) == 1 and int(t.src_off[0]) == 0 and int(t.dst_off[0]) == 0 and n is not None and int(t.lens[0]) == n

    def run(self, src_flat: Optional[torch.Tensor], dst_flat: Optional[torch.Tensor], eta: float = 1.0,
            peer_dst_ptrs: Optional[Dict[int, int]] = None, group=None, notify: bool = False, skip_local: bool = False):
        """peer_dst_ptrs: dst worker -> device address of its destination flat buffer mapped into this process.
        When given for every send, transfers are direct peer stores; else pack + isend/irecv + unpack.
        `skip_local`: the destination aliases the source on this worker (see `whole_local_copy`)."""
        if not skip_local:
            for pl in self.local_plans:
                pl.run(src_flat, dst_flat, eta=eta)
        direct = peer_dst_ptrs is not None and all(t.dst_worker in peer_dst_ptrs for t in self.sends)
        if direct:
            for t, pl in zip(self.sends, self.direct_plans):
                pl.run(src_flat, None, dst_ptr=peer_dst_ptrs[t.dst_worker], eta=eta)
            if notify:
                # completion tokens: a 4-byte send ordered after each store kernel, a matching recv on the destination, so the
                # destination's stream continues only once every sender's stores have landed (no group / barrier needed)
                dev = self.device
                tok = torch.ones(1, dtype=torch.int32, device=dev)
                ops = [dist.P2POp(dist.isend, tok, self.w2r(t.dst_worker), group) for t in self.sends]
                bufs = [torch.empty(1, dtype=torch.int32, device=dev) for _ in self.recvs]
                ops += [dist.P2POp(dist.irecv, b, self.w2r(t.src_worker), group) for b, t in zip(bufs, self.recvs)]
                if ops:
                    for w in dist.batch_isend_irecv(ops):
                        w.wait()
            return
        ops, staged = [], []
        for t, pl in zip(self.sends, self.pack_plans):
            buf = torch.empty(t.numel, dtype=src_flat.dtype, device=src_flat.device)
            pl.run(src_flat, buf)
            ops.append(dist.P2POp(dist.isend, buf, self.w2r(t.dst_worker), group))
            staged.append(buf)
        rbufs = []
        for t in self.recvs:
            buf = torch.empty(t.numel, dtype=dst_flat.dtype, device=dst_flat.device)
            ops.append(dist.P2POp(dist.irecv, buf, self.w2r(t.src_worker), group))
            rbufs.append(buf)
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for buf, pl in zip(rbufs, self.unpack_plans):
            pl.run(buf, dst_flat, eta=eta)


def _prefix(lens: List[int]) -> np.ndarray:
    """Exclusive prefix sums: offsets of the segments inside a packed transfer buffer."""
    ln = np.asarray(lens, dtype=np.int64)
    return np.cumsum(ln) - ln
