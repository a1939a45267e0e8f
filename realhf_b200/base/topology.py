"""3-D process topology and the explicit `ParallelContext` handed to every layer.

Parity: `realhf/base/topology.py` (ProcessTopology, PipeModelDataParallelTopology, ParallelGrid,
new_or_get_group, FakeGrid).  Design difference: the reference resolves process groups through a
per-process global keyed by the "current model name" (`constants.model_scope`).  Here every model
shard owns a `ParallelContext` object that carries its coordinates and groups explicitly, so several
models with different layouts can live in one process without global state.
"""

from __future__ import annotations

import dataclasses
import itertools
from typing import Dict, List, NamedTuple, Optional, Tuple

import torch
import torch.distributed as dist


class ProcessCoord(NamedTuple):
    pipe: int
    data: int
    model: int  # tensor-parallel axis (the reference calls it "model")


class ProcessTopology:
    """Cartesian rank <-> (pipe, data, model) mapping; `model` is the fastest-varying axis so that a
    tensor-parallel group is a run of consecutive ranks (same NVSwitch domain)."""

    AXES = ("pipe", "data", "model")

    def __init__(self, num_pp: int, num_dp: int, num_tp: int):
        self.dims = (num_pp, num_dp, num_tp)
        self._coord2rank: Dict[ProcessCoord, int] = {}
        for r, c in enumerate(itertools.product(*(range(d) for d in self.dims))):
            self._coord2rank[ProcessCoord(*c)] = r
        self._rank2coord = {r: c for c, r in self._coord2rank.items()}

    def __eq__(self, other):
        return isinstance(other, ProcessTopology) and self.dims == other.dims

    def __hash__(self):
        return hash(self.dims)

    def __repr__(self):
        return f"Topology(pp={self.dims[0]}, dp={self.dims[1]}, tp={self.dims[2]})"

    def get_dim(self, axis: str) -> int:
        return self.dims[self.AXES.index(axis)]

    def world_size(self) -> int:
        return self.dims[0] * self.dims[1] * self.dims[2]

    def get_rank(self, **coord) -> int:
        return self._coord2rank[ProcessCoord(**coord)]

    def get_coord(self, rank: int) -> ProcessCoord:
        return self._rank2coord[rank]

    def get_axis_comm_lists(self, axis: str) -> List[List[int]]:
        """Rank lists that vary only along `axis`."""
        ai = self.AXES.index(axis)
        others = [i for i in range(3) if i != ai]
        out = []
        for fixed in itertools.product(*(range(self.dims[i]) for i in others)):
            ranks = []
            for v in range(self.dims[ai]):
                c = [0, 0, 0]
                c[ai] = v
                for i, f in zip(others, fixed):
                    c[i] = f
                ranks.append(self._coord2rank[ProcessCoord(*c)])
            out.append(ranks)
        return out

    def filter_match(self, **kw) -> List[int]:
        return sorted(r for r, c in self._rank2coord.items() if all(getattr(c, k) == v for k, v in kw.items()))

    def get_axis_list(self, axis: str, idx: int) -> List[int]:
        return self.filter_match(**{axis: idx})


class PipeModelDataParallelTopology(ProcessTopology):
    """Name kept from the reference; also records sequence-parallel / grad-checkpoint flags of the layout."""

    def __init__(self, num_pp: int, num_mp: int, num_dp: int, sequence_parallel: bool = False,
                 gradient_checkpointing: bool = False, max_prompt_len: Optional[int] = None,
                 gradient_accumulation_fusion: bool = False):
        super().__init__(num_pp=num_pp, num_dp=num_dp, num_tp=num_mp)
        self.sequence_parallel = sequence_parallel
        self.gradient_checkpointing = gradient_checkpointing
        self.max_prompt_len = max_prompt_len
        self.gradient_accumulation_fusion = gradient_accumulation_fusion


def decompose_to_three_factors(n: int) -> List[Tuple[int, int, int]]:
    out = []
    for a in range(1, n + 1):
        if n % a:
            continue
        for b in range(1, n // a + 1):
            if (n // a) % b == 0:
                out.append((a, b, n // a // b))
    return out


# ------------------------------------------------------------------------------------------- groups

_GROUPS: Dict[Tuple[Tuple[int, ...], Optional[str]], "dist.ProcessGroup"] = {}


def new_or_get_group(ranks: List[int], backend: Optional[str] = None):
    """Cached `dist.new_group`; every process of the world must call it with the same arguments in the same order."""
    key = (tuple(sorted(ranks)), backend)
    if key not in _GROUPS:
        _GROUPS[key] = dist.new_group(list(key[0]), backend=backend)
    return _GROUPS[key]


def destroy_all_comm_groups():
    for g in _GROUPS.values():
        try:
            dist.destroy_process_group(g)
        except Exception:
            pass
    _GROUPS.clear()
    if dist.is_initialized():
        dist.destroy_process_group()


@dataclasses.dataclass
class ParallelContext:
    """Coordinates + process groups of ONE model shard.  `ranks[i]` is the global rank of model-local rank i."""

    topo: ProcessTopology
    ranks: List[int]
    local_rank: int  # model-local rank of this process, -1 if this process is not part of the model
    tp_group: Optional["dist.ProcessGroup"] = None
    dp_group: Optional["dist.ProcessGroup"] = None
    pp_group: Optional["dist.ProcessGroup"] = None
    model_group: Optional["dist.ProcessGroup"] = None
    tp_dp_group: Optional["dist.ProcessGroup"] = None
    embedding_group: Optional["dist.ProcessGroup"] = None
    sequence_parallel: bool = False
    gradient_checkpointing: bool = False
    symm: Optional[object] = None  # parallel.symm_mem.SymmetricWorkspace for fused TP kernels

    @property
    def coord(self) -> ProcessCoord:
        return self.topo.get_coord(self.local_rank)

    pp_rank = property(lambda self: self.coord.pipe)
    dp_rank = property(lambda self: self.coord.data)
    tp_rank = property(lambda self: self.coord.model)
    pp_size = property(lambda self: self.topo.dims[0])
    dp_size = property(lambda self: self.topo.dims[1])
    tp_size = property(lambda self: self.topo.dims[2])

    @property
    def is_member(self) -> bool:
        return self.local_rank >= 0

    def global_rank(self, **coord) -> int:
        return self.ranks[self.topo.get_rank(**coord)]

    @property
    def is_dp_head(self) -> bool:
        """The shard that reports results to the master: tp rank 0 of the last pipeline stage."""
        return self.tp_rank == 0 and self.pp_rank == self.pp_size - 1

    def pp_prev(self) -> int:
        c = self.coord
        return self.global_rank(pipe=(c.pipe - 1) % self.pp_size, data=c.data, model=c.model)

    def pp_next(self) -> int:
        c = self.coord
        return self.global_rank(pipe=(c.pipe + 1) % self.pp_size, data=c.data, model=c.model)

    @classmethod
    def single(cls) -> "ParallelContext":
        """World of one: no groups, every collective is the identity."""
        return cls(topo=ProcessTopology(1, 1, 1), ranks=[0], local_rank=0)

    @classmethod
    def fake(cls, topo: ProcessTopology, local_rank: int) -> "ParallelContext":
        """Rank math without process groups (the reference's FakeGrid): for planners and unit tests."""
        return cls(topo=topo, ranks=list(range(topo.world_size())), local_rank=local_rank)

    @classmethod
    def build(cls, topo: ProcessTopology, ranks: List[int], my_global_rank: int, backend: Optional[str] = None,
              sequence_parallel: bool = False, gradient_checkpointing: bool = False) -> "ParallelContext":
        """Create every group of the layout.  Collective over the whole world (all processes call it with
        identical arguments), also on processes that are not members of this model."""
        assert len(ranks) == topo.world_size(), (ranks, topo)
        local = ranks.index(my_global_rank) if my_global_rank in ranks else -1
        ctx = cls(topo=topo, ranks=list(ranks), local_rank=local, sequence_parallel=sequence_parallel,
                  gradient_checkpointing=gradient_checkpointing)
        g = lambda rs: [ranks[r] for r in rs]
        ctx.model_group = new_or_get_group(list(ranks), backend)
        for axis, attr in (("model", "tp_group"), ("data", "dp_group"), ("pipe", "pp_group")):
            for rs in topo.get_axis_comm_lists(axis):
                grp = new_or_get_group(g(rs), backend)
                if local in rs:
                    setattr(ctx, attr, grp)
        for pp in range(topo.dims[0]):  # tp x dp slab of one stage (grad-norm / stats reductions)
            rs = topo.filter_match(pipe=pp)
            grp = new_or_get_group(g(rs), backend)
            if local in rs:
                ctx.tp_dp_group = grp
        if topo.dims[0] > 1:  # tied embeddings: first + last stage with the same (dp, tp)
            for dp in range(topo.dims[1]):
                for tp in range(topo.dims[2]):
                    rs = sorted({topo.get_rank(pipe=0, data=dp, model=tp), topo.get_rank(pipe=topo.dims[0] - 1, data=dp, model=tp)})
                    grp = new_or_get_group(g(rs), backend)
                    if local in rs:
                        ctx.embedding_group = grp
        return ctx


# Names the reference exposes
ParallelGrid = ParallelContext
FakeGrid = ParallelContext.fake
