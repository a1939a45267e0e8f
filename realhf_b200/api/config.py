"""Config atoms shared by every layer.

Parity: reference `realhf/api/core/config.py:8-188` (Abstraction records, ModelName,
ModelFamily, ModelShardID and its `role@ppXX@mpXX@dpXX` string form).
"""

from __future__ import annotations

import dataclasses
import enum
import re
from typing import Any, Dict, Optional


def _abstraction(name: str, default_type: Optional[str] = None, extra: Optional[dict] = None):
    """Build a `(type_, args)` record class resolved by a `make_*` factory."""
    fields = [("type_", Optional[str], dataclasses.field(default=default_type)),
              ("args", Dict[str, Any], dataclasses.field(default_factory=dict))]
    for k, v in (extra or {}).items():
        fields.append((k, Any, dataclasses.field(default_factory=v)))
    cls = dataclasses.make_dataclass(name, fields)
    cls.__module__ = __name__
    return cls


DatasetAbstraction = _abstraction("DatasetAbstraction")
DataLoaderAbstraction = _abstraction("DataLoaderAbstraction", "packed")
ModelWrapperAbstraction = _abstraction("ModelWrapperAbstraction")
ModelAbstraction = _abstraction("ModelAbstraction", extra={"wrappers": list})
ModelBackendAbstraction = _abstraction("ModelBackendAbstraction")
ModelInterfaceAbstraction = _abstraction("ModelInterfaceAbstraction")


class ModelInterfaceType(enum.Enum):
    GENERATE = "generate"
    TRAIN_STEP = "train_step"
    EVALUATE = "evaluate"
    INFERENCE = "inference"


@dataclasses.dataclass(unsafe_hash=True, order=True, frozen=True)
class ModelName:
    """`role` identifies one set of weights; replicas of a role differ in layout/mesh."""

    role: str = "default"
    replica_id: int = 0

    @property
    def name(self):
        return str(self)

    def __str__(self):
        return f"{self.role}@{self.replica_id}"

    @classmethod
    def parse(cls, s: str) -> "ModelName":
        role, rid = s.rsplit("@", 1)
        return cls(role, int(rid))


@dataclasses.dataclass(unsafe_hash=True)
class ModelFamily:
    """e.g. ModelFamily("llama", 7, is_critic=False)."""

    _class: str = "llama"
    size: int = 0
    is_critic: bool = False

    def __repr__(self):
        s = f"{self._class}-{self.size}"
        return s + ("-critic" if self.is_critic else "")


@dataclasses.dataclass(unsafe_hash=True)
class ModelShardID:
    """One (pp, tp, dp) shard of a model replica."""

    model_name: ModelName
    dp_rank: int
    tp_rank: int
    pp_rank: int
    topo: Any = dataclasses.field(default=None, hash=False, compare=False)

    # reference code calls the tensor-parallel axis "mp"; keep an alias for users
    @property
    def mp_rank(self):
        return self.tp_rank

    @property
    def parallelism_rank(self):
        return self.topo.get_rank(pipe=self.pp_rank, model=self.tp_rank, data=self.dp_rank)

    @classmethod
    def from_parallelism_rank(cls, model_name, topo, parallelism_rank):
        c = topo.get_coord(parallelism_rank)
        return cls(model_name=model_name, dp_rank=c.data, tp_rank=c.model, pp_rank=c.pipe, topo=topo)

    def __repr__(self):
        return f"{self.model_name}@pp{self.pp_rank:02d}@mp{self.tp_rank:02d}@dp{self.dp_rank:02d}"

    _PAT = re.compile(r"^(.+)@pp(\d+)@mp(\d+)@dp(\d+)$")

    @classmethod
    def parse(cls, s: str, topo=None) -> "ModelShardID":
        m = cls._PAT.match(s)
        if m is None:
            raise ValueError(f"not a shard id: {s}")
        return cls(ModelName.parse(m.group(1)), dp_rank=int(m.group(4)), tp_rank=int(m.group(3)),
                   pp_rank=int(m.group(2)), topo=topo)


@dataclasses.dataclass
class StandaloneModelShardAbstraction:
    """Everything a model worker needs to build one shard."""

    id: ModelShardID
    model: ModelAbstraction
    backend: ModelBackendAbstraction
    eval_dataset: Optional[DatasetAbstraction] = None
    eval_bs: int = 128
    should_instantiate: bool = True
