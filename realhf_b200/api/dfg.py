"""Dataflow graph of model function calls (MFCs).

Parity: `realhf/api/core/dfg.py` — `MFCDef` (name, n_seqs, interface type/impl, model name, input /
output keys with remaps, n_mbs, balanced_dp, hooks), `ParamReallocHook`, `OffloadHook`, and
`build_graph` (edge producer -> consumer per data key; keys with no producer come from the dataset).
"""

from __future__ import annotations

import dataclasses
from typing import Any, Dict, List, Optional, Tuple, Union

import networkx as nx

from realhf_b200.api.config import ModelFamily, ModelInterfaceAbstraction, ModelInterfaceType, ModelName


@dataclasses.dataclass
class OffloadHook:
    """Move the (non-trainable) model's parameters to pinned host memory after the call."""


@dataclasses.dataclass
class ParamReallocHook:
    """Re-shard parameters between two replicas of a role.

    Exactly one of source/target is set; the other end is the MFC's own model.
    target <- eta * source + (1 - eta) * target   (eta=1: plain copy; eta<1: EMA update).
    """

    source: Optional[ModelName] = None
    target: Optional[ModelName] = None
    eta: float = 1.0

    def __post_init__(self):
        if (self.source is None) == (self.target is None):
            raise ValueError("ParamReallocHook needs exactly one of `source` / `target`")


RPCHook = Union[OffloadHook, ParamReallocHook]


@dataclasses.dataclass
class MFCDef:
    name: str
    n_seqs: int
    interface_type: ModelInterfaceType
    interface_impl: ModelInterfaceAbstraction
    model_name: Union[str, ModelName]
    input_keys: Tuple = ()
    input_key_remap: Dict[str, str] = dataclasses.field(default_factory=dict)
    output_keys: Tuple = ()
    output_key_remap: Dict[str, str] = dataclasses.field(default_factory=dict)
    n_mbs: Optional[int] = None
    balanced_dp: bool = False
    log_return_value: bool = False
    model_type: Optional[Union[Any, ModelFamily]] = None
    model_path: Optional[str] = None
    _G: Optional[nx.DiGraph] = dataclasses.field(default=None, repr=False, compare=False)
    _pre_hooks: List[RPCHook] = dataclasses.field(default_factory=list, repr=False)
    _post_hooks: List[RPCHook] = dataclasses.field(default_factory=list, repr=False)

    def __post_init__(self):
        if isinstance(self.model_name, str):
            self.model_name = ModelName(role=self.model_name, replica_id=0)
        self.input_keys, self.output_keys = tuple(self.input_keys), tuple(self.output_keys)

    def __repr__(self):
        return f"MFCDef[{self.name}]"

    def __hash__(self):
        return hash(self.name)

    def __eq__(self, other):
        return isinstance(other, MFCDef) and other.name == self.name

    @property
    def role(self) -> str:
        return self.model_name.role

    def add_pre_hook(self, h: RPCHook):
        if isinstance(h, OffloadHook):
            raise ValueError("offload can only be a post hook")
        self._pre_hooks.append(h)

    def add_post_hook(self, h: RPCHook):
        self._post_hooks.append(h)

    # ---- graph queries (valid after build_graph)
    @property
    def is_src(self) -> bool:
        return len(list(self._G.predecessors(self.name))) == 0

    @property
    def is_dst(self) -> bool:
        return len(list(self._G.successors(self.name))) == 0

    @property
    def data_producers(self) -> Dict[str, ModelName]:
        return self._G.graph["data_producers"]

    @property
    def data_consumers(self) -> Dict[str, List[str]]:
        return self._G.graph["data_consumers"]

    @property
    def parents(self) -> List["MFCDef"]:
        return [self._G.nodes[x]["object"] for x in self._G.predecessors(self.name)]

    @property
    def children(self) -> List["MFCDef"]:
        return [self._G.nodes[x]["object"] for x in self._G.successors(self.name)]

    def all_successors(self) -> List["MFCDef"]:
        return [self._G.nodes[x]["object"] for x in nx.descendants(self._G, self.name)]

    @property
    def is_dst_of_model_role(self) -> bool:
        """No downstream MFC uses the same role (so post-hooks like offload are safe here)."""
        return not any(r.role == self.role for r in self.all_successors())

    @property
    def max_min_flow_seqs(self) -> int:
        return self._G.graph["max_min_flow_seqs"]


def build_graph(rpcs: List[MFCDef], verbose: bool = False) -> nx.DiGraph:
    names = [r.name for r in rpcs]
    if len(set(names)) != len(names):
        raise ValueError(f"duplicate MFC names: {names}")
    G = nx.DiGraph()
    producers: Dict[str, MFCDef] = {}
    for r in rpcs:
        G.add_node(r.name, object=r)
        for k in r.output_keys:
            if k in producers:
                raise ValueError(f"key `{k}` is produced by both {producers[k].name} and {r.name}")
            producers[k] = r
    consumers: Dict[str, List[str]] = {}
    for r in rpcs:
        for k in r.input_keys:
            consumers.setdefault(k, []).append(r.name)
            if k in producers and producers[k].name != r.name:
                p = producers[k]
                if G.has_edge(p.name, r.name):
                    G[p.name][r.name]["keys"].append(k)
                else:
                    G.add_edge(p.name, r.name, keys=[k])
    if not nx.is_directed_acyclic_graph(G):
        raise ValueError("the MFC graph has a cycle")
    G.graph["data_producers"] = {k: p.model_name for k, p in producers.items()}
    G.graph["data_consumers"] = consumers
    G.graph["dataset_keys"] = sorted(k for k in consumers if k not in producers)
    G.graph["max_min_flow_seqs"] = max(r.n_seqs for r in rpcs)
    for r in rpcs:
        r._G = G
    if verbose:
        for u, v, d in G.edges(data=True):
            print(f"{u} -> {v}: {d['keys']}")
    return G


def topological_levels(G: nx.DiGraph) -> List[List[str]]:
    return [sorted(gen) for gen in nx.topological_generations(G)]
