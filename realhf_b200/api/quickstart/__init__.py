"""User-facing configuration dataclasses and the dotted-override command-line parser.

Parity: `realhf/api/quickstart/{model,dataset,device_mesh,entrypoint}.py`.  Hydra / OmegaConf are not
available offline, so `parse_overrides` implements the same surface natively: `key.sub=value` pairs over nested
dataclasses, `null` -> None, booleans, numbers, `[a,b]` lists, and names of registered experiments as sub-commands.
"""

from __future__ import annotations

import dataclasses
import enum
import json
import re
import typing
from typing import Any, Dict, List, Optional, Tuple, Union, get_args, get_origin, get_type_hints

import numpy as np

from realhf_b200.api.config import ModelFamily
from realhf_b200.engine.optim import OptimizerConfig  # same knobs as the reference's OptimizerConfig

# ------------------------------------------------------------------------------------------- model / parallelism


@dataclasses.dataclass(unsafe_hash=True)
class ParallelismConfig:
    model_parallel_size: int = 1       # tensor parallel degree (name kept from the reference)
    pipeline_parallel_size: int = 1
    data_parallel_size: int = 1
    use_sequence_parallel: bool = False

    def __str__(self):
        return (f"Parallel(mp={self.model_parallel_size},pp={self.pipeline_parallel_size},dp={self.data_parallel_size}"
                f"{',sp' if self.use_sequence_parallel else ''})")

    @property
    def world_size(self) -> int:
        return self.model_parallel_size * self.pipeline_parallel_size * self.data_parallel_size


def parallelism_eq(a: ParallelismConfig, b: ParallelismConfig) -> bool:
    return (a.model_parallel_size, a.pipeline_parallel_size, a.data_parallel_size) == \
        (b.model_parallel_size, b.pipeline_parallel_size, b.data_parallel_size)


@dataclasses.dataclass
class LoRAConfig:
    dim: int = 32
    scaling: float = 32.0


@dataclasses.dataclass
class ModelTrainEvalConfig:
    type: ModelFamily = dataclasses.field(default_factory=lambda: ModelFamily("llama", 7, False))
    backend: str = "train"  # train (aliases: megatron, deepspeed) | inference
    path: str = ""
    lora: Optional[LoRAConfig] = None
    gradient_checkpointing: Union[bool, str] = True   # true | false | auto (recompute only the blocks the free HBM requires)
    enable_fp16: bool = False
    enable_bf16: bool = True
    offload: bool = False
    zero_stage: int = 1
    optimizer: Optional[OptimizerConfig] = dataclasses.field(default_factory=OptimizerConfig)
    init_from_scratch: bool = False
    init_critic_from_actor: bool = False
    expert_parallel: bool = False   # MoE models: partition whole experts over the tensor-parallel group (parallel/ep.py)


# ------------------------------------------------------------------------------------------- datasets


@dataclasses.dataclass
class PromptAnswerDatasetConfig:
    train_path: str = ""
    valid_path: str = ""
    max_seqlen: int = 1024
    train_bs_n_seqs: int = 256
    valid_bs_n_seqs: int = 256
    pad_to_max_length: bool = False


@dataclasses.dataclass
class PairedComparisonDatasetConfig:
    train_path: str = ""
    valid_path: str = ""
    max_pairs_per_prompt: int = 2
    max_seqlen: int = 1024
    train_bs_n_seqs: int = 256
    valid_bs_n_seqs: int = 256


@dataclasses.dataclass
class PromptOnlyDatasetConfig:
    path: str = ""
    max_prompt_len: int = 256
    train_bs_n_seqs: int = 256
    pad_to_max_length: bool = False


# ------------------------------------------------------------------------------------------- device meshes


@dataclasses.dataclass
class DeviceMesh:
    """A 0/1 map over (nodes x gpus) inside the global cluster mesh.  Valid sub-meshes are 1/2/4/8 consecutive
    GPUs of one node, or whole nodes (reference: device_mesh.py:145-182)."""

    n_nodes: int
    n_gpus_per_node: int
    mapping: np.ndarray
    global_mesh_name: Optional[str] = None
    name: Optional[str] = None
    gpu_memory_capacity: int = 180 * (1024 ** 3)  # B200

    def __post_init__(self):
        self.mapping = np.asarray(self.mapping, dtype=np.int32)
        assert self.mapping.shape == (self.n_nodes, self.n_gpus_per_node), (self.mapping.shape, self.n_nodes, self.n_gpus_per_node)

    def __eq__(self, other):
        return isinstance(other, DeviceMesh) and self.mapping.shape == other.mapping.shape and bool((self.mapping == other.mapping).all())

    def __hash__(self):
        return hash(self.mapping.tobytes())

    def __repr__(self):
        return f"DeviceMesh({self.name or self.global_ranks()})"

    @property
    def n_gpus(self) -> int:
        return int(self.mapping.sum())

    def global_ranks(self) -> List[int]:
        """Worker indices (node-major) covered by this mesh."""
        return [int(i) for i in np.flatnonzero(self.mapping.reshape(-1))]

    def overlap(self, other: "DeviceMesh") -> bool:
        return bool((self.mapping * other.mapping).any())

    def contain(self, other: "DeviceMesh") -> bool:
        return bool(((self.mapping - other.mapping) >= 0).all())

    def sub_device_meshes(self, min_n_gpus: int = 1) -> List["DeviceMesh"]:
        out = []
        sizes = [s for s in (1, 2, 4, 8, 16) if min_n_gpus <= s <= self.n_gpus_per_node and self.n_gpus_per_node % s == 0]
        for node in range(self.n_nodes):
            if not self.mapping[node].any():
                continue
            for s in sizes:
                for start in range(0, self.n_gpus_per_node, s):
                    m = np.zeros_like(self.mapping)
                    m[node, start:start + s] = 1
                    if ((self.mapping - m) >= 0).all():
                        out.append(DeviceMesh(self.n_nodes, self.n_gpus_per_node, m, self.global_mesh_name))
        for n in range(2, self.n_nodes + 1):  # whole consecutive nodes
            for start in range(0, self.n_nodes - n + 1):
                m = np.zeros_like(self.mapping)
                m[start:start + n] = 1
                if ((self.mapping - m) >= 0).all():
                    out.append(DeviceMesh(self.n_nodes, self.n_gpus_per_node, m, self.global_mesh_name))
        return out


def make_device_mesh_from_name(global_mesh_name: Optional[str], name: str, n_nodes: int = 1, n_gpus_per_node: int = 8) -> DeviceMesh:
    """`NODE01:0,1,2,3` (GPUs of one node) or `NODE[01-02]` / `NODE01,NODE02` (whole nodes).  Node numbers are
    1-based positions inside the cluster's node list; any prefix is accepted."""
    m = np.zeros((n_nodes, n_gpus_per_node), dtype=np.int32)

    def node_idx(tok: str) -> int:
        num = re.search(r"(\d+)$", tok)
        if not num:
            raise ValueError(f"cannot parse node name `{tok}`")
        return (int(num.group(1)) - 1) % max(n_nodes, 1)

    if ":" in name:
        node, gpus = name.split(":")
        gl = [int(x) for x in gpus.split(",")]
        if len(gl) not in (1, 2, 4, 8, 16) or gl != list(range(gl[0], gl[0] + len(gl))) or gl[0] % len(gl) != 0:
            raise ValueError(f"invalid device mesh `{name}`: need 1/2/4/8 consecutive, aligned GPUs")
        m[node_idx(node), gl] = 1
    else:
        from realhf_b200.base.cluster import parse_nodelist
        for nd in parse_nodelist(name):  # full Slurm hostlist syntax: NODE[01-02,05],NODE07
            m[node_idx(nd)] = 1
    return DeviceMesh(n_nodes, n_gpus_per_node, m, global_mesh_name, name)


def find_parallel_strategies(mesh: DeviceMesh) -> List[ParallelismConfig]:
    """All (tp, pp, dp) factorizations of the mesh with tp confined to one node."""
    n = mesh.n_gpus
    out = []
    for tp in (1, 2, 4, 8):
        if tp > min(n, mesh.n_gpus_per_node) or n % tp:
            continue
        rest = n // tp
        for pp in range(1, rest + 1):
            if rest % pp == 0:
                out.append(ParallelismConfig(tp, pp, rest // pp))
    return out


@dataclasses.dataclass
class MFCConfig:
    """Per-MFC knobs exposed on the command line (e.g. `actor_train.parallel.model_parallel_size=2`)."""

    n_mbs: Optional[int] = None
    parallel: ParallelismConfig = dataclasses.field(default_factory=ParallelismConfig)
    device_mesh: Optional[str] = None


@dataclasses.dataclass
class RPCAllocation:
    rpc: Any  # MFCDef | str
    device_mesh: DeviceMesh
    parallel: ParallelismConfig


# ------------------------------------------------------------------------------------------- override parser

_NONE = ("null", "none", "None", "~")


def _convert(value: str, tp) -> Any:
    origin = get_origin(tp)
    if origin is Union:
        args = [a for a in get_args(tp) if a is not type(None)]
        if value in _NONE:
            return None
        for a in args:
            try:
                return _convert(value, a)
            except (ValueError, TypeError):
                continue
        raise ValueError(f"cannot parse `{value}` as {tp}")
    if value in _NONE and tp is not str:
        return None
    if tp is bool:
        if value.lower() in ("true", "1", "yes"):
            return True
        if value.lower() in ("false", "0", "no"):
            return False
        raise ValueError(f"`{value}` is not a bool")
    if tp is int:
        return int(value)
    if tp is float:
        return float(value)
    if tp is str or tp is Any:
        return value
    if isinstance(tp, type) and issubclass(tp, enum.Enum):
        return tp(value)
    if origin in (list, List, tuple, Tuple):
        inner = get_args(tp)[0] if get_args(tp) else str
        body = value.strip()
        if body.startswith("[") and body.endswith("]"):
            body = body[1:-1]
        items = [x.strip() for x in body.split(",") if x.strip()]
        return [(_convert(x, inner)) for x in items]
    if origin in (dict, Dict):
        return json.loads(value)
    try:
        return json.loads(value)
    except json.JSONDecodeError:
        return value


def apply_override(obj: Any, dotted: str, raw: str):
    parts = dotted.split(".")
    cur = obj
    for i, p in enumerate(parts):
        if isinstance(cur, dict):
            if i == len(parts) - 1:
                cur[p] = _convert(raw, Any)
                return
            cur = cur.setdefault(p, {})
            continue
        if not dataclasses.is_dataclass(cur):
            raise AttributeError(f"`{'.'.join(parts[:i])}` is not a config group")
        fields = {f.name: f for f in dataclasses.fields(cur)}
        if p not in fields:
            raise AttributeError(f"unknown option `{dotted}` (no field `{p}` in {type(cur).__name__}; "
                                 f"choices: {sorted(fields)})")
        hints = get_type_hints(type(cur))
        if i == len(parts) - 1:
            setattr(cur, p, _convert(raw, hints.get(p, Any)))
            return
        nxt = getattr(cur, p)
        if nxt is None:  # Optional[dataclass] left empty: instantiate it
            tp = hints[p]
            cands = [a for a in get_args(tp) if dataclasses.is_dataclass(a)] if get_origin(tp) is Union else [tp]
            nxt = cands[0]()
            setattr(cur, p, nxt)
        cur = nxt


def parse_overrides(cfg: Any, argv: List[str]) -> Any:
    """Apply `a.b.c=value` arguments to a dataclass instance (in place) and re-run `__post_init__` checks."""
    for arg in argv:
        if "=" not in arg:
            raise ValueError(f"expected key=value, got `{arg}`")
        k, v = arg.split("=", 1)
        k = k.lstrip("+")
        if len(v) >= 2 and v[0] == v[-1] and v[0] in "'\"":
            v = v[1:-1]
        apply_override(cfg, k, v)
    post = getattr(cfg, "__post_init__", None)
    if post is not None:
        post()
    return cfg


QUICKSTART_EXPERIMENTS: Dict[str, Any] = {}


def register_quickstart_exp(name: str, cls):
    """Make `python -m realhf_b200.apps.quickstart <name> k=v ...` available."""
    if "_" in name:
        raise ValueError("experiment names must not contain `_` (used as a separator in run names)")
    QUICKSTART_EXPERIMENTS[name] = cls
