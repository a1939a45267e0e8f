"""Cluster specification: what the launcher needs to know about the machines it runs on.

Parity: `realhf/base/cluster.py:17-133` (JSON file at `$CLUSTER_SPEC_PATH`: cluster type / name, fileroot, container images and
mounts for Slurm + pyxis, regexes from node names to node / GPU types, node-name prefix).  Additions for this framework: the
node shape (`n_gpus_per_node`, `gpu_memory_gb`: the allocation search and the heuristic allocator size their plans from them;
the defaults describe an 8 x B200 NVSwitch node) and the Slurm `partition`.

Without a spec file the defaults describe one local node, so single-node use needs no configuration at all.
"""

from __future__ import annotations

import dataclasses
import json
import os
import re
from typing import Dict, List, Optional


@dataclasses.dataclass
class ClusterSpec:
    cluster_type: str = "local"                 # local | slurm
    cluster_name: str = "local"
    fileroot: Optional[str] = None              # shared filesystem root for logs / checkpoints / recover states
    default_mount: Optional[str] = None         # "src:dst,src:dst" for the container runtime
    node_type_from_node_name: Dict[str, str] = dataclasses.field(default_factory=dict)   # regex -> node type
    gpu_type_from_node_name: Dict[str, str] = dataclasses.field(default_factory=dict)    # regex -> gpu type
    cpu_image: Optional[str] = None
    gpu_image: Optional[str] = None
    node_name_prefix: str = "NODE"
    partition: Optional[str] = None
    n_gpus_per_node: int = 8
    gpu_memory_gb: float = 180.0

    @classmethod
    def from_file(cls, path: str) -> "ClusterSpec":
        with open(path) as f:
            raw = json.load(f)
        known = {f.name for f in dataclasses.fields(cls)}
        unknown = sorted(set(raw) - known)
        if unknown:
            raise ValueError(f"unknown keys in cluster spec {path}: {unknown}")
        for k in ("cluster_type", "cluster_name", "fileroot"):
            if k not in raw:
                raise ValueError(f"cluster spec {path} lacks the required key `{k}`")
        spec = cls(**raw)
        if spec.cluster_type not in ("local", "slurm"):
            raise ValueError(f"cluster_type must be `local` or `slurm`, got {spec.cluster_type!r}")
        return spec

    # ---- node-name helpers (Slurm)
    @staticmethod
    def _lookup(table: Dict[str, str], node_name: str, what: str) -> str:
        for pattern, value in table.items():
            if re.match(pattern, node_name):
                return value
        raise KeyError(f"no {what} pattern matches node `{node_name}`")

    def node_type(self, node_name: str) -> str:
        return "default" if self.cluster_type != "slurm" else self._lookup(self.node_type_from_node_name, node_name, "node-type")

    def gpu_type(self, node_name: str) -> str:
        return "b200" if self.cluster_type != "slurm" else self._lookup(self.gpu_type_from_node_name, node_name, "gpu-type")

    def node_is_type(self, node_name: str, node_type) -> bool:
        """`node_type` None matches everything; a list matches any member."""
        if node_type is None:
            return True
        types = [node_type] if isinstance(node_type, str) else list(node_type)
        return self.node_type(node_name) in types

    def node_names(self, indices: List[int], width: int = 2) -> List[str]:
        """1-based node indices -> names (`NODE01`, ...), the naming the `nodelist` / device-mesh strings use."""
        return [f"{self.node_name_prefix}{i:0{width}d}" for i in indices]

    def image(self, gpu: bool) -> Optional[str]:
        return self.gpu_image if gpu else self.cpu_image


def parse_nodelist(nodelist: str) -> List[str]:
    """Slurm hostlist syntax -> node names: `NODE[01-03,07],gpu12` -> NODE01 NODE02 NODE03 NODE07 gpu12 (zero padding kept).
    (Reference: base/slurm_utils.py:13-32, which shells out to `scontrol show hostnames`.)"""
    out: List[str] = []
    for item in re.findall(r"[^,\[]+(?:\[[^\]]*\])?", nodelist.replace(" ", "")):
        m = re.fullmatch(r"(.*)\[([^\]]*)\]", item)
        if not m:
            out.append(item)
            continue
        prefix, body = m.groups()
        for part in body.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                if int(hi) < int(lo):
                    raise ValueError(f"descending range `{part}` in nodelist `{nodelist}`")
                out.extend(f"{prefix}{i:0{len(lo)}d}" for i in range(int(lo), int(hi) + 1))
            elif part:
                out.append(prefix + part)
    if not out or len(set(out)) != len(out):
        raise ValueError(f"empty or repeated nodes in nodelist `{nodelist}`")
    return out


def format_nodelist(nodes: List[str]) -> str:
    """Inverse of `parse_nodelist` for names of the form <prefix><digits>: consecutive numbers are folded into ranges."""
    groups: Dict[tuple, List[int]] = {}
    plain: List[str] = []
    for n in nodes:
        m = re.fullmatch(r"(.*?)(\d+)", n)
        if m:
            groups.setdefault((m.group(1), len(m.group(2))), []).append(int(m.group(2)))
        else:
            plain.append(n)
    parts = []
    for (prefix, width), nums in groups.items():
        nums = sorted(set(nums))
        runs, start, prev = [], nums[0], nums[0]
        for x in nums[1:] + [None]:
            if x is not None and x == prev + 1:
                prev = x
                continue
            runs.append(f"{start:0{width}d}" if start == prev else f"{start:0{width}d}-{prev:0{width}d}")
            if x is not None:
                start = prev = x
        parts.append(f"{prefix}{runs[0]}" if len(runs) == 1 and "-" not in runs[0] else f"{prefix}[{','.join(runs)}]")
    return ",".join(parts + plain)


_SPEC: Optional[ClusterSpec] = None
_SPEC_PATH: Optional[str] = None


def spec() -> ClusterSpec:
    """The process-wide spec: loaded from `$CLUSTER_SPEC_PATH` on first use (re-read if the variable changes)."""
    global _SPEC, _SPEC_PATH
    path = os.environ.get("CLUSTER_SPEC_PATH") or None
    if _SPEC is None or path != _SPEC_PATH:
        _SPEC = ClusterSpec.from_file(path) if path else ClusterSpec()
        _SPEC_PATH = path
    return _SPEC


def node_name_is_node_type(node_name: str, node_type=None) -> bool:
    return spec().node_is_type(node_name, node_type)
