"""Slurm host-list helpers (parity: `realhf/base/slurm_utils.py`).  The reference shells out to `scontrol show hostnames` /
`scontrol show hostlistsorted`; these are pure functions (`base/cluster.py::parse_nodelist` does the expansion), so they also work
on a machine without Slurm -- which is where allocations are planned and tested."""

from __future__ import annotations

import re
import shutil
from typing import List, Sequence

import numpy as np

from realhf_b200.base.cluster import parse_nodelist as _expand


def parse_node_id(node_name: str, prefix: str) -> int:
    return int(node_name.split(prefix)[-1])


def parse_nodelist(nodelist: str, prefix: str) -> List[str]:
    """`NODE[01-03],NODE07` -> names, checked against the cluster's node-name prefix."""
    nodes = _expand(nodelist)
    bad = [n for n in nodes if not n.startswith(prefix)]
    if bad:
        raise ValueError(f"nodes {bad} do not carry the cluster prefix `{prefix}`")
    return nodes


def nodelist_from_nodes(nodes: Sequence[str], prefix: str) -> str:
    """Names -> compact host list (`NODE[01-03,07]`): consecutive ids of equal width fold into ranges."""
    if not nodes:
        return ""
    ids = sorted({(len(n) - len(prefix), parse_node_id(n, prefix)) for n in nodes})
    parts, i = [], 0
    while i < len(ids):
        w, a = ids[i]
        j = i
        while j + 1 < len(ids) and ids[j + 1] == (w, ids[j][1] + 1):
            j += 1
        parts.append(f"{a:0{w}d}" if j == i else f"{a:0{w}d}-{ids[j][1]:0{w}d}")
        i = j + 1
    if len(parts) == 1 and "-" not in parts[0]:
        return prefix + parts[0]
    return f"{prefix}[{','.join(parts)}]"


def are_ones_contiguous(binary_array: np.ndarray) -> bool:
    """The 1-entries of a 0/1 vector form one block (a device mesh must own contiguous GPUs of a node)."""
    ones = np.flatnonzero(np.asarray(binary_array).reshape(-1))
    return ones.size == 0 or bool(ones[-1] - ones[0] + 1 == ones.size)


def slurm_hostname_key(hostname: str):
    """Sort key that orders `node2` before `node10` (text and number runs alternate)."""
    return [int(p) if p.isdigit() else p for p in re.split(r"(\d+)", hostname) if p != ""]


def check_slurm_availability() -> bool:
    return shutil.which("sbatch") is not None and shutil.which("squeue") is not None and shutil.which("scontrol") is not None
