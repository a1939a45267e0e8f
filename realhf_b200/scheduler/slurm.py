"""Slurm plumbing: node inventory, resource allocation of job steps onto nodes, hostfile / multi-prog / sbatch generation and
job-state parsing.

Parity: `realhf/scheduler/slurm/utils.py` (SlurmResource :59-130, SlurmLaunchInfo :133-355, allocate_resources :357-471,
state parsing :474-822).  The reference packs the job steps of ALL worker types of a trial onto the partition's nodes itself
(instead of leaving placement to Slurm) because model workers must land on known hosts in a known order — rank r of the
process group is task r of the array, and device meshes name hosts explicitly — and writes the placement as an
`--distribution=arbitrary` hostfile next to an `srun --multi-prog` file.  Same here, with B200 node shapes coming from the
cluster spec instead of hard-coded 8-GPU / 80 GB assumptions.  Everything that talks to Slurm goes through `run_cmd`, so the
logic is unit-tested offline against canned `scontrol` / `squeue` output (`tests/test_slurm_cpu.py`).
"""

from __future__ import annotations

import dataclasses
import math
import os
import re
import shlex
import subprocess
from typing import Callable, Dict, List, Optional, Sequence, Tuple


# ------------------------------------------------------------------------------------------------ host lists


def parse_nodelist(expr: Optional[str]) -> List[str]:
    """`node[01-03,07],gpu5` -> [node01, node02, node03, node07, gpu5] (Slurm hostlist syntax, one bracket group per name)."""
    if not expr:
        return []
    out: List[str] = []
    depth, cur = 0, ""
    parts = []
    for ch in expr:
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
            continue
        depth += ch == "["
        depth -= ch == "]"
        cur += ch
    if cur:
        parts.append(cur)
    for p in parts:
        m = re.fullmatch(r"([^\[\]]*)\[([^\]]+)\](.*)", p.strip())
        if not m:
            if p.strip():
                out.append(p.strip())
            continue
        prefix, body, suffix = m.groups()
        for piece in body.split(","):
            if "-" in piece:
                a, b = piece.split("-")
                width = len(a)
                for i in range(int(a), int(b) + 1):
                    out.append(f"{prefix}{i:0{width}d}{suffix}")
            else:
                out.append(f"{prefix}{piece}{suffix}")
    return out


def compress_nodelist(hosts: Sequence[str]) -> str:
    """Inverse of `parse_nodelist` for names of the form <prefix><digits>: [n01, n02, n04] -> n[01-02,04]."""
    groups: Dict[Tuple[str, int], List[int]] = {}
    plain: List[str] = []
    for h in hosts:
        m = re.fullmatch(r"(.*?)(\d+)", h)
        if not m:
            plain.append(h)
            continue
        groups.setdefault((m.group(1), len(m.group(2))), []).append(int(m.group(2)))
    parts = list(plain)
    for (prefix, width), nums in sorted(groups.items()):
        nums = sorted(set(nums))
        runs, start, prev = [], nums[0], nums[0]
        for n in nums[1:] + [None]:
            if n is not None and n == prev + 1:
                prev = n
                continue
            runs.append(f"{start:0{width}d}" if start == prev else f"{start:0{width}d}-{prev:0{width}d}")
            if n is not None:
                start = prev = n
        parts.append(f"{prefix}[{','.join(runs)}]" if (len(runs) > 1 or "-" in runs[0]) else f"{prefix}{runs[0]}")
    return ",".join(parts)


# ------------------------------------------------------------------------------------------------ resources


@dataclasses.dataclass
class SlurmResource:
    """CPU cores, memory (MB) and GPUs of a node — or the demand of one job step."""

    cpu: int = 0
    mem: int = 0
    gpu: int = 0
    gpu_type: Optional[str] = None

    def __add__(self, o: "SlurmResource") -> "SlurmResource":
        return SlurmResource(self.cpu + o.cpu, self.mem + o.mem, self.gpu + o.gpu, self.gpu_type or o.gpu_type)

    def __sub__(self, o: "SlurmResource") -> "SlurmResource":
        return SlurmResource(self.cpu - o.cpu, self.mem - o.mem, self.gpu - o.gpu, self.gpu_type or o.gpu_type)

    def fits(self, demand: "SlurmResource") -> bool:
        if demand.gpu and demand.gpu_type and self.gpu_type and demand.gpu_type.lower() != self.gpu_type.lower():
            return False
        return self.cpu >= demand.cpu and self.mem >= demand.mem and self.gpu >= demand.gpu

    def valid(self) -> bool:
        return self.cpu >= 0 and self.mem >= 0 and self.gpu >= 0


def _gres_gpus(text: str) -> Tuple[int, Optional[str]]:
    """`gpu:b200:8(S:0-1)` / `gpu:8` / `gres/gpu=8,gres/gpu:b200=8` -> (count, type)."""
    best, typ = 0, None
    for m in re.finditer(r"gpu(?::([A-Za-z][\w.-]*))?[:=](\d+)", text or ""):
        n = int(m.group(2))
        if n >= best:
            best, typ = n, (m.group(1) or typ)
    return best, typ


def parse_scontrol_nodes(text: str, usable_states: Sequence[str] = ("IDLE", "MIXED", "ALLOCATED", "COMPLETING")) -> Dict[str, SlurmResource]:
    """FREE resources per node from `scontrol show nodes -o` (one node per line).  Nodes that are down / drained are left out."""
    nodes: Dict[str, SlurmResource] = {}
    for line in text.splitlines():
        kv = dict(m.groups() for m in re.finditer(r"(\w+)=(\S*)", line))
        name = kv.get("NodeName")
        if not name:
            continue
        state = kv.get("State", "").split("+")[0].rstrip("*~#$@")
        flags = kv.get("State", "").upper()
        if state.upper() not in usable_states or any(f in flags for f in ("DRAIN", "DOWN", "FAIL", "MAINT", "NOT_RESPONDING")):
            continue
        cpu = int(kv.get("CPUTot", 0)) - int(kv.get("CPUAlloc", 0))
        mem = int(kv.get("RealMemory", 0)) - int(kv.get("AllocMem", 0))
        total, typ = _gres_gpus(kv.get("Gres", ""))
        used, _ = _gres_gpus(kv.get("AllocTRES", "") or kv.get("GresUsed", ""))
        nodes[name] = SlurmResource(cpu, mem, max(0, total - used), typ)
    return nodes


# ------------------------------------------------------------------------------------------------ launch descriptions


@dataclasses.dataclass
class SlurmLaunchInfo:
    """One worker type of a trial: `n_jobsteps` Slurm tasks, each running `wprocs_per_jobstep` worker processes that share the
    step's resources (the reference packs several CPU-side workers into one task the same way)."""

    run_name: str
    worker_type: str
    cmd: str                      # template with {jobstep_id} {n_jobsteps} {worker_submission_index} {wprocs_per_jobstep} {wprocs_in_job} {wproc_offset}
    wprocs_in_job: int
    resource: SlurmResource       # per job step
    wprocs_per_jobstep: int = 1
    worker_submission_index: int = 0
    wproc_offset: int = 0
    partition: str = "dev"
    nodelist: Optional[str] = None
    exclude: Optional[str] = None
    container_image: Optional[str] = None
    container_mounts: Optional[str] = None
    env_vars: Dict[str, str] = dataclasses.field(default_factory=dict)
    time_limit: Optional[str] = None
    begin: Optional[str] = None
    deadline: Optional[str] = None
    log_dir: str = "."
    hosts: Optional[List[str]] = None   # filled by `allocate_resources`: host of every job step, in task order
    job_id: Optional[str] = None

    @property
    def n_jobsteps(self) -> int:
        return math.ceil(self.wprocs_in_job / self.wprocs_per_jobstep)

    @property
    def slurm_name(self) -> str:
        return f"{self.run_name}:{self.worker_type}" + (f":{self.worker_submission_index}" if self.worker_submission_index else "")

    def _path(self, ext: str) -> str:
        tag = self.worker_type + (f"-{self.worker_submission_index}" if self.worker_submission_index else "")
        return os.path.join(self.log_dir, f"{tag}.{ext}")

    # ---- files
    def multiprog(self) -> str:
        lines = []
        for step in range(self.n_jobsteps):
            first = step * self.wprocs_per_jobstep
            n_here = min(self.wprocs_per_jobstep, self.wprocs_in_job - first)
            cmd = self.cmd.format(jobstep_id=step, n_jobsteps=self.n_jobsteps, worker_submission_index=self.worker_submission_index,
                                  wprocs_per_jobstep=n_here, wprocs_in_job=self.wprocs_in_job, wproc_offset=self.wproc_offset)
            lines.append(f"{step} {cmd}")
        return "\n".join(lines) + "\n"

    def hostfile(self) -> str:
        assert self.hosts is not None and len(self.hosts) == self.n_jobsteps, "allocate_resources() first"
        return "\n".join(self.hosts) + "\n"

    def sbatch_script(self) -> str:
        assert self.hosts is not None, "allocate_resources() first"
        uniq = sorted(set(self.hosts), key=self.hosts.index)
        r = self.resource
        lines = ["#!/bin/bash", f"#SBATCH --job-name={self.slurm_name}", f"#SBATCH --partition={self.partition}",
                 f"#SBATCH --ntasks={self.n_jobsteps}", f"#SBATCH --nodes={len(uniq)}", f"#SBATCH --nodelist={compress_nodelist(uniq)}",
                 f"#SBATCH --cpus-per-task={max(1, r.cpu)}", f"#SBATCH --mem-per-cpu={max(1, r.mem // max(1, r.cpu))}M",
                 "#SBATCH --distribution=arbitrary", f"#SBATCH --output={self._path('out')}", "#SBATCH --open-mode=append"]
        if r.gpu:
            lines.append(f"#SBATCH --gpus-per-task={(r.gpu_type + ':') if r.gpu_type else ''}{r.gpu}")
        for flag, v in (("time", self.time_limit), ("begin", self.begin), ("deadline", self.deadline)):
            if v:
                lines.append(f"#SBATCH --{flag}={v}")
        lines.append(f"export SLURM_HOSTFILE={shlex.quote(self._path('hostfile'))}")
        for k, v in self.env_vars.items():
            lines.append(f"export {k}={shlex.quote(str(v))}")
        container = ""
        if self.container_image:
            container = f"--container-image={self.container_image} "
            if self.container_mounts:
                container += f"--container-mounts={self.container_mounts} "
        lines.append(f"srun -K -l {container}--ntasks={self.n_jobsteps} --distribution=arbitrary --multi-prog {shlex.quote(self._path('multiprog'))}")
        return "\n".join(lines) + "\n"

    def commit(self) -> str:
        """Write hostfile, multi-prog file and sbatch script; returns the script path."""
        os.makedirs(self.log_dir, exist_ok=True)
        for ext, body in (("hostfile", self.hostfile()), ("multiprog", self.multiprog()), ("sbatch", self.sbatch_script())):
            with open(self._path(ext), "w") as f:
                f.write(body)
        return self._path("sbatch")


class SlurmResourceNotEnoughException(Exception):
    pass


def allocate_resources(infos: List[SlurmLaunchInfo], nodes: Dict[str, SlurmResource], strategy: str = "pack") -> List[SlurmLaunchInfo]:
    """Place every job step of every launch on a node (sets `info.hosts`).

    GPU-demanding launches go first, larger demands before smaller ones (first-fit decreasing).  `pack` fills a node before
    moving to the next one, so the consecutive tasks of a worker array — consecutive ranks of the process group, hence the
    ranks of one TP / DP group — share an NVSwitch domain; `spread` round-robins over the eligible nodes (CPU-side workers).
    Respects per-launch `nodelist` / `exclude`.  Raises SlurmResourceNotEnoughException with the unmet demand."""
    free = {k: dataclasses.replace(v) for k, v in nodes.items()}
    order = sorted(infos, key=lambda i: (-i.resource.gpu, -i.resource.cpu * max(1, i.n_jobsteps), i.worker_type))
    for info in order:
        allowed = [h for h in (parse_nodelist(info.nodelist) or sorted(free)) if h in free and h not in set(parse_nodelist(info.exclude))]
        hosts: List[str] = []
        rr = 0
        for step in range(info.n_jobsteps):
            cand = allowed if strategy == "pack" or info.resource.gpu else allowed[rr:] + allowed[:rr]
            host = next((h for h in cand if free[h].fits(info.resource)), None)
            if host is None:
                have = {h: dataclasses.asdict(free[h]) for h in allowed}
                raise SlurmResourceNotEnoughException(
                    f"{info.slurm_name}: job step {step}/{info.n_jobsteps} needs {dataclasses.asdict(info.resource)}, free on the eligible nodes: {have}")
            free[host] = free[host] - info.resource
            hosts.append(host)
            rr = (allowed.index(host) + 1) % max(1, len(allowed))
        info.hosts = hosts
    return infos


# ------------------------------------------------------------------------------------------------ talking to slurm


RunCmd = Callable[[List[str]], str]


def run_cmd(argv: List[str]) -> str:
    return subprocess.run(argv, capture_output=True, text=True, check=True).stdout


def query_nodes(partition: Optional[str] = None, run: RunCmd = run_cmd) -> Dict[str, SlurmResource]:
    nodes = parse_scontrol_nodes(run(["scontrol", "show", "nodes", "-o"]))
    if partition:
        try:
            in_part = set(parse_nodelist(",".join(run(["sinfo", "-h", "-p", partition, "-o", "%N"]).split())))
            nodes = {k: v for k, v in nodes.items() if k in in_part}
        except (subprocess.CalledProcessError, OSError):
            pass
    return nodes


_STATES = {"PENDING": "PENDING", "CONFIGURING": "PENDING", "RUNNING": "RUNNING", "COMPLETING": "RUNNING", "COMPLETED": "COMPLETED",
           "CANCELLED": "CANCELLED", "FAILED": "FAILED", "TIMEOUT": "FAILED", "NODE_FAIL": "FAILED", "OUT_OF_MEMORY": "FAILED",
           "PREEMPTED": "CANCELLED", "BOOT_FAIL": "FAILED", "DEADLINE": "FAILED", "SUSPENDED": "PENDING"}


def parse_job_states(text: str) -> Dict[str, Tuple[str, str, str]]:
    """`squeue -h -o "%i|%T|%j|%N"` / `sacct -n -P -o JobID,State,JobName,NodeList` lines -> {job id: (state, name, nodes)};
    job steps (`123.0`, `123.batch`) are folded into their job, a failed step fails the job."""
    out: Dict[str, Tuple[str, str, str]] = {}
    for line in text.splitlines():
        f = [x.strip() for x in line.split("|")]
        if len(f) < 2 or not f[0]:
            continue
        jid = f[0].split(".")[0].split("_")[0]
        st = _STATES.get(f[1].split()[0].upper(), "FAILED")
        name = f[2] if len(f) > 2 else ""
        nodes = f[3] if len(f) > 3 else ""
        if jid in out:
            prev = out[jid]
            worse = st if st in ("FAILED", "CANCELLED") else prev[0]
            out[jid] = (worse, prev[1] or name, prev[2] or nodes)
        else:
            out[jid] = (st, name, nodes)
    return out


def job_states(job_ids: Sequence[str], run: RunCmd = run_cmd) -> Dict[str, Tuple[str, str, str]]:
    """Live state from squeue; jobs that left the queue are looked up in the accounting database (COMPLETED if unknown there)."""
    if not job_ids:
        return {}
    ids = ",".join(job_ids)
    try:
        live = parse_job_states(run(["squeue", "-h", "-j", ids, "-o", "%i|%T|%j|%N"]))
    except (subprocess.CalledProcessError, OSError):
        live = {}
    missing = [j for j in job_ids if j not in live]
    if missing:
        try:
            done = parse_job_states(run(["sacct", "-n", "-P", "-j", ",".join(missing), "-o", "JobID,State,JobName,NodeList"]))
        except (subprocess.CalledProcessError, OSError):
            done = {}
        for j in missing:
            live[j] = done.get(j, ("COMPLETED", "", ""))
    return live
