"""Which GPU a worker process uses.  Parity: `realhf/base/gpu_utils.py` (gpu_count / set_cuda_device / isolate_cuda_device /
reveal_pg_identity).

The reference isolates every worker with `CUDA_VISIBLE_DEVICES=<one id>`.  On an NVSwitch node that is the wrong default: the
peer-memory paths (direct-store reallocation, fused TP kernels, the NVLS optimizer) need the peers' devices in the process.  So
the default here is "all GPUs visible, `torch.cuda.set_device(local id)`"; `REAL_ISOLATE_GPUS=1` restores the reference's
behaviour.  The local id comes from the scheduler when it binds GPUs itself (`REAL_LOCAL_GPU`, or one visible device under
`--gpus-per-task=1`), and otherwise from a rendezvous of the workers of one host through name_resolve (`local_gpu_index`)."""

from __future__ import annotations

import os
import socket
import subprocess
from typing import List, Optional

from realhf_b200.base import name_resolve


def gpu_count() -> int:
    """GPUs this process may use: entries of CUDA_VISIBLE_DEVICES if set, else what `nvidia-smi -L` lists (0 without a driver)."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis is not None:
        return len([x for x in vis.split(",") if x.strip() != ""])
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=10).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


def set_cuda_device(device) -> None:
    import torch
    if device is not None and str(device) != "cpu" and torch.cuda.is_available():
        torch.cuda.set_device(device)


def _identity_key(exp: str, trial: str, worker_type: str, index: int) -> str:
    return f"{exp}/{trial}/gpu_identity/{worker_type}/{index}"


def reveal_identity(exp: str, trial: str, worker_type: str, index: int, host: Optional[str] = None) -> None:
    """Publish which host this worker runs on (the other half of `local_gpu_index`)."""
    name_resolve.add(_identity_key(exp, trial, worker_type, index), host or socket.gethostname(), replace=True)


def local_gpu_index(exp: str, trial: str, worker_type: str, index: int, world: int, host: Optional[str] = None,
                    n_gpus: Optional[int] = None, timeout: float = 300.0) -> int:
    """Local GPU id of worker `index` of `world`: its position among the workers that published the SAME host name, in worker
    order.  Every worker calls this (it publishes its own identity first, then waits for all peers).  Raises when a host got
    more workers than it has GPUs."""
    host = host or socket.gethostname()
    reveal_identity(exp, trial, worker_type, index, host)
    hosts: List[str] = [name_resolve.wait(_identity_key(exp, trial, worker_type, i), timeout=timeout) for i in range(world)]
    mates = [i for i in range(world) if hosts[i] == host]
    n = gpu_count() if n_gpus is None else n_gpus
    if n and len(mates) > n:
        raise RuntimeError(f"host {host} runs {len(mates)} {worker_type}s but has {n} GPUs")
    return mates.index(index)


def isolate_cuda_device(exp: str, trial: str, worker_type: str, index: int, world: int, **kw) -> int:
    """Resolve the local GPU id and make it the process's device: with `REAL_ISOLATE_GPUS=1` by narrowing CUDA_VISIBLE_DEVICES to
    it (must run before CUDA is initialised; the device is then cuda:0), else by returning it for `torch.cuda.set_device`."""
    if "REAL_LOCAL_GPU" in os.environ:
        local = int(os.environ["REAL_LOCAL_GPU"])
    else:
        local = local_gpu_index(exp, trial, worker_type, index, world, **kw)
    if os.environ.get("REAL_ISOLATE_GPUS", "0") == "1":
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        ids = [x.strip() for x in vis.split(",")] if vis else [str(i) for i in range(max(gpu_count(), local + 1))]
        os.environ["CUDA_VISIBLE_DEVICES"] = ids[local]
        os.environ["REAL_LOCAL_GPU"] = "0"
        return 0
    return local
