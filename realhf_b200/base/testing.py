"""Multi-process test harness (parity: `realhf/base/testing.py` LocalMultiProcessTest): spawn `world_size` processes on
this host, initialise a gloo (CPU) or nccl (one GPU per rank) process group, run a function, collect per-rank results;
the first failure kills everyone."""

from __future__ import annotations

import os
import queue
import socket
import time
import traceback
from typing import Any, Callable, Dict, List

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(rank: int, world_size: int, backend: str, port: int, fn: Callable, kwargs: Dict, q):
    try:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        torch.set_num_threads(max(1, (os.cpu_count() or 4) // world_size))
        if backend == "nccl":
            torch.cuda.set_device(rank)
            dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=torch.device("cuda", rank))
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world_size)
        res = fn(rank, world_size, **kwargs)
        dist.barrier()
        q.put((rank, "ok", res))
    except Exception:
        q.put((rank, "err", traceback.format_exc()))
    finally:
        try:
            dist.destroy_process_group()
        except Exception:
            pass


def run_distributed(fn: Callable, world_size: int, backend: str = "gloo", timeout: float = 600.0, **kwargs) -> List[Any]:
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_entry, args=(r, world_size, backend, port, fn, kwargs, q)) for r in range(world_size)]
    for p in procs:
        p.start()
    results: Dict[int, Any] = {}
    t0 = time.monotonic()
    err = None
    while len(results) < world_size and err is None:
        try:
            rank, status, payload = q.get(timeout=1.0)
            if status == "err":
                err = f"rank {rank} failed:\n{payload}"
            else:
                results[rank] = payload
        except queue.Empty:
            if time.monotonic() - t0 > timeout:
                err = f"timeout after {timeout}s (finished ranks: {sorted(results)})"
            elif any(p.exitcode not in (None, 0) for p in procs):
                err = f"a worker process died: exit codes {[p.exitcode for p in procs]}"
    for p in procs:
        if err is not None and p.is_alive():
            p.terminate()
        p.join(timeout=10)
    if err is not None:
        raise RuntimeError(err)
    return [results[r] for r in range(world_size)]
