"""Ray mode: the worker processes of a trial are Ray tasks instead of OS processes started by this shell or by Slurm.

Parity: `realhf/system/controller.py:348-575` (`run_ray_worker`, `RayController._launch_workers`: resource check against
`ray.available_resources()`, GPU workers packed onto nodes in rank order through the per-node custom resource
`node:<ip>`, CPU workers left to Ray's default placement, a first `ray.get(timeout)` poll so that import errors of the
remote side surface at launch) and `scheduler/client.py:145` (`mode="ray"`).

The reference needs a second controller class for Ray because its controller *is* the process launcher.  Here the launcher
talks to a `SchedulerClient`, so Ray is one more client: `submit_array` turns every worker into one remote task that runs
the same `apps.remote.main_worker` a local / Slurm worker process runs, `find_all` folds `ray.wait` into job states, and
everything above it (`apps/main.py::Controller`, status leases, the control panel, recover) is unchanged.

Placement.  A B200 node is one NVSwitch domain and the peer-memory paths (direct-store reallocation, NVLS ZeRO, fused TP
kernels) need the ranks of a group on ONE node with every GPU of the node visible in each process.  GPU workers are
therefore packed in rank order, `gpus_per_node` consecutive ranks per node; each task reserves one GPU for Ray's accounting
(`num_gpus=1`) but keeps the whole node visible and takes its own device from `REAL_LOCAL_GPU` (`REAL_ISOLATE_GPUS=1`
restores Ray's one-visible-GPU behaviour).

`ray` is imported lazily: the package does not depend on it.
"""

from __future__ import annotations

import os
import re
import time
from typing import Dict, List, Optional

from realhf_b200.base import logging
from realhf_b200.scheduler.client import JobException, JobInfo, JobState, SchedulerClient
from realhf_b200.scheduler.ray_entry import run_ray_worker

logger = logging.getLogger("scheduler.ray")

_NODE_RESOURCE = re.compile(r"node:(\b(?:\d{1,3}\.){3}\d{1,3}\b)")


def _import_ray():
    try:
        import ray  # noqa: WPS433
    except ImportError as e:  # pragma: no cover - depends on the environment
        raise RuntimeError("mode=ray needs the `ray` package on the launcher and on every node "
                           "(`pip install ray`); use mode=local or mode=slurm otherwise") from e
    return ray


def required_resources(requests: List[dict]) -> Dict[str, float]:
    """Sum of what a list of `submit_array` calls asks for: {"CPU", "GPU", "memory" (GiB)}."""
    cpu = sum(r["cpu"] * r["count"] for r in requests)
    gpu = sum(r["gpu"] * r["count"] for r in requests)
    mem = sum(r["mem"] * r["count"] for r in requests) / 1024.0
    return {"CPU": float(cpu), "GPU": float(gpu), "memory": mem}


def pack_gpu_workers(count: int, node_resources: List[str], total_gpus: float) -> List[tuple]:
    """[(node resource name, slot on that node)] for GPU workers 0..count-1: `gpus_per_node` consecutive ranks per node, in
    the order Ray lists the nodes (reference: controller.py:463-505; heterogeneous GPU counts are rejected there too)."""
    if not node_resources:
        raise RuntimeError("Ray reports no `node:<ip>` resources: is the cluster up?")
    if total_gpus % len(node_resources) != 0:
        raise ValueError("cannot place GPU workers on nodes with different numbers of GPUs")
    per_node = int(total_gpus // len(node_resources))
    if per_node == 0 or total_gpus < count:
        raise RuntimeError(f"{count} GPU workers requested but Ray has {int(total_gpus)} GPUs")
    return [(node_resources[i // per_node], i % per_node) for i in range(count)]


class RaySchedulerClient(SchedulerClient):
    def __init__(self, expr_name: str, trial_name: str, address: Optional[str] = None, **init_kwargs):
        super().__init__(expr_name, trial_name)
        self._ray = _import_ray()
        if not self._ray.is_initialized():
            # RAY_ADDRESS / address="auto" joins a running cluster; nothing given starts a local instance (single node)
            addr = address or os.environ.get("RAY_ADDRESS")
            self._ray.init(**(dict(address=addr) if addr else {}), **init_kwargs)
            self._own_ray = True
        else:
            self._own_ray = False
        self._refs: Dict[str, object] = {}
        self._requests: List[dict] = []
        self._final: Dict[str, JobInfo] = {}

    # ---- submission
    def submit_array(self, worker_type: str, cmd: str, count: int, cpu: int = 1, gpu: int = 0, mem: int = 1024,
                     env_vars: Optional[Dict[str, str]] = None, **kw):
        """`cmd` (the `apps.remote worker ...` command line of the process-based clients) is not executed: a Ray task calls
        the same entry point in-process.  `mem` is in MiB like everywhere else."""
        ray = self._ray
        req = dict(worker_type=worker_type, count=count, cpu=cpu, gpu=gpu, mem=mem)
        self._requests.append(req)
        avail = ray.available_resources()
        need = required_resources([req])
        have = {"CPU": avail.get("CPU", 0.0), "GPU": avail.get("GPU", 0.0), "memory": avail.get("memory", 0.0) / 1024 ** 3}
        short = [f"{k}: need {need[k]:g}, available {have[k]:g}" for k in need if have[k] < need[k]]
        if short:  # Ray would queue the tasks forever: say so now (reference: controller.py:441-447 logs the same condition)
            logger.critical(f"Ray does not have the resources for {count} x {worker_type} ({'; '.join(short)}); "
                            "the experiment will wait until more nodes join")
        env = dict(env_vars or {})
        env.setdefault("OMP_NUM_THREADS", str(max(1, int(cpu))))
        if gpu > 0:
            if gpu != 1:
                raise ValueError("Ray mode places exactly one GPU per model worker")
            nodes = sorted(k for k in avail if _NODE_RESOURCE.fullmatch(k))
            placement = pack_gpu_workers(count, nodes, avail.get("GPU", 0.0))
            per_node = int(avail["GPU"] // len(nodes))
        refs = []
        for i in range(count):
            opts = dict(num_cpus=cpu, num_gpus=gpu, memory=int(mem) * 1024 ** 2, name=f"{worker_type}/{i}",
                        runtime_env=dict(env_vars={k: str(v) for k, v in env.items()}), max_retries=0)
            slot = None
            if gpu > 0:
                node, slot = placement[i]
                opts["resources"] = {node: 1.0 / per_node}
            ref = ray.remote(**opts)(run_ray_worker).remote(worker_type, i, count, self.expr_name, self.trial_name, env, slot)
            if gpu > 0:
                # one by one, so that the ranks of a node grab its GPUs in rank order
                self._poll_once([ref], 0.1)
            refs.append(ref)
            self._refs[f"{worker_type}/{i}"] = ref
        self._poll_once(refs, 1.0)   # let import errors / missing packages on the remote side raise here
        logger.info(f"launched {count} x {worker_type} as Ray tasks")

    def _poll_once(self, refs, timeout):
        try:
            self._ray.get(refs, timeout=timeout)
        except self._ray.exceptions.GetTimeoutError:
            pass

    # ---- queries
    def find(self, job_name: str) -> Optional[JobInfo]:
        if job_name in self._final:
            return self._final[job_name]
        ref = self._refs.get(job_name)
        if ref is None:
            return JobInfo(job_name, JobState.NOT_FOUND)
        ready, _ = self._ray.wait([ref], timeout=0)
        if not ready:
            return JobInfo(job_name, JobState.RUNNING, host="ray")
        try:
            self._ray.get(ref)
            info = JobInfo(job_name, JobState.COMPLETED, host="ray", returncode=0)
        except self._ray.exceptions.TaskCancelledError:
            info = JobInfo(job_name, JobState.CANCELLED, host="ray")
        except Exception as e:
            logger.error(f"{job_name} failed: {type(e).__name__}: {e}")
            info = JobInfo(job_name, JobState.FAILED, host="ray", returncode=1)
        self._final[job_name] = info
        return info

    def find_all(self, job_name_regex: str = ".*") -> List[JobInfo]:
        return [self.find(n) for n in list(self._refs) if re.fullmatch(job_name_regex, n)]

    def wait(self, timeout=None, check_status=(JobState.FAILED, JobState.CANCELLED, JobState.NOT_FOUND),
             remove_status=(JobState.COMPLETED,), update=False, poll=0.5):
        t0 = time.monotonic()
        left = set(self._refs)
        while left:
            for n in list(left):
                info = self.find(n)
                if info.state in check_status:
                    raise JobException(self.run_name, n, "ray", info.state)
                if info.state in remove_status:
                    left.discard(n)
            if timeout is not None and time.monotonic() - t0 > timeout:
                raise TimeoutError(f"Ray tasks still running after {timeout}s: {sorted(left)}")
            if left:
                time.sleep(poll)

    def stop_all(self, signal_=None):
        for n, ref in list(self._refs.items()):
            if n in self._final:
                continue
            try:
                self._ray.cancel(ref, force=True)
            except Exception:
                pass
        self._refs.clear()
        self._final.clear()
        if self._own_ray:
            try:
                self._ray.shutdown()
            except Exception:
                pass
            self._own_ray = False
