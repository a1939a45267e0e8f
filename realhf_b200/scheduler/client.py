"""Scheduler clients: launch / watch / stop the worker processes of a trial.

Parity: `realhf/scheduler/client.py` (SchedulerClient API :44-113, make :145), `scheduler/local/client.py`
(subprocess + psutil) and `scheduler/slurm/*` (sbatch / srun command construction).
"""

from __future__ import annotations

import dataclasses
import enum
import os
import shlex
import signal
import subprocess
import sys
import time
from typing import Dict, List, Optional

import psutil

from realhf_b200.base import constants, logging

logger = logging.getLogger("scheduler")


class JobState(enum.Enum):
    NOT_FOUND = 0
    PENDING = 1
    RUNNING = 2
    COMPLETED = 3
    FAILED = 4
    CANCELLED = 5

    def active(self):
        return self in (JobState.PENDING, JobState.RUNNING)


class JobException(Exception):
    def __init__(self, run_name, worker_type, host, reason: JobState):
        super().__init__(f"job {run_name}:{worker_type} {reason} at {host}")
        self.run_name, self.worker_type, self.host, self.reason = run_name, worker_type, host, reason


@dataclasses.dataclass
class JobInfo:
    name: str
    state: JobState
    host: Optional[str] = None
    pid: Optional[int] = None
    returncode: Optional[int] = None


class SchedulerClient:
    def __init__(self, expr_name: str, trial_name: str):
        self.expr_name, self.trial_name = expr_name, trial_name
        self.run_name = f"{expr_name}_{trial_name}"

    def submit(self, worker_type: str, cmd: str, **kw):
        return self.submit_array(worker_type, cmd, count=1, **kw)

    def submit_array(self, worker_type: str, cmd: str, count: int, **kw):
        raise NotImplementedError()

    def stop_all(self, signal_=None):
        raise NotImplementedError()

    def find(self, job_name: str) -> Optional[JobInfo]:
        raise NotImplementedError()

    def find_all(self, job_name_regex: str = ".*") -> List[JobInfo]:
        raise NotImplementedError()

    def wait(self, timeout=None, **kw):
        raise NotImplementedError()


def remote_worker_cmd(expr_name: str, trial_name: str, debug: bool, worker_type: str) -> str:
    flags = "" if debug else "-O "
    return (f"{sys.executable} {flags}-m realhf_b200.apps.remote worker -w {worker_type} -e {expr_name} -f {trial_name} "
            "-i {jobstep_id} -g {n_jobsteps} -r {worker_submission_index} -p {wprocs_per_jobstep} -j {wprocs_in_job} -o {wproc_offset}")


class LocalSchedulerClient(SchedulerClient):
    """One OS process per worker.  Model worker i drives GPU i (`REAL_LOCAL_GPU`) but keeps every GPU of the node visible:
    peer-mapped symmetric memory (CUDA IPC: custom all-reduce, fused TP GEMMs, direct-store realloc) needs the peers' devices
    in the process.  `REAL_ISOLATE_GPUS=1` restores the reference's one-visible-GPU-per-worker behaviour."""

    def __init__(self, expr_name, trial_name):
        super().__init__(expr_name, trial_name)
        self._procs: Dict[str, subprocess.Popen] = {}
        self._logs: Dict[str, str] = {}

    def submit_array(self, worker_type: str, cmd: str, count: int, gpu: int = 0, env_vars: Optional[Dict[str, str]] = None, **kw):
        log_dir = constants.run_dirs(self.expr_name, self.trial_name)["log"]
        n_gpus = int(os.environ.get("REAL_N_VISIBLE_GPUS", "0")) or _count_gpus()
        for i in range(count):
            env = dict(os.environ)
            env.update(env_vars or {})
            # co-located workers must not each spin up one OpenMP thread per core (the reference forwards OMP_NUM_THREADS too)
            env.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or 8) // (count + 1))))
            if gpu > 0 and n_gpus > 0:
                if os.environ.get("REAL_ISOLATE_GPUS", "0") == "1":
                    env["CUDA_VISIBLE_DEVICES"] = str(i % n_gpus)
                    env["REAL_LOCAL_GPU"] = "0"
                else:
                    env["REAL_LOCAL_GPU"] = str(i % n_gpus)
            c = cmd.format(jobstep_id=i, n_jobsteps=count, worker_submission_index=0, wprocs_per_jobstep=1, wprocs_in_job=count,
                           wproc_offset=0)
            name = f"{worker_type}/{i}"
            log = os.path.join(log_dir, f"{worker_type}-{i}")
            self._logs[name] = log
            f = open(log, "a")
            self._procs[name] = subprocess.Popen(shlex.split(c), env=env, stdout=f, stderr=subprocess.STDOUT, start_new_session=True)
            logger.info(f"started {name} (pid {self._procs[name].pid}), log {log}")

    def find(self, job_name: str) -> Optional[JobInfo]:
        p = self._procs.get(job_name)
        if p is None:
            return JobInfo(job_name, JobState.NOT_FOUND)
        rc = p.poll()
        st = JobState.RUNNING if rc is None else (JobState.COMPLETED if rc == 0 else JobState.FAILED)
        return JobInfo(job_name, st, "localhost", p.pid, rc)

    def find_all(self, job_name_regex: str = ".*"):
        import re
        return [self.find(n) for n in self._procs if re.fullmatch(job_name_regex, n)]

    def wait(self, timeout=None, check_status=(JobState.FAILED, JobState.CANCELLED, JobState.NOT_FOUND),
             remove_status=(JobState.COMPLETED,), update=False, poll=0.5):
        t0 = time.monotonic()
        left = set(self._procs)
        while left:
            for n in list(left):
                info = self.find(n)
                if info.state in check_status:
                    raise JobException(self.run_name, n, "localhost", info.state)
                if info.state in remove_status:
                    left.discard(n)
            if timeout is not None and time.monotonic() - t0 > timeout:
                raise TimeoutError(f"jobs still running after {timeout}s: {sorted(left)}")
            time.sleep(poll)

    def stop_all(self, signal_=signal.SIGTERM):
        for n, p in self._procs.items():
            if p.poll() is None:
                try:
                    parent = psutil.Process(p.pid)
                    for ch in parent.children(recursive=True):
                        ch.send_signal(signal_)
                    parent.send_signal(signal_)
                except psutil.NoSuchProcess:
                    pass
        t0 = time.monotonic()
        for p in self._procs.values():
            try:
                p.wait(timeout=max(0.1, 20 - (time.monotonic() - t0)))
            except subprocess.TimeoutExpired:
                p.kill()
        self._procs.clear()

    def log_tail(self, job_name: str, n: int = 40) -> str:
        try:
            with open(self._logs[job_name]) as f:
                return "".join(f.readlines()[-n:])
        except (KeyError, OSError):
            return ""


def _count_gpus() -> int:
    from realhf_b200.base.gpu_utils import gpu_count
    return gpu_count()


class SlurmSchedulerClient(SchedulerClient):
    """Slurm mode (reference: scheduler/slurm/client.py + utils.py).  `submit_array` only records a launch; `commit()` (called by
    `wait()` or explicitly) queries the free resources of the partition, places the job steps of ALL worker types together
    (`scheduler.slurm.allocate_resources`: GPU workers packed in rank order, so the ranks of a TP / DP group share an NVSwitch
    domain), writes one hostfile + multi-prog file + sbatch script per worker type and submits them.  Every Slurm command goes
    through `run`, so the whole flow is tested offline with canned command output."""

    def __init__(self, expr_name, trial_name, partition: Optional[str] = None, container_image: Optional[str] = None,
                 container_mounts: Optional[str] = None, run=None):
        super().__init__(expr_name, trial_name)
        from realhf_b200.base import cluster
        from realhf_b200.scheduler import slurm
        cs = cluster.spec()  # partition / images / mounts default to the cluster spec ($CLUSTER_SPEC_PATH)
        self.partition = partition or cs.partition or "dev"
        self.image, self.cpu_image = container_image or cs.gpu_image, container_image or cs.cpu_image
        self.mounts = container_mounts or cs.default_mount or None
        self.gpu_type = os.environ.get("REAL_SLURM_GPU_TYPE") or None  # e.g. "b200": only request typed GPUs when asked to
        self._run = run or slurm.run_cmd
        self._pending: List = []
        self._launched: Dict[str, object] = {}

    def submit_array(self, worker_type, cmd, count, cpu: int = 4, gpu: int = 0, mem: int = 10000, nodelist: Optional[str] = None,
                     exclude: Optional[str] = None, time_limit: Optional[str] = None, env_vars: Optional[Dict[str, str]] = None,
                     container_image: Optional[str] = None, wprocs_per_jobstep: int = 1, begin: Optional[str] = None,
                     deadline: Optional[str] = None, **_ignored):
        from realhf_b200.scheduler import slurm
        n_prev = sum(1 for i in self._pending + list(self._launched.values()) if i.worker_type == worker_type)
        offset = sum(i.wprocs_in_job for i in self._pending + list(self._launched.values()) if i.worker_type == worker_type)
        info = slurm.SlurmLaunchInfo(
            run_name=self.run_name, worker_type=worker_type, cmd=cmd, wprocs_in_job=count, wprocs_per_jobstep=wprocs_per_jobstep,
            worker_submission_index=n_prev, wproc_offset=offset,
            resource=slurm.SlurmResource(cpu=cpu, mem=mem, gpu=gpu, gpu_type=self.gpu_type if gpu else None), partition=self.partition,
            nodelist=nodelist, exclude=exclude, container_image=container_image or (self.image if gpu else self.cpu_image),
            container_mounts=self.mounts, env_vars=dict(env_vars or {}), time_limit=time_limit, begin=begin, deadline=deadline,
            log_dir=constants.run_dirs(self.expr_name, self.trial_name)["log"])
        self._pending.append(info)

    def build_script(self, worker_type: str, cmd: str, count: int, **kw) -> str:
        """The sbatch script `submit_array` + `commit` would produce for one worker type placed on its own (for inspection)."""
        from realhf_b200.scheduler import slurm
        saved, self._pending = self._pending, []
        try:
            self.submit_array(worker_type, cmd, count, **kw)
            info = self._pending[0]
        finally:
            self._pending = saved
        try:
            nodes = slurm.query_nodes(self.partition, self._run)
        except (OSError, subprocess.CalledProcessError):
            # no Slurm on this machine (dry run / tests): an idle partition shaped by the cluster spec
            from realhf_b200.base import cluster
            cs = cluster.spec()
            per = max(1, cs.n_gpus_per_node)
            names = slurm.parse_nodelist(info.nodelist) or cs.node_names(list(range(1, 2 + (count * max(1, info.resource.gpu)) // per)))
            nodes = {n: slurm.SlurmResource(cpu=1 << 20, mem=1 << 40, gpu=per, gpu_type=info.resource.gpu_type) for n in names}
        slurm.allocate_resources([info], nodes)
        info.commit()  # the script refers to its hostfile / multi-prog file
        return info.sbatch_script()

    def commit(self):
        from realhf_b200.scheduler import slurm
        if not self._pending:
            return
        slurm.allocate_resources(self._pending, slurm.query_nodes(self.partition, self._run))
        for info in self._pending:
            path = info.commit()
            out = self._run(["sbatch", "--parsable", path]).strip()
            info.job_id = out.split(";")[0]
            key = info.worker_type if info.worker_submission_index == 0 else f"{info.worker_type}:{info.worker_submission_index}"
            self._launched[key] = info
            logger.info(f"submitted {info.slurm_name} as slurm job {info.job_id} on {sorted(set(info.hosts))}")
        self._pending = []

    def find_all(self, job_name_regex=".*"):
        import re

        from realhf_b200.scheduler import slurm
        self.commit()
        ids = [i.job_id for i in self._launched.values()]
        st = slurm.job_states(ids, self._run)
        m = {"PENDING": JobState.PENDING, "RUNNING": JobState.RUNNING, "COMPLETED": JobState.COMPLETED, "FAILED": JobState.FAILED,
             "CANCELLED": JobState.CANCELLED}
        return [JobInfo(k, m[st[i.job_id][0]], host=st[i.job_id][2] or None) for k, i in self._launched.items() if re.fullmatch(job_name_regex, k)]

    def find(self, job_name: str):
        return next(iter(self.find_all(job_name)), None)

    def wait(self, timeout=None, poll=10, **kw):
        self.commit()
        t0 = time.monotonic()
        while True:
            infos = self.find_all()
            bad = next((i for i in infos if i.state in (JobState.FAILED, JobState.CANCELLED)), None)
            if bad is not None:
                raise JobException(self.run_name, bad.name, bad.host or "slurm", bad.state)
            if all(i.state == JobState.COMPLETED for i in infos):
                return
            if timeout is not None and time.monotonic() - t0 > timeout:
                raise TimeoutError()
            time.sleep(poll)

    def stop_all(self, signal_=None):
        sig = ["-s", "INT"] if signal_ == signal.SIGINT else []
        self._pending = []
        if self._launched:
            for info in self._launched.values():
                self._run(["scancel"] + sig + [info.job_id])
        else:  # a fresh client (`apps.main stop` from another shell): address the trial's jobs by name
            for wt in ("master_worker", "model_worker"):
                try:
                    self._run(["scancel"] + sig + ["--name", f"{self.run_name}:{wt}"])
                except Exception:
                    pass


def make(mode: str, expr_name: str, trial_name: str, **kw) -> SchedulerClient:
    if mode == "local":
        return LocalSchedulerClient(expr_name, trial_name)
    if mode == "slurm":
        return SlurmSchedulerClient(expr_name, trial_name, **kw)
    if mode == "ray":
        from realhf_b200.scheduler.ray import RaySchedulerClient
        return RaySchedulerClient(expr_name, trial_name, **kw)
    raise NotImplementedError(mode)
