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
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=10).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


class SlurmSchedulerClient(SchedulerClient):
    """Builds one sbatch script per worker type with `srun --multi-prog` (reference: scheduler/slurm/utils.py:357-471).
    Submission needs a Slurm cluster; command construction is unit-testable offline."""

    def __init__(self, expr_name, trial_name, partition: Optional[str] = None, container_image: Optional[str] = None,
                 container_mounts: Optional[str] = None):
        super().__init__(expr_name, trial_name)
        from realhf_b200.base import cluster
        cs = cluster.spec()  # partition / images / mounts default to the cluster spec ($CLUSTER_SPEC_PATH)
        self.partition = partition or cs.partition or "dev"
        self.image, self.cpu_image = container_image or cs.gpu_image, container_image or cs.cpu_image
        self.mounts = container_mounts or cs.default_mount or "/:/host"
        self._job_ids: Dict[str, str] = {}

    def build_script(self, worker_type: str, cmd: str, count: int, cpu: int = 4, gpu: int = 0, mem: int = 10000,
                     nodelist: Optional[str] = None, exclude: Optional[str] = None, time_limit: Optional[str] = None,
                     env_vars: Optional[Dict[str, str]] = None, gpus_per_node: Optional[int] = None,
                     container_image: Optional[str] = None) -> str:
        if gpus_per_node is None:
            from realhf_b200.base import cluster
            gpus_per_node = cluster.spec().n_gpus_per_node
        log_dir = constants.run_dirs(self.expr_name, self.trial_name)["log"]
        n_nodes = max(1, (count * max(gpu, 0) + gpus_per_node - 1) // gpus_per_node) if gpu else 1
        lines = ["#!/bin/bash", f"#SBATCH --job-name={self.run_name}:{worker_type}", f"#SBATCH --partition={self.partition}",
                 f"#SBATCH --ntasks={count}", f"#SBATCH --nodes={n_nodes}", f"#SBATCH --cpus-per-task={cpu}", f"#SBATCH --mem-per-cpu={max(1, mem // max(cpu, 1))}M",
                 f"#SBATCH --output={log_dir}/{worker_type}-%t.out", "#SBATCH --open-mode=append"]
        if gpu:
            lines.append(f"#SBATCH --gpus-per-task={gpu}")
        if nodelist:
            lines.append(f"#SBATCH --nodelist={nodelist}")
        if exclude:
            lines.append(f"#SBATCH --exclude={exclude}")
        if time_limit:
            lines.append(f"#SBATCH --time={time_limit}")
        for k, v in (env_vars or {}).items():
            lines.append(f"export {k}={shlex.quote(str(v))}")
        multiprog = os.path.join(log_dir, f"{worker_type}.multiprog")
        with open(multiprog, "w") as f:
            for i in range(count):
                f.write(f"{i} " + cmd.format(jobstep_id=i, n_jobsteps=count, worker_submission_index=0, wprocs_per_jobstep=1,
                                             wprocs_in_job=count, wproc_offset=0) + "\n")
        image = container_image or (self.image if gpu else self.cpu_image)
        container = f"--container-image={image} --container-mounts={self.mounts} " if image else ""
        lines.append(f"srun {container}--multi-prog {multiprog}")
        return "\n".join(lines) + "\n"

    def submit_array(self, worker_type, cmd, count, **kw):
        allowed = ("cpu", "gpu", "mem", "nodelist", "exclude", "time_limit", "env_vars", "gpus_per_node", "container_image")
        script = self.build_script(worker_type, cmd, count, **{k: v for k, v in kw.items() if k in allowed and v is not None})
        path = os.path.join(constants.run_dirs(self.expr_name, self.trial_name)["log"], f"{worker_type}.sbatch")
        with open(path, "w") as f:
            f.write(script)
        out = subprocess.run(["sbatch", "--parsable", path], capture_output=True, text=True, check=True).stdout.strip()
        self._job_ids[worker_type] = out.split(";")[0]

    def _state(self, job_id: str) -> JobState:
        out = subprocess.run(["squeue", "-h", "-j", job_id, "-o", "%T"], capture_output=True, text=True).stdout.strip()
        m = {"PENDING": JobState.PENDING, "RUNNING": JobState.RUNNING, "COMPLETED": JobState.COMPLETED, "FAILED": JobState.FAILED,
             "CANCELLED": JobState.CANCELLED, "": JobState.COMPLETED}
        return m.get(out.split("\n")[0], JobState.FAILED)

    def find_all(self, job_name_regex=".*"):
        return [JobInfo(k, self._state(v)) for k, v in self._job_ids.items()]

    def wait(self, timeout=None, poll=10, **kw):
        t0 = time.monotonic()
        while True:
            infos = self.find_all()
            if any(i.state in (JobState.FAILED, JobState.CANCELLED) for i in infos):
                bad = next(i for i in infos if i.state in (JobState.FAILED, JobState.CANCELLED))
                raise JobException(self.run_name, bad.name, "slurm", bad.state)
            if all(i.state == JobState.COMPLETED for i in infos):
                return
            if timeout is not None and time.monotonic() - t0 > timeout:
                raise TimeoutError()
            time.sleep(poll)

    def stop_all(self, signal_=None):
        sig = ["-s", "INT"] if signal_ == signal.SIGINT else []
        if self._job_ids:
            for jid in self._job_ids.values():
                subprocess.run(["scancel"] + sig + [jid])
        else:  # a fresh client (`apps.main stop` from another shell): address the trial's jobs by name
            for wt in ("master_worker", "model_worker"):
                subprocess.run(["scancel"] + sig + ["--name", f"{self.run_name}:{wt}"])


def make(mode: str, expr_name: str, trial_name: str, **kw) -> SchedulerClient:
    if mode == "local":
        return LocalSchedulerClient(expr_name, trial_name)
    if mode == "slurm":
        return SlurmSchedulerClient(expr_name, trial_name, **kw)
    if mode == "ray":
        raise NotImplementedError("ray is not available in this image; use `local` or `slurm`")
    raise NotImplementedError(mode)
