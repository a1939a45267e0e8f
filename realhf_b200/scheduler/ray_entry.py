"""Body of a Ray worker task (see `scheduler/ray.py`).  Kept in a module that imports nothing of the framework at import
time: Ray unpickles the task function by importing this module in a fresh worker process, and the launcher's environment
(`REAL_FILEROOT`, the name-resolve backend, ...) has to be in `os.environ` BEFORE `realhf_b200.base.constants` is first
imported there.  Parity: `realhf/system/controller.py:348-395` (`run_ray_worker`)."""

from __future__ import annotations

import argparse
import os
from typing import Dict, Optional


def run_ray_worker(worker_type: str, index: int, world: int, exp: str, trial: str, env: Dict[str, str], slot: Optional[int] = None):
    """Body of one Ray task = one worker process.  Runs in the Ray worker process (a fresh interpreter per task), so process
    state (environment, CUDA device) can be set freely.  `slot` is the rank's position on its node (packed placement)."""
    os.environ.update({k: str(v) for k, v in (env or {}).items()})
    if worker_type == "model_worker" and slot is not None:
        import torch  # must not have touched CUDA yet: the visible-device list is still ours to set
        if torch.cuda.is_initialized():
            raise RuntimeError("CUDA was initialised before the worker chose its device")
        try:
            assigned = [int(g) for g in __import__("ray").get_gpu_ids()]
        except Exception:
            assigned = []
        if os.environ.get("REAL_ISOLATE_GPUS", "0") == "1":
            os.environ["REAL_LOCAL_GPU"] = "0"        # Ray already narrowed CUDA_VISIBLE_DEVICES to the reserved GPU
        else:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)  # peers' devices stay in the process (peer memory / multicast)
            os.environ["REAL_LOCAL_GPU"] = str(assigned[0] if assigned else slot)
    from realhf_b200.apps import remote
    from realhf_b200.base import constants
    # same log file a process-based scheduler would have given this worker (Ray's own log routing keeps working for
    # whatever is printed before / after); restored afterwards because Ray may reuse this process for another task
    log = os.path.join(constants.run_dirs(exp, trial)["log"], f"{worker_type}-{index}")
    saved = _redirect_output(log)
    args = argparse.Namespace(worker_type=worker_type, experiment_name=exp, trial_name=trial, jobstep_id=index, n_jobsteps=world,
                              worker_submission_index=0, wprocs_per_jobstep=1, wprocs_in_job=world, wproc_offset=0)
    try:
        remote.main_worker(args)
    except SystemExit as e:  # main_worker exits non-zero after publishing ERROR: the task must fail, not return
        if e.code not in (0, None):
            raise RuntimeError(f"{worker_type}/{index} failed (see {log} / the ERROR status key)") from e
    finally:
        _restore_output(saved)
    return 0


def _redirect_output(path: str):
    import sys
    sys.stdout.flush()
    sys.stderr.flush()
    saved = (os.dup(1), os.dup(2))
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_APPEND, 0o644)
    os.dup2(fd, 1)
    os.dup2(fd, 2)
    os.close(fd)
    return saved


def _restore_output(saved):
    import sys
    sys.stdout.flush()
    sys.stderr.flush()
    os.dup2(saved[0], 1)
    os.dup2(saved[1], 2)
    os.close(saved[0])
    os.close(saved[1])
