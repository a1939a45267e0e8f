"""Process-wide constants (paths, env flags) and the model-construction scope.

Parity: `realhf/base/constants.py` paths (:57-112: LOG_ROOT, MODEL_SAVE_ROOT, RECOVER_ROOT, env flags).  The
reference also keeps ~40 global accessors (current model's TP group etc.) behind `model_scope`; here layers get an
explicit `ParallelContext`, so `model_scope` only conveys that context to model *factories* during worker setup.
"""

import contextlib
import contextvars
import getpass
import os
from typing import Optional

USER = os.environ.get("USER") or getpass.getuser()


def _default_fileroot() -> str:
    # REAL_FILEROOT > the cluster spec's shared filesystem ($CLUSTER_SPEC_PATH) > a per-user directory under /tmp
    if os.environ.get("REAL_FILEROOT"):
        return os.environ["REAL_FILEROOT"]
    from realhf_b200.base import cluster
    return cluster.spec().fileroot or f"/tmp/realhf_b200/{USER}"


FILEROOT = _default_fileroot()
LOG_ROOT = os.path.join(FILEROOT, "logs")
MODEL_SAVE_ROOT = os.path.join(FILEROOT, "checkpoints")
RECOVER_ROOT = os.path.join(FILEROOT, "recover")
DATASET_CACHE_PATH = os.path.join(FILEROOT, "datasets_cache")
PROFILER_CACHE_PATH = os.path.join(FILEROOT, "profiler")
PARAM_REALLOC_PATH = os.path.join(FILEROOT, "param_realloc")
NCCL_TIMEOUT_MIN = 30

# environment flags forwarded to workers (same names as the reference, constants.py:77-112)
FORWARDED_ENV = ["REAL_MODE", "REAL_RECOVER_RUN", "REAL_SAVE_RECOVER_STATES", "REAL_CUDA_TMARK", "REAL_TIME_MARK", "REAL_DUMP_TRACE",
                 "REAL_DUMP_MEMORY", "REAL_SAVE_MAX_SHARD_SIZE_BYTE", "REAL_FILEROOT", "REAL_NAME_RESOLVE",
                 "REAL_NAME_RESOLVE_ROOT", "REAL_GEMM", "REAL_PDL", "REAL_GEMM_2CTA", "REAL_FUSED_TP", "REAL_REALLOC_DIRECT", "REAL_ISOLATE_GPUS",
                 "REAL_DATASET_CACHE", "REAL_FAULT_INJECT", "REAL_STATUS_TTL", "REAL_TENSORBOARD", "WANDB_MODE", "WANDB_API_KEY", "REAL_WATCH_CONTROLLER", "REAL_ATTN", "REAL_ATTN_BWD", "REAL_LAYERNORM", "REAL_MOE_GROUPED_WGRAD", "CLUSTER_SPEC_PATH", "PYTHONPATH"]

_SCOPE = contextvars.ContextVar("real_model_scope", default=None)


@contextlib.contextmanager
def model_scope(name, ctx, instantiate: bool = True):
    tok = _SCOPE.set(dict(name=name, ctx=ctx, instantiate=instantiate))
    try:
        yield
    finally:
        _SCOPE.reset(tok)


def current_scope() -> Optional[dict]:
    return _SCOPE.get()


def run_dirs(experiment_name: str, trial_name: str):
    d = dict(log=os.path.join(LOG_ROOT, experiment_name, trial_name),
             save=os.path.join(MODEL_SAVE_ROOT, USER, experiment_name, trial_name),
             recover=os.path.join(RECOVER_ROOT, experiment_name, trial_name))
    for p in d.values():
        os.makedirs(p, exist_ok=True)
    return d
