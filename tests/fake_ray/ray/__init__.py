"""A stand-in for the slice of the Ray API that `realhf_b200.scheduler.ray` uses, for machines without Ray (this image).

Tasks run in spawned subprocesses (fresh interpreters, like Ray worker processes): `remote(**opts)(fn).remote(*args)` starts
one, `get` / `wait` / `cancel` act on it, `available_resources()` reports what `FAKE_RAY_RESOURCES` (JSON) says.  Only what
the scheduler client needs is implemented; semantics follow Ray's documented behaviour for those calls.  Put the parent
directory on `sys.path` / `PYTHONPATH` to activate it."""

from __future__ import annotations

import json
import multiprocessing as mp
import os
import pickle
import tempfile
import time
from typing import Any, Dict, List

from . import exceptions  # noqa: F401
from . import util  # noqa: F401

_CTX = mp.get_context("spawn")
_STATE: Dict[str, Any] = dict(init=False, launched=[], gpu_cursor={})


def init(address=None, **kw):
    _STATE["init"] = True
    _STATE["address"] = address
    return dict(address=address)


def is_initialized() -> bool:
    return _STATE["init"]


def shutdown():
    for rec in _STATE["launched"]:
        _kill(rec["ref"])
    _STATE["init"] = False


def available_resources() -> Dict[str, float]:
    spec = os.environ.get("FAKE_RAY_RESOURCES")
    if spec:
        return json.loads(spec)
    return {"CPU": 64.0, "GPU": 0.0, "memory": 64.0 * 1024 ** 3, "node:127.0.0.1": 1.0}


def get_gpu_ids() -> List[int]:
    ids = os.environ.get("FAKE_RAY_GPU_IDS", "")
    return [int(x) for x in ids.split(",") if x != ""]


class ObjectRef:
    def __init__(self, proc, result_path, name):
        self._proc, self._result_path, self.name = proc, result_path, name
        self._cancelled = False

    def __repr__(self):
        return f"ObjectRef({self.name})"


def _task_main(fn, args, opts, result_path, gpu_ids):
    env = ((opts.get("runtime_env") or {}).get("env_vars")) or {}
    os.environ.update(env)
    os.environ["FAKE_RAY_GPU_IDS"] = ",".join(str(g) for g in gpu_ids)
    if gpu_ids:
        os.environ["CUDA_VISIBLE_DEVICES"] = ",".join(str(g) for g in gpu_ids)  # what Ray does for num_gpus > 0
    try:
        out = ("ok", fn(*args))
    except BaseException as e:  # noqa: BLE001 - the driver re-raises it
        import traceback
        out = ("err", f"{type(e).__name__}: {e}\n{traceback.format_exc()}")
    with open(result_path + ".tmp", "wb") as f:
        pickle.dump(out, f)
    os.replace(result_path + ".tmp", result_path)


class _RemoteFunction:
    def __init__(self, fn, opts):
        self._fn, self._opts = fn, opts

    def remote(self, *args):
        assert _STATE["init"], "ray.init() was not called"
        fd, path = tempfile.mkstemp(prefix="fake_ray_result_")
        os.close(fd)
        os.unlink(path)
        gpu_ids = []
        if self._opts.get("num_gpus", 0):
            node = next(iter(self._opts.get("resources") or {"node:127.0.0.1": 1}))
            cur = _STATE["gpu_cursor"].get(node, 0)
            gpu_ids = list(range(cur, cur + int(self._opts["num_gpus"])))
            _STATE["gpu_cursor"][node] = cur + int(self._opts["num_gpus"])
        p = _CTX.Process(target=_task_main, args=(self._fn, args, self._opts, path, gpu_ids), daemon=False)
        p.start()
        ref = ObjectRef(p, path, self._opts.get("name"))
        _STATE["launched"].append(dict(ref=ref, opts=dict(self._opts), args=args, gpu_ids=gpu_ids))
        return ref


def remote(*a, **opts):
    if a and callable(a[0]) and not opts:
        return _RemoteFunction(a[0], {})
    return lambda fn: _RemoteFunction(fn, opts)


def _done(ref: ObjectRef) -> bool:
    return ref._cancelled or not ref._proc.is_alive()


def _result(ref: ObjectRef):
    if ref._cancelled:
        raise exceptions.TaskCancelledError(ref.name)
    ref._proc.join()
    if not os.path.exists(ref._result_path):
        raise exceptions.WorkerCrashedError(f"{ref.name}: the worker process died (exit code {ref._proc.exitcode})")
    with open(ref._result_path, "rb") as f:
        kind, payload = pickle.load(f)
    if kind == "err":
        raise exceptions.RayTaskError(payload)
    return payload


def get(refs, timeout=None):
    single = isinstance(refs, ObjectRef)
    lst = [refs] if single else list(refs)
    deadline = None if timeout is None else time.monotonic() + timeout
    for r in lst:
        while not _done(r):
            if deadline is not None and time.monotonic() >= deadline:
                raise exceptions.GetTimeoutError(f"{r.name} not ready after {timeout}s")
            time.sleep(0.02)
    out = [_result(r) for r in lst]
    return out[0] if single else out


def wait(refs, num_returns=1, timeout=None):
    deadline = None if timeout is None else time.monotonic() + timeout
    while True:
        ready = [r for r in refs if _done(r)]
        if len(ready) >= num_returns or (deadline is not None and time.monotonic() >= deadline):
            ready = ready[:max(num_returns, 0)] if len(ready) > num_returns else ready
            return ready, [r for r in refs if r not in ready]
        time.sleep(0.02)


def _kill(ref: ObjectRef):
    if ref._proc.is_alive():
        try:
            import psutil
            parent = psutil.Process(ref._proc.pid)
            for ch in parent.children(recursive=True):
                ch.kill()
        except Exception:
            pass
        ref._proc.kill()
        ref._proc.join(timeout=10)


def cancel(ref: ObjectRef, force=False, recursive=True):
    if not _done(ref):
        ref._cancelled = True
        _kill(ref)
