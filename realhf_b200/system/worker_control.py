"""Request / response control plane of the worker processes.

Parity: `realhf/system/worker_base.py` (`WorkerServer` :77-196, `WorkerControlPanel` :217-455, statuses :42-57) and
`system/worker_control.py` (ZMQ REQ/REP transport :17-115).  Every worker process runs a `WorkerServer`: a REP socket on a
random port whose address is published in name_resolve; the launcher (or an operator's shell: `python -m
realhf_b200.apps.main status|pause|resume|stop|ping`) talks to the workers through a `WorkerControlPanel`.

Differences from the reference: the server lives on its own daemon thread instead of being polled from the worker's main
loop — a model worker in the middle of a 10 s generation MFC still answers `ping` / `status` at once, which is what makes
the answers useful for liveness and progress monitoring — and handlers only flip flags / read state that the main loop owns
(`pause`, `resume`, `exit` are picked up between steps, exactly where the name_resolve control key was read before; that
key keeps working for shells that cannot reach the worker's port).  The Ray transport of the reference is not provided.
"""

from __future__ import annotations

import enum
import pickle
import threading
import time
from typing import Any, Callable, Dict, List, Optional, Sequence

import zmq

from realhf_b200.base import logging, name_resolve

logger = logging.getLogger("worker_control")


class WorkerServerStatus(str, enum.Enum):
    """Same vocabulary as the reference (`worker_base.py:42-57`)."""

    READY = "READY"
    RUNNING = "RUNNING"
    PAUSED = "PAUSED"
    COMPLETED = "COMPLETED"
    UNKNOWN = "UNKNOWN"
    INTERRUPTED = "INTERRUPTED"
    ERROR = "ERROR"
    LOST = "LOST"


def control_addr_key(exp: str, trial: str, worker_name: str) -> str:
    return f"{exp}/{trial}/worker_control/{worker_name}"


class WorkerException(Exception):
    def __init__(self, worker_name: str, status: WorkerServerStatus, scenario: str):
        super().__init__(f"worker {worker_name} is {status.value} while {scenario}")
        self.worker_name, self.worker_status, self.scenario = worker_name, status, scenario


class WorkerServer:
    """Serves `command -> handler(**kwargs)` requests for one worker.  Built-in commands: ping, status, pause, resume, exit,
    interrupt; workers add their own with `register_handler` (e.g. `progress`, `memory`)."""

    def __init__(self, exp: str, trial: str, worker_name: str, host: Optional[str] = None):
        from realhf_b200.system.stream import host_ip
        self.exp, self.trial, self.worker_name = exp, trial, worker_name
        self._ctx = zmq.Context.instance()
        self._sock = self._ctx.socket(zmq.REP)
        self._sock.setsockopt(zmq.LINGER, 0)
        port = self._sock.bind_to_random_port("tcp://*")
        self.address = f"tcp://{host or host_ip()}:{port}"
        self._handlers: Dict[str, Callable[..., Any]] = {}
        self.status = WorkerServerStatus.READY
        self.paused = threading.Event()      # set: the main loop must not start new work
        self.exit_requested = threading.Event()
        self._stop = threading.Event()
        self._t0 = time.time()
        self.n_served = 0
        for name, fn in (("ping", lambda: "pong"), ("status", self._status), ("pause", self._pause), ("resume", self._resume),
                         ("exit", self._exit), ("interrupt", self._interrupt)):
            self._handlers[name] = fn
        self._thread = threading.Thread(target=self._serve, name=f"worker-server-{worker_name}", daemon=True)
        self._thread.start()
        name_resolve.add(control_addr_key(exp, trial, worker_name), self.address, replace=True, keepalive_ttl=_ttl())

    # ---- built-ins
    def _status(self):
        return dict(status=self.status.value, uptime_s=round(time.time() - self._t0, 1), served=self.n_served)

    def _pause(self):
        self.paused.set()
        return "pausing"

    def _resume(self):
        self.paused.clear()
        return "resuming"

    def _exit(self):
        self.exit_requested.set()
        self.paused.clear()
        return "exiting"

    def _interrupt(self):
        self.set_status(WorkerServerStatus.INTERRUPTED)
        self.exit_requested.set()
        self.paused.clear()
        return "interrupted"

    # ---- API for the owning worker
    def register_handler(self, command: str, fn: Callable[..., Any]):
        self._handlers[command] = fn

    def set_status(self, status: WorkerServerStatus):
        self.status = status

    def wait_while_paused(self, poll: float = 0.1) -> bool:
        """Blocks while paused; returns False when the worker should exit instead of continuing."""
        if self.paused.is_set() and not self.exit_requested.is_set():
            prev, self.status = self.status, WorkerServerStatus.PAUSED
            while self.paused.is_set() and not self.exit_requested.is_set():
                time.sleep(poll)
            self.status = prev
        return not self.exit_requested.is_set()

    def _serve(self):
        poller = zmq.Poller()
        poller.register(self._sock, zmq.POLLIN)
        while not self._stop.is_set():
            try:
                if not dict(poller.poll(100)):
                    continue
                cmd, kwargs = pickle.loads(self._sock.recv())
            except zmq.ZMQError:
                return
            try:
                fn = self._handlers.get(cmd)
                if fn is None:
                    raise KeyError(f"worker {self.worker_name} has no handler for `{cmd}` (known: {sorted(self._handlers)})")
                reply = ("ok", fn(**(kwargs or {})))
            except Exception as e:  # the caller gets the error; the worker keeps running
                reply = ("err", f"{type(e).__name__}: {e}")
            self.n_served += 1
            try:
                self._sock.send(pickle.dumps(reply))
            except zmq.ZMQError:
                return

    def close(self):
        self._stop.set()
        self._thread.join(timeout=2)
        try:
            name_resolve.delete(control_addr_key(self.exp, self.trial, self.worker_name))
        except Exception:
            pass
        self._sock.close(0)


def _ttl() -> float:
    import os
    return float(os.environ.get("REAL_STATUS_TTL", "120"))


class WorkerControlPanel:
    """Client side: connect to workers by name, send single or group requests, poll statuses."""

    def __init__(self, exp: str, trial: str, timeout: float = 10.0):
        self.exp, self.trial, self.timeout = exp, trial, timeout
        self._ctx = zmq.Context.instance()
        self._socks: Dict[str, zmq.Socket] = {}
        self._addr: Dict[str, str] = {}

    @property
    def worker_names(self) -> List[str]:
        return sorted(self._socks)

    def discover(self) -> List[str]:
        """Names of every worker of the trial that published a control address."""
        root = control_addr_key(self.exp, self.trial, "")
        return sorted(k[len(root):] for k in name_resolve.find_subtree(root.rstrip("/")) if k.startswith(root))

    def connect(self, worker_names: Optional[Sequence[str]] = None, timeout: Optional[float] = None) -> List[str]:
        names = list(worker_names) if worker_names is not None else self.discover()
        for n in names:
            if n in self._socks:
                continue
            addr = name_resolve.wait(control_addr_key(self.exp, self.trial, n), timeout=timeout if timeout is not None else self.timeout)
            self._open(n, addr)
        return names

    def _open(self, name: str, addr: str):
        s = self._ctx.socket(zmq.REQ)
        s.setsockopt(zmq.LINGER, 0)
        s.connect(addr)
        self._socks[name], self._addr[name] = s, addr

    def request(self, worker_name: str, command: str, timeout: Optional[float] = None, **kwargs) -> Any:
        s = self._socks[worker_name]
        s.send(pickle.dumps((command, kwargs)))
        if not s.poll(int(1000 * (timeout if timeout is not None else self.timeout))):
            # a REQ socket that missed its reply is stuck in the wrong state: replace it, report the worker as lost
            s.close(0)
            self._open(worker_name, self._addr[worker_name])
            raise WorkerException(worker_name, WorkerServerStatus.LOST, f"waiting for the reply to `{command}`")
        kind, payload = pickle.loads(s.recv())
        if kind == "err":
            raise RuntimeError(f"worker {worker_name} failed `{command}`: {payload}")
        return payload

    def group_request(self, command: str, worker_names: Optional[Sequence[str]] = None, timeout: Optional[float] = None,
                      worker_kwargs: Optional[Dict[str, Dict]] = None, **kwargs) -> Dict[str, Any]:
        """Send `command` to all (or the given) workers first, then collect: the workers process it concurrently."""
        names = list(worker_names) if worker_names is not None else self.worker_names
        for n in names:
            kw = dict(kwargs, **((worker_kwargs or {}).get(n, {})))
            self._socks[n].send(pickle.dumps((command, kw)))
        out: Dict[str, Any] = {}
        deadline = time.monotonic() + (timeout if timeout is not None else self.timeout)
        for n in names:
            s = self._socks[n]
            if not s.poll(max(0, int(1000 * (deadline - time.monotonic())))):
                s.close(0)
                self._open(n, self._addr[n])
                out[n] = WorkerException(n, WorkerServerStatus.LOST, f"waiting for the reply to `{command}`")
                continue
            kind, payload = pickle.loads(s.recv())
            out[n] = payload if kind == "ok" else RuntimeError(f"worker {n} failed `{command}`: {payload}")
        return out

    def pulse(self) -> Dict[str, WorkerServerStatus]:
        res = self.group_request("status", timeout=min(self.timeout, 5.0))
        return {n: (WorkerServerStatus(r["status"]) if isinstance(r, dict) else WorkerServerStatus.LOST) for n, r in res.items()}

    def close(self):
        for s in self._socks.values():
            s.close(0)
        self._socks.clear()
