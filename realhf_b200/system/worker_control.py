"""Request / response control plane of the worker processes.

Parity: `realhf/system/worker_base.py` (`WorkerServer` :77-196, `WorkerControlPanel` :217-455, statuses :42-57) and
`system/worker_control.py` (ZMQ REQ/REP transport :17-115).  Every worker process runs a `WorkerServer`: a REP socket on a
random port whose address is published in name_resolve; the launcher (or an operator's shell: `python -m
realhf_b200.apps.main status|pause|resume|stop|ping`) talks to the workers through a `WorkerControlPanel`.

Differences from the reference: the server lives on its own daemon thread instead of being polled from the worker's main
loop — a model worker in the middle of a 10 s generation MFC still answers `ping` / `status` at once, which is what makes
the answers useful for liveness and progress monitoring — and handlers only flip flags / read state that the main loop owns
(`pause`, `resume`, `exit` are picked up between steps, exactly where the name_resolve control key was read before; that
key keeps working for shells that cannot reach the worker's port).

Two transports, as in the reference: ZMQ REQ/REP (default; the address is published in name_resolve) and a pair of queues
(`comm=(request_queue, reply_queue)`: anything with `put(item, timeout=)` / `get(timeout=)` — `ray.util.queue.Queue` in Ray
mode (reference `worker_control.py:48-64`, `:118-150`), `multiprocessing` / `queue.Queue` elsewhere), for deployments where
the workers' ports are not reachable from the controller.
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


class _ZmqRepTransport:
    def __init__(self, host: Optional[str]):
        from realhf_b200.system.stream import host_ip
        self._sock = zmq.Context.instance().socket(zmq.REP)
        self._sock.setsockopt(zmq.LINGER, 0)
        port = self._sock.bind_to_random_port("tcp://*")
        self.address = f"tcp://{host or host_ip()}:{port}"
        self._poller = zmq.Poller()
        self._poller.register(self._sock, zmq.POLLIN)

    def recv(self, timeout_ms: int) -> Optional[bytes]:
        """None: nothing arrived in time.  Raises `ConnectionError` once the socket is gone."""
        try:
            if not dict(self._poller.poll(timeout_ms)):
                return None
            return self._sock.recv()
        except zmq.ZMQError as e:
            raise ConnectionError(str(e)) from e

    def send(self, data: bytes):
        try:
            self._sock.send(data)
        except zmq.ZMQError as e:
            raise ConnectionError(str(e)) from e

    def close(self):
        self._sock.close(0)


class _QueueTransport:
    """Server side of the queue transport: requests arrive on `comm[0]`, replies leave on `comm[1]`."""

    address = "queue://"

    def __init__(self, comm):
        self._req, self._rep = comm

    def recv(self, timeout_ms: int) -> Optional[bytes]:
        try:
            return self._req.get(timeout=timeout_ms / 1000.0)
        except Exception as e:  # queue.Empty / ray.util.queue.Empty (distinct classes, same meaning)
            if type(e).__name__ == "Empty":
                return None
            raise ConnectionError(f"{type(e).__name__}: {e}") from e

    def send(self, data: bytes):
        self._rep.put(data)

    def close(self):
        pass


class WorkerServer:
    """Serves `command -> handler(**kwargs)` requests for one worker.  Built-in commands: ping, status, pause, resume, exit,
    interrupt; workers add their own with `register_handler` (e.g. `progress`, `memory`)."""

    def __init__(self, exp: str, trial: str, worker_name: str, host: Optional[str] = None, comm=None):
        self.exp, self.trial, self.worker_name = exp, trial, worker_name
        self._transport = _QueueTransport(comm) if comm is not None else _ZmqRepTransport(host)
        self.address = self._transport.address
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
        while not self._stop.is_set():
            try:
                raw = self._transport.recv(100)
                if raw is None:
                    continue
                cmd, kwargs = pickle.loads(raw)
            except ConnectionError:
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
                self._transport.send(pickle.dumps(reply))
            except ConnectionError:
                return

    def close(self):
        self._stop.set()
        self._thread.join(timeout=2)
        try:
            name_resolve.delete(control_addr_key(self.exp, self.trial, self.worker_name))
        except Exception:
            pass
        self._transport.close()


def _ttl() -> float:
    import os
    return float(os.environ.get("REAL_STATUS_TTL", "120"))


class _ZmqReqChannel:
    def __init__(self, addr: str):
        self.addr = addr
        self._open()

    def _open(self):
        self._s = zmq.Context.instance().socket(zmq.REQ)
        self._s.setsockopt(zmq.LINGER, 0)
        self._s.connect(self.addr)

    def send(self, data: bytes):
        self._s.send(data)

    def recv(self, timeout_s: float) -> Optional[bytes]:
        if not self._s.poll(max(0, int(1000 * timeout_s))):
            # a REQ socket that missed its reply is stuck in the wrong state: replace it
            self._s.close(0)
            self._open()
            return None
        return self._s.recv()

    def close(self):
        self._s.close(0)


class _QueueChannel:
    """Client side of the queue transport.  A reply that arrives after its request timed out would be mistaken for the
    answer to the next request, so the reply queue is drained before every new request."""

    def __init__(self, request_q, reply_q):
        self._req, self._rep = request_q, reply_q

    def send(self, data: bytes):
        while True:
            try:
                self._rep.get(block=False)
            except Exception as e:
                if type(e).__name__ == "Empty":
                    break
                raise
        self._req.put(data)

    def recv(self, timeout_s: float) -> Optional[bytes]:
        try:
            return self._rep.get(timeout=max(timeout_s, 1e-3))
        except Exception as e:
            if type(e).__name__ == "Empty":
                return None
            raise

    def close(self):
        pass


class WorkerControlPanel:
    """Client side: connect to workers by name, send single or group requests, poll statuses."""

    def __init__(self, exp: str, trial: str, timeout: float = 10.0):
        self.exp, self.trial, self.timeout = exp, trial, timeout
        self._socks: Dict[str, Any] = {}

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
        if not addr.startswith("tcp://"):
            raise ValueError(f"worker {name} serves `{addr}`: queue-transport workers are reached with `attach_queues`")
        self._socks[name] = _ZmqReqChannel(addr)

    def attach_queues(self, worker_name: str, request_q, reply_q):
        """Queue transport: talk to `worker_name` through the queue pair its `WorkerServer(comm=...)` was built with."""
        self._socks[worker_name] = _QueueChannel(request_q, reply_q)

    def request(self, worker_name: str, command: str, timeout: Optional[float] = None, **kwargs) -> Any:
        s = self._socks[worker_name]
        s.send(pickle.dumps((command, kwargs)))
        raw = s.recv(timeout if timeout is not None else self.timeout)
        if raw is None:
            raise WorkerException(worker_name, WorkerServerStatus.LOST, f"waiting for the reply to `{command}`")
        kind, payload = pickle.loads(raw)
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
            raw = self._socks[n].recv(max(0.0, deadline - time.monotonic()))
            if raw is None:
                out[n] = WorkerException(n, WorkerServerStatus.LOST, f"waiting for the reply to `{command}`")
                continue
            kind, payload = pickle.loads(raw)
            out[n] = payload if kind == "ok" else RuntimeError(f"worker {n} failed `{command}`: {payload}")
        return out

    def pulse(self) -> Dict[str, WorkerServerStatus]:
        res = self.group_request("status", timeout=min(self.timeout, 5.0))
        return {n: (WorkerServerStatus(r["status"]) if isinstance(r, dict) else WorkerServerStatus.LOST) for n, r in res.items()}

    def close(self):
        for s in self._socks.values():
            s.close()
        self._socks.clear()
