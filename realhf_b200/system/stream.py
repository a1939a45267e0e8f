"""Master <-> model-worker request / reply transport over ZMQ.

Parity: `realhf/system/request_reply_stream.py` (Payload :32-59, master/worker streams).  One ROUTER socket on
the master, one DEALER per worker (identity = worker index).  The reference needs a 3-phase syn/ack handshake so
that all workers enqueue overlapping MFCs in the same order (master_worker.py:438-451); here the master posts the
requests of one MFC to all of its workers back-to-back from a single-threaded event loop and every
master->worker channel is FIFO, so every worker observes the master's global dispatch order by construction —
no handshake round trip.
"""

from __future__ import annotations

import dataclasses
import pickle
import os
import socket
import time
import uuid
from typing import Any, List, Optional

import zmq

from realhf_b200.base import name_resolve


def master_addr_key(exp: str, trial: str) -> str:
    return f"{exp}/{trial}/stream/master_addr"


from realhf_b200.base.network import find_free_port as free_port  # noqa: E402,F401  (kept under the names the workers import)
from realhf_b200.base.network import gethostip as host_ip  # noqa: E402,F401


@dataclasses.dataclass
class Payload:
    handler: int                       # model worker index
    handle_name: str                   # spec | fetch | initialize | model_config | generate | inference | train_step | ...
    request_id: str = dataclasses.field(default_factory=lambda: uuid.uuid4().hex)
    data: Any = None
    model_name: Any = None             # ModelName the request addresses (if any)
    pre_hooks: List[str] = dataclasses.field(default_factory=list)
    pre_hook_data: List[Any] = dataclasses.field(default_factory=list)
    post_hooks: List[str] = dataclasses.field(default_factory=list)
    post_hook_data: List[Any] = dataclasses.field(default_factory=list)
    is_reply: bool = False
    error: Optional[str] = None


class MasterStream:
    def __init__(self, exp: str, trial: str, n_workers: int):
        self.ctx = zmq.Context.instance()
        self.sock = self.ctx.socket(zmq.ROUTER)
        self.sock.setsockopt(zmq.LINGER, 0)
        port = self.sock.bind_to_random_port("tcp://*")
        name_resolve.add(master_addr_key(exp, trial), f"tcp://{host_ip()}:{port}", replace=True)
        self.n_workers = n_workers
        self._ready = set()

    def wait_workers(self, timeout: float = 600.0):
        t0 = time.monotonic()
        while len(self._ready) < self.n_workers:
            if self.sock.poll(100):
                ident, raw = self.sock.recv_multipart()
                msg = pickle.loads(raw)
                if msg == "hello":
                    self._ready.add(ident)
            if time.monotonic() - t0 > timeout:
                raise TimeoutError(f"only {len(self._ready)}/{self.n_workers} model workers connected")

    def post(self, p: Payload) -> str:
        self.sock.send_multipart([f"mw{p.handler}".encode(), pickle.dumps(p)])
        return p.request_id

    def poll(self, timeout_ms: int = 0) -> Optional[Payload]:
        if not self.sock.poll(timeout_ms):
            return None
        _, raw = self.sock.recv_multipart()
        msg = pickle.loads(raw)
        return msg if isinstance(msg, Payload) else None

    def close(self):
        self.sock.close()


class WorkerStream:
    def __init__(self, exp: str, trial: str, worker_index: int, timeout: float = 600.0):
        addr = name_resolve.wait(master_addr_key(exp, trial), timeout=timeout)
        self.ctx = zmq.Context.instance()
        self.sock = self.ctx.socket(zmq.DEALER)
        self.sock.setsockopt(zmq.IDENTITY, f"mw{worker_index}".encode())
        self.sock.setsockopt(zmq.LINGER, 0)
        self.sock.connect(addr)
        self.sock.send(pickle.dumps("hello"))

    def poll(self, timeout_ms: int = 0) -> Optional[Payload]:
        if not self.sock.poll(timeout_ms):
            return None
        return pickle.loads(self.sock.recv())

    def reply(self, req: Payload, data: Any = None, error: Optional[str] = None):
        self.sock.send(pickle.dumps(Payload(handler=req.handler, handle_name=req.handle_name, request_id=req.request_id,
                                            data=data, model_name=req.model_name, is_reply=True, error=error)))

    def close(self):
        self.sock.close()
