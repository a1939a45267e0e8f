"""Key-value "name resolve" store used for rendezvous and liveness.

Parity: `realhf/base/name_resolve.py` — `add / add_subentry / get / get_subtree / find_subtree / wait /
delete / clear_subtree / watch_names / reset`, with an in-memory repository (single process, tests) and a
shared-filesystem repository (the reference's default, `:265-355`); keys may carry a TTL kept alive by a
background thread, and a Redis repository (`:357-500`).  The Redis backend speaks the wire protocol (RESP2) itself over a
TCP socket, so it needs no `redis` package; `MiniRedisServer` is a small RESP server with the command subset the
repository uses (SET NX/EX, GET, DEL, KEYS, EXPIRE, TTL, AUTH, PING) for clusters without a Redis deployment and for tests:
`python -m realhf_b200.base.name_resolve --serve 6379`.
"""

from __future__ import annotations

import os
import random
import shutil
import threading
import time
import uuid
from typing import Callable, Dict, List, Optional


class ArgumentError(Exception):
    pass


class NameEntryExistsError(Exception):
    pass


class NameEntryNotFoundError(Exception):
    pass


class NameRecordRepository:
    def add(self, name, value, delete_on_exit=True, keepalive_ttl=None, replace=False):
        raise NotImplementedError()

    def add_subentry(self, name, value, **kw):
        sub = f"{name.rstrip('/')}/{uuid.uuid4().hex[:8]}"
        self.add(sub, value, **kw)
        return sub

    def delete(self, name):
        raise NotImplementedError()

    def clear_subtree(self, name_root):
        raise NotImplementedError()

    def get(self, name):
        raise NotImplementedError()

    def get_subtree(self, name_root) -> List[str]:
        raise NotImplementedError()

    def find_subtree(self, name_root) -> List[str]:
        raise NotImplementedError()

    def wait(self, name, timeout: Optional[float] = None, poll_frequency: float = 0.05):
        t0 = time.monotonic()
        while True:
            try:
                return self.get(name)
            except NameEntryNotFoundError:
                pass
            if timeout is not None and time.monotonic() - t0 > timeout:
                raise TimeoutError(f"timed out waiting for key `{name}`")
            time.sleep(poll_frequency + random.random() * 0.01)

    def watch_names(self, names: List[str], call_back: Callable, poll_frequency: float = 15, wait_timeout: float = 300):
        """Call `call_back()` once any of `names` disappears (used by workers to exit when the controller dies)."""
        if isinstance(names, str):
            names = [names]

        def _watch():
            for n in names:
                try:
                    self.wait(n, timeout=wait_timeout)
                except TimeoutError:
                    call_back()
                    return
            while True:
                for n in names:
                    try:
                        self.get(n)
                    except NameEntryNotFoundError:
                        call_back()
                        return
                time.sleep(poll_frequency)

        t = threading.Thread(target=_watch, daemon=True)
        t.start()
        return t

    def reset(self):
        pass


class MemoryNameRecordRepository(NameRecordRepository):
    def __init__(self):
        self._store: Dict[str, str] = {}
        self._lock = threading.Lock()

    def add(self, name, value, delete_on_exit=True, keepalive_ttl=None, replace=False):
        name = name.rstrip("/")
        if not name:
            raise ArgumentError("empty name")
        with self._lock:
            if name in self._store and not replace:
                raise NameEntryExistsError(name)
            self._store[name] = str(value)
            # the owner lives in this process, so a TTL'd key of the in-memory store can never go stale

    def delete(self, name):
        with self._lock:
            if name not in self._store:
                raise NameEntryNotFoundError(name)
            del self._store[name]

    def clear_subtree(self, name_root):
        root = name_root.rstrip("/")
        with self._lock:
            for k in [k for k in self._store if k == root or k.startswith(root + "/")]:
                del self._store[k]

    def get(self, name):
        name = name.rstrip("/")
        with self._lock:
            if name not in self._store:
                raise NameEntryNotFoundError(name)
            return self._store[name]

    def find_subtree(self, name_root):
        root = name_root.rstrip("/")
        with self._lock:
            return sorted(k for k in self._store if k == root or k.startswith(root + "/"))

    def get_subtree(self, name_root):
        with self._lock:
            store = dict(self._store)
        root = name_root.rstrip("/")
        return [store[k] for k in sorted(store) if k == root or k.startswith(root + "/")]

    def reset(self):
        with self._lock:
            self._store.clear()


class NfsNameRecordRepository(NameRecordRepository):
    """One file per key under a shared directory; writes are atomic renames."""

    def __init__(self, record_root: Optional[str] = None):
        self.root = record_root or os.environ.get("REAL_NAME_RESOLVE_ROOT") or self._default_root()
        self._to_delete = set()
        self._keepalive: Dict[str, float] = {}
        self._ka_thread: Optional[threading.Thread] = None

    @staticmethod
    def _default_root() -> str:
        # multi-node runs need the store on the shared filesystem: follow the cluster spec's fileroot when there is one
        from realhf_b200.base import cluster
        fr = cluster.spec().fileroot
        return os.path.join(fr, "name_resolve") if fr else "/tmp/realhf_b200/name_resolve"

    def _dir(self, name):
        return os.path.join(self.root, name.strip("/"))

    def _file(self, name):
        return os.path.join(self._dir(name), "ENTRY")

    def _ttl_file(self, name):
        return os.path.join(self._dir(name), "TTL")

    def _expired(self, name) -> bool:
        """A key written with `keepalive_ttl` is live only while its owner keeps touching it: once the entry has not been
        refreshed for a full TTL (owner killed, node lost) readers treat it as gone (reference: etcd / redis lease expiry)."""
        try:
            with open(self._ttl_file(name)) as fh:
                ttl = float(fh.read())
            return time.time() - os.path.getmtime(self._file(name)) > ttl
        except (OSError, ValueError):
            return False

    def add(self, name, value, delete_on_exit=True, keepalive_ttl=None, replace=False):
        if not name.strip("/"):
            raise ArgumentError("empty name")
        f = self._file(name)
        if os.path.isfile(f) and not replace:
            raise NameEntryExistsError(name)
        os.makedirs(os.path.dirname(f), exist_ok=True)
        tmp = f + f".tmp{uuid.uuid4().hex[:6]}"
        with open(tmp, "w") as fh:
            fh.write(str(value))
        os.replace(tmp, f)
        if delete_on_exit:
            self._to_delete.add(name)
        if keepalive_ttl is not None:
            with open(self._ttl_file(name), "w") as fh:
                fh.write(str(float(keepalive_ttl)))
            self._keepalive[name] = keepalive_ttl
            self._ensure_keepalive()
        else:
            self._keepalive.pop(name, None)
            try:
                os.remove(self._ttl_file(name))
            except OSError:
                pass

    def _ensure_keepalive(self):
        if self._ka_thread is not None:
            return

        def _touch():
            while True:
                for n in list(self._keepalive):
                    try:
                        os.utime(self._file(n))
                    except OSError:
                        pass
                time.sleep(max(1.0, min(self._keepalive.values(), default=10) / 3))

        self._ka_thread = threading.Thread(target=_touch, daemon=True)
        self._ka_thread.start()

    def delete(self, name):
        f = self._file(name)
        if not os.path.isfile(f):
            raise NameEntryNotFoundError(name)
        os.remove(f)
        try:
            os.remove(self._ttl_file(name))
        except OSError:
            pass
        self._to_delete.discard(name)
        self._keepalive.pop(name, None)
        d = os.path.dirname(f)
        while d != self.root and d.startswith(self.root):
            try:
                os.rmdir(d)
            except OSError:
                break
            d = os.path.dirname(d)

    def clear_subtree(self, name_root):
        shutil.rmtree(self._dir(name_root), ignore_errors=True)

    def get(self, name):
        f = self._file(name)
        if self._expired(name):
            raise NameEntryNotFoundError(f"{name} (lease expired)")
        for _ in range(3):
            try:
                with open(f) as fh:
                    return fh.read()
            except FileNotFoundError:
                raise NameEntryNotFoundError(name)
            except OSError:
                time.sleep(0.01)
        raise NameEntryNotFoundError(name)

    def find_subtree(self, name_root):
        base = self._dir(name_root)
        out = []
        for d, _, files in os.walk(base):
            if "ENTRY" in files:
                out.append(os.path.relpath(d, self.root))
        return sorted(out)

    def get_subtree(self, name_root):
        out = []
        for k in self.find_subtree(name_root):
            try:
                out.append(self.get(k))
            except NameEntryNotFoundError:
                pass
        return out

    def reset(self):
        for n in list(self._to_delete):
            try:
                self.delete(n)
            except NameEntryNotFoundError:
                pass
        self._to_delete.clear()



# ------------------------------------------------------------------------------------------------ redis (RESP2 over TCP)


class _RespConnection:
    """Minimal RESP2 client: one blocking TCP connection, commands as arrays of bulk strings."""

    def __init__(self, host: str, port: int, password: Optional[str] = None, timeout: float = 10.0):
        import socket
        self._sock = socket.create_connection((host, port), timeout=timeout)
        self._rf = self._sock.makefile("rb")
        self._lock = threading.Lock()
        if password:
            self.call("AUTH", password)

    @staticmethod
    def encode(*args) -> bytes:
        out = [b"*%d\r\n" % len(args)]
        for a in args:
            b = a if isinstance(a, bytes) else str(a).encode()
            out.append(b"$%d\r\n%s\r\n" % (len(b), b))
        return b"".join(out)

    def _read(self):
        line = self._rf.readline()
        if not line:
            raise ConnectionError("redis connection closed")
        t, rest = line[:1], line[1:-2]
        if t == b"+":
            return rest.decode()
        if t == b"-":
            raise RuntimeError(f"redis error: {rest.decode()}")
        if t == b":":
            return int(rest)
        if t == b"$":
            n = int(rest)
            if n < 0:
                return None
            data = self._rf.read(n + 2)
            return data[:-2].decode()
        if t == b"*":
            n = int(rest)
            return None if n < 0 else [self._read() for _ in range(n)]
        raise RuntimeError(f"bad RESP reply: {line!r}")

    def call(self, *args):
        with self._lock:
            self._sock.sendall(self.encode(*args))
            return self._read()

    def close(self):
        try:
            self._rf.close()
            self._sock.close()
        except OSError:
            pass


class RedisNameRecordRepository(NameRecordRepository):
    """Keys live in a Redis server (`REAL_REDIS_HOST` / `REAL_REDIS_PORT` / `REAL_REDIS_PASSWORD`, or constructor arguments).
    TTL'd keys use the server's own expiry (`SET ... EX`), refreshed by a keep-alive thread at a third of the TTL, so a
    process that dies silently loses its keys without anybody cleaning up (reference: name_resolve.py:357-500)."""

    KEEPALIVE_POLL = 1.0

    def __init__(self, host: Optional[str] = None, port: Optional[int] = None, password: Optional[str] = None):
        self.host = host or os.environ.get("REAL_REDIS_HOST", "127.0.0.1")
        self.port = int(port or os.environ.get("REAL_REDIS_PORT", "6379"))
        self._conn = _RespConnection(self.host, self.port, password or os.environ.get("REAL_REDIS_PASSWORD"))
        self._to_delete: set = set()
        self._ttl: Dict[str, float] = {}
        self._lock = threading.Lock()
        self._stop = threading.Event()
        self._ka: Optional[threading.Thread] = None

    def _keepalive(self):
        last: Dict[str, float] = {}
        while not self._stop.wait(self.KEEPALIVE_POLL):
            now = time.monotonic()
            with self._lock:
                items = list(self._ttl.items())
            for name, ttl in items:
                if now - last.get(name, 0.0) >= ttl / 3:
                    try:
                        self._conn.call("EXPIRE", name, max(1, int(ttl)))
                        last[name] = now
                    except (OSError, RuntimeError, ConnectionError):
                        pass

    def add(self, name, value, delete_on_exit=True, keepalive_ttl=None, replace=False):
        name = name.rstrip("/")
        if not name:
            raise ArgumentError("empty name")
        args = ["SET", name, str(value)]
        if keepalive_ttl is not None:
            args += ["EX", max(1, int(keepalive_ttl))]
        if not replace:
            args.append("NX")
        if self._conn.call(*args) is None:
            raise NameEntryExistsError(name)
        with self._lock:
            if delete_on_exit:
                self._to_delete.add(name)
            if keepalive_ttl is not None:
                self._ttl[name] = float(keepalive_ttl)
                if self._ka is None:
                    self._ka = threading.Thread(target=self._keepalive, daemon=True)
                    self._ka.start()

    def delete(self, name):
        name = name.rstrip("/")
        with self._lock:
            self._to_delete.discard(name)
            self._ttl.pop(name, None)
        if self._conn.call("DEL", name) == 0:
            raise NameEntryNotFoundError(name)

    def _keys(self, name_root) -> List[str]:
        root = name_root.rstrip("/")
        esc = "".join("\\" + ch if ch in "*?[]\\" else ch for ch in root)
        keys = set(self._conn.call("KEYS", esc + "/*") or [])
        if self._conn.call("EXISTS", root):
            keys.add(root)
        return sorted(keys)

    def clear_subtree(self, name_root):
        keys = self._keys(name_root)
        if keys:
            self._conn.call("DEL", *keys)
        with self._lock:
            for k in keys:
                self._to_delete.discard(k)
                self._ttl.pop(k, None)

    def get(self, name):
        v = self._conn.call("GET", name.rstrip("/"))
        if v is None:
            raise NameEntryNotFoundError(name)
        return v

    def find_subtree(self, name_root):
        return self._keys(name_root)

    def get_subtree(self, name_root):
        out = []
        for k in self._keys(name_root):
            v = self._conn.call("GET", k)
            if v is not None:
                out.append(v)
        return out

    def reset(self):
        self._stop.set()
        with self._lock:
            names, self._to_delete, self._ttl = list(self._to_delete), set(), {}
        for n in names:
            try:
                self._conn.call("DEL", n)
            except (OSError, RuntimeError, ConnectionError):
                pass
        self._stop = threading.Event()
        self._ka = None


class MiniRedisServer:
    """A threaded RESP2 server with the command subset `RedisNameRecordRepository` uses.  Not a database: an in-memory dict with
    expiry, meant as the rendezvous store of a cluster that has no Redis (and as the test double of one)."""

    def __init__(self, host: str = "127.0.0.1", port: int = 0, password: Optional[str] = None):
        import socketserver
        store: Dict[str, str] = {}
        expiry: Dict[str, float] = {}
        lock = threading.Lock()

        def alive(k):
            e = expiry.get(k)
            if e is not None and time.monotonic() >= e:
                store.pop(k, None)
                expiry.pop(k, None)
            return k in store

        def run(cmd: List[str], authed: List[bool]):
            op = cmd[0].upper()
            if op == "AUTH":
                authed[0] = cmd[-1] == password
                return "+OK" if authed[0] else "-ERR invalid password"
            if password and not authed[0]:
                return "-NOAUTH Authentication required."
            with lock:
                if op == "PING":
                    return "+PONG"
                if op == "SET":
                    k, v, rest = cmd[1], cmd[2], [c.upper() for c in cmd[3:]]
                    if "NX" in rest and alive(k):
                        return None
                    store[k] = v
                    expiry.pop(k, None)
                    if "EX" in rest:
                        expiry[k] = time.monotonic() + float(cmd[3 + rest.index("EX") + 1])
                    return "+OK"
                if op == "GET":
                    return ("$", store[cmd[1]]) if alive(cmd[1]) else None
                if op == "DEL":
                    n = 0
                    for k in cmd[1:]:
                        if alive(k):
                            store.pop(k)
                            expiry.pop(k, None)
                            n += 1
                    return n
                if op == "EXISTS":
                    return sum(1 for k in cmd[1:] if alive(k))
                if op == "EXPIRE":
                    if not alive(cmd[1]):
                        return 0
                    expiry[cmd[1]] = time.monotonic() + float(cmd[2])
                    return 1
                if op == "TTL":
                    if not alive(cmd[1]):
                        return -2
                    return -1 if cmd[1] not in expiry else max(0, int(expiry[cmd[1]] - time.monotonic()))
                if op == "KEYS":
                    import fnmatch
                    import re
                    rx = re.compile(fnmatch.translate(cmd[1]))
                    return [k for k in sorted(store) if alive(k) and rx.match(k)]
            return f"-ERR unknown command '{cmd[0]}'"

        def enc(r) -> bytes:
            if r is None:
                return b"$-1\r\n"
            if isinstance(r, int):
                return b":%d\r\n" % r
            if isinstance(r, tuple):
                b = r[1].encode()
                return b"$%d\r\n%s\r\n" % (len(b), b)
            if isinstance(r, list):
                return b"*%d\r\n" % len(r) + b"".join(enc(("$", x)) for x in r)
            return r.encode() + b"\r\n"

        class Handler(socketserver.StreamRequestHandler):
            def handle(self):
                authed = [False]
                while True:
                    line = self.rfile.readline()
                    if not line:
                        return
                    if not line.startswith(b"*"):
                        continue
                    parts = []
                    for _ in range(int(line[1:-2])):
                        n = int(self.rfile.readline()[1:-2])
                        parts.append(self.rfile.read(n + 2)[:-2].decode())
                    self.wfile.write(enc(run(parts, authed)))
                    self.wfile.flush()

        class Server(socketserver.ThreadingTCPServer):
            allow_reuse_address = True
            daemon_threads = True

        self._srv = Server((host, port), Handler)
        self.host, self.port = self._srv.server_address
        self._thread = threading.Thread(target=self._srv.serve_forever, daemon=True)

    def start(self) -> "MiniRedisServer":
        self._thread.start()
        return self

    def stop(self):
        self._srv.shutdown()
        self._srv.server_close()


def make_repository(type_: str = "nfs", **kw) -> NameRecordRepository:
    if type_ == "memory":
        return MemoryNameRecordRepository(**kw)
    if type_ == "nfs":
        return NfsNameRecordRepository(**kw)
    if type_ == "redis":
        return RedisNameRecordRepository(**kw)
    raise NotImplementedError(type_)


DEFAULT_REPOSITORY_TYPE = os.environ.get("REAL_NAME_RESOLVE", "nfs")
try:
    DEFAULT_REPOSITORY = make_repository(DEFAULT_REPOSITORY_TYPE)
except OSError as _e:  # REAL_NAME_RESOLVE=redis without a reachable server: say so at first use instead of at import
    import warnings
    warnings.warn(f"name_resolve backend {DEFAULT_REPOSITORY_TYPE!r} is unavailable ({_e}); falling back to the file store")
    DEFAULT_REPOSITORY_TYPE = "nfs"
    DEFAULT_REPOSITORY = make_repository("nfs")


def reconfigure(*a, **kw):
    global DEFAULT_REPOSITORY, DEFAULT_REPOSITORY_TYPE
    DEFAULT_REPOSITORY.reset()
    DEFAULT_REPOSITORY = make_repository(*a, **kw)
    DEFAULT_REPOSITORY_TYPE = a[0] if a else kw.get("type_", "nfs")


def add(*a, **kw):
    return DEFAULT_REPOSITORY.add(*a, **kw)


def add_subentry(*a, **kw):
    return DEFAULT_REPOSITORY.add_subentry(*a, **kw)


def delete(*a, **kw):
    return DEFAULT_REPOSITORY.delete(*a, **kw)


def clear_subtree(*a, **kw):
    return DEFAULT_REPOSITORY.clear_subtree(*a, **kw)


def get(*a, **kw):
    return DEFAULT_REPOSITORY.get(*a, **kw)


def get_subtree(*a, **kw):
    return DEFAULT_REPOSITORY.get_subtree(*a, **kw)


def find_subtree(*a, **kw):
    return DEFAULT_REPOSITORY.find_subtree(*a, **kw)


def wait(*a, **kw):
    return DEFAULT_REPOSITORY.wait(*a, **kw)


def watch_names(*a, **kw):
    return DEFAULT_REPOSITORY.watch_names(*a, **kw)


def reset():
    return DEFAULT_REPOSITORY.reset()


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="RESP key-value server for name_resolve (REAL_NAME_RESOLVE=redis) on clusters without Redis")
    ap.add_argument("--serve", type=int, required=True, metavar="PORT")
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--password", default=os.environ.get("REAL_REDIS_PASSWORD"))
    a = ap.parse_args()
    srv = MiniRedisServer(a.host, a.serve, a.password).start()
    print(f"name_resolve store listening on {srv.host}:{srv.port}", flush=True)
    try:
        while True:
            time.sleep(3600)
    except KeyboardInterrupt:
        srv.stop()
