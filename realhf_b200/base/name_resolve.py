"""Key-value "name resolve" store used for rendezvous and liveness.

Parity: `realhf/base/name_resolve.py` — `add / add_subentry / get / get_subtree / find_subtree / wait /
delete / clear_subtree / watch_names / reset`, with an in-memory repository (single process, tests) and a
shared-filesystem repository (the reference's default, `:265-355`); keys may carry a TTL kept alive by a
background thread.  A Redis backend is not provided (redis is not in this image); the file backend works
across nodes on any shared mount, and `realhf_b200.system.rendezvous` offers a TCP store alternative.
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


def make_repository(type_: str = "nfs", **kw) -> NameRecordRepository:
    if type_ == "memory":
        return MemoryNameRecordRepository(**kw)
    if type_ == "nfs":
        return NfsNameRecordRepository(**kw)
    if type_ == "redis":
        raise NotImplementedError("the redis backend needs the `redis` package, which is not available offline")
    raise NotImplementedError(type_)


DEFAULT_REPOSITORY_TYPE = os.environ.get("REAL_NAME_RESOLVE", "nfs")
DEFAULT_REPOSITORY = make_repository(DEFAULT_REPOSITORY_TYPE)


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
