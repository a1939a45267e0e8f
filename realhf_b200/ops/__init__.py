"""Native op loader.

`lib()` returns `torch.ops.realhf_b200` after loading `_C/librealhf_b200_ops.so`.  There is exactly
one capability check (`use_native(t)`: the tensor is on a CUDA device): CUDA tensors always take the
sm_100a kernels — a missing library on a GPU box is a hard error, never a silent PyTorch fallback —
while CPU tensors take the plain-PyTorch reference used by the CPU test-suite.
"""

from __future__ import annotations

import os
from pathlib import Path

import torch

_C_DIR = Path(__file__).resolve().parent.parent / "_C"
_LIB_PATH = _C_DIR / "librealhf_b200_ops.so"
_lib = None
_host = None


# kernels launched per op call (for the `gpu_launches` accounting of bench.py)
_KERNELS_PER_OP = {"rmsnorm_bwd": 2, "adamw_step": 1, "decode_attention": 1}


class _LaunchCounter:
    """Counts launches of OUR kernels.  Launches recorded while a CUDA graph is being captured are attributed
    to that graph (`begin_capture` / `end_capture`) and re-counted on every replay (`count_replay`)."""

    def __init__(self):
        self.total = 0
        self.by_op = {}
        self._capturing = None

    def add(self, op: str, n: int = 1):
        self.total += n
        self.by_op[op] = self.by_op.get(op, 0) + n
        if self._capturing is not None:
            self._capturing[0] += n

    def begin_capture(self):
        self._capturing = [0]

    def end_capture(self) -> int:
        n, self._capturing = self._capturing[0], None
        self.total -= n  # a capture does not execute the kernels
        return n

    def count_replay(self, n: int):
        self.total += n
        self.by_op["<graph replay>"] = self.by_op.get("<graph replay>", 0) + n

    def reset(self):
        self.total, self.by_op = 0, {}


launches = _LaunchCounter()


class _CountingLib:
    def __init__(self, ns):
        self._ns = ns
        self._cache = {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            op = getattr(self._ns, name)
            k = _KERNELS_PER_OP.get(name, 1)

            def fn(*a, _op=op, _k=k, _name=name, **kw):
                launches.add(_name, _k)
                return _op(*a, **kw)

            self._cache[name] = fn
        return fn


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            # a checkout without build artefacts (the .so files are git-ignored): build in-tree once, loudly
            try:
                import sys
                print(f"[realhf_b200] {_LIB_PATH.name} missing: building the sm_100a kernels in-tree (takes a few minutes)", file=sys.stderr)
                from realhf_b200.ops import build as _build
                _build.build_all()
            except Exception as e:
                raise RuntimeError(
                    f"{_LIB_PATH} is missing and building it failed ({e}). Build it with `python -m realhf_b200.ops.build` "
                    "(or `python -c 'import __graft_entry__ as g; g.build()'`).") from e
        torch.ops.load_library(str(_LIB_PATH))
        _lib = _CountingLib(torch.ops.realhf_b200)
    return _lib


def host():
    """pybind11 module with host-side native helpers; None if not built (pure-Python fallbacks exist)."""
    global _host
    if _host is None:
        try:
            from realhf_b200._C import host_ext  # type: ignore
            _host = host_ext
        except ImportError:
            if os.environ.get("REAL_REQUIRE_NATIVE", "0") == "1":
                raise
            _host = False
    return _host or None


def use_native(t: torch.Tensor) -> bool:
    return t.is_cuda


def native_built() -> bool:
    return _LIB_PATH.exists()
