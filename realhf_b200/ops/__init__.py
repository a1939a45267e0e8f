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


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(
                f"{_LIB_PATH} is missing. Build it with `python -m realhf_b200.ops.build` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`).")
        torch.ops.load_library(str(_LIB_PATH))
        _lib = torch.ops.realhf_b200
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
