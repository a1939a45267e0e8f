"""Is Ray usable here?  (parity: `realhf/base/ray_utils.py`; the scheduler client `scheduler/ray.py` imports Ray lazily.)"""

from __future__ import annotations

import importlib.util
import shutil


def check_ray_availability() -> bool:
    """The `ray` package is importable and its CLI is on PATH (a launcher needs both: `ray.init` and `ray status` / `ray start`)."""
    return importlib.util.find_spec("ray") is not None and shutil.which("ray") is not None
