"""Secrets kept out of configs.  Parity: `realhf/base/security.py` (read_key)."""

from __future__ import annotations

import os


def read_key(service: str, name: str = "default") -> str:
    """Content of `$REAL_KEY_ROOT/<service>/<name>` (default root `~/.real_keys`), stripped; e.g. a wandb or Redis password."""
    root = os.environ.get("REAL_KEY_ROOT", os.path.expanduser("~/.real_keys"))
    with open(os.path.join(root, service, name)) as f:
        return f.read().strip()
