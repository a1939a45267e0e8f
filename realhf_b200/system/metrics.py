"""Where the master worker's per-step statistics go.

Always: `stats.jsonl` in the run's log directory (one JSON object per MFC reply that carries statistics).
Optional: TensorBoard event files (`REAL_TENSORBOARD=1` or `tensorboard=True` on the command line; written under
`<log dir>/tensorboard`) and Weights & Biases (`wandb_mode=online|offline`, i.e. `WANDB_MODE`; offline runs stay on disk under
the log directory).  The reference carries wandb / tensorboard fields in its configs but never writes to either
(`api/core/system_api.py:96-104`).  A sink that fails is switched off with one warning: statistics must
never stop a training run."""

from __future__ import annotations

import json
import numbers
import os
from typing import Callable, Dict, List, Optional

from realhf_b200.base import logging

logger = logging.getLogger("metrics")


class MetricSinks:
    def __init__(self, exp: str, trial: str, log_dir: str):
        self.exp, self.trial, self.log_dir = exp, trial, log_dir
        self._sinks: List[List] = []  # [name, callable(rec), close()]
        os.makedirs(log_dir, exist_ok=True)
        f = open(os.path.join(log_dir, "stats.jsonl"), "a")
        self._sinks.append(["jsonl", lambda rec, f=f: (f.write(json.dumps(rec) + "\n"), f.flush()), f.close])
        if os.environ.get("REAL_TENSORBOARD", "0") == "1":
            self._try("tensorboard", self._open_tensorboard)
        if os.environ.get("WANDB_MODE", "disabled") in ("online", "offline"):
            self._try("wandb", self._open_wandb)

    def _try(self, name: str, opener: Callable):
        try:
            write, close = opener()
            self._sinks.append([name, write, close])
        except Exception as e:  # missing package, no credentials, read-only disk ...
            logger.warning(f"{name} logging disabled: {e!r}")

    @staticmethod
    def _scalars(rec: Dict) -> Dict[str, float]:
        rpc = rec.get("rpc", "train")
        return {f"{rpc}/{k}": float(v) for k, v in rec.items()
                if k not in ("rpc", "step", "epoch", "time") and isinstance(v, numbers.Real) and not isinstance(v, bool)}

    def _open_tensorboard(self):
        from torch.utils.tensorboard import SummaryWriter
        w = SummaryWriter(log_dir=os.path.join(self.log_dir, "tensorboard"))

        def write(rec):
            for k, v in self._scalars(rec).items():
                w.add_scalar(k, v, global_step=int(rec.get("step", 0)), walltime=rec.get("time"))
            w.flush()
        return write, w.close

    def _open_wandb(self):
        import wandb
        os.environ.setdefault("WANDB_DIR", self.log_dir)
        run = wandb.init(project=self.exp, name=self.trial, mode=os.environ["WANDB_MODE"], dir=self.log_dir, resume="allow")

        def write(rec):
            run.log({**self._scalars(rec), "epoch": rec.get("epoch", 0)}, step=int(rec.get("step", 0)))
        return write, run.finish

    def log(self, rec: Dict):
        for sink in list(self._sinks):
            try:
                sink[1](rec)
            except Exception as e:
                logger.warning(f"{sink[0]} logging disabled after an error: {e!r}")
                self._sinks.remove(sink)

    def close(self):
        for name, _, close in self._sinks:
            try:
                close()
            except Exception as e:
                logger.debug(f"closing {name}: {e!r}")
        self._sinks.clear()

    @property
    def active(self) -> List[str]:
        return [s[0] for s in self._sinks]
