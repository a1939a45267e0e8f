"""Recover bookkeeping (parity: `realhf/base/recover.py:12-54`).  Beyond the reference, model workers also save
optimizer / LR-scheduler state (`TrainBackend.save`), and interfaces expose KL-controller / value-normaliser state."""

import dataclasses
import os
import pickle
from typing import Hashable, List, Optional

from realhf_b200.base import constants


@dataclasses.dataclass
class StepInfo:
    epoch: int = 0
    epoch_step: int = 0
    global_step: int = 0


@dataclasses.dataclass
class RecoverInfo:
    recover_start: StepInfo
    last_step_info: StepInfo
    hash_vals_to_ignore: List[Hashable] = dataclasses.field(default_factory=list)


def _path(exp: str, trial: str) -> str:
    d = os.path.join(constants.RECOVER_ROOT, exp, trial)
    os.makedirs(d, exist_ok=True)
    return os.path.join(d, "recover_info.pkl")


def dump_recover_info(info: RecoverInfo, exp: str, trial: str):
    tmp = _path(exp, trial) + ".tmp"
    with open(tmp, "wb") as f:
        pickle.dump(info, f)
    os.replace(tmp, _path(exp, trial))


def load_recover_info(exp: str, trial: str) -> Optional[RecoverInfo]:
    p = _path(exp, trial)
    if not os.path.exists(p):
        return None
    with open(p, "rb") as f:
        return pickle.load(f)
