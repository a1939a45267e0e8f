"""Frequency controls (parity: `realhf/base/timeutil.py`: FrequencyControl, EpochStepTimeFreqCtl)."""

import dataclasses
import time
from typing import Optional


class FrequencyControl:
    """Fires when `frequency_steps` calls or `frequency_seconds` seconds have passed since the last fire."""

    def __init__(self, frequency_seconds: Optional[float] = None, frequency_steps: Optional[int] = None,
                 initial_value: bool = False):
        self.secs, self.steps = frequency_seconds, frequency_steps
        self._last = time.monotonic()
        self._count = 0
        self._initial = initial_value

    def check(self, steps: int = 1) -> bool:
        if self._initial:
            self._initial = False
            self._last, self._count = time.monotonic(), 0
            return True
        self._count += steps
        fire = False
        if self.steps is not None and self._count >= self.steps:
            fire = True
        if self.secs is not None and time.monotonic() - self._last >= self.secs:
            fire = True
        if fire:
            self._last, self._count = time.monotonic(), 0
        return fire

    def state_dict(self):
        return dict(count=self._count, elapsed=time.monotonic() - self._last)

    def load_state_dict(self, sd):
        self._count = sd["count"]
        self._last = time.monotonic() - sd["elapsed"]


@dataclasses.dataclass
class EpochStepTimeFreqCtl:
    freq_epoch: Optional[int] = None
    freq_step: Optional[int] = None
    freq_sec: Optional[float] = None

    def __post_init__(self):
        self._e = FrequencyControl(frequency_steps=self.freq_epoch)
        self._s = FrequencyControl(frequency_steps=self.freq_step)
        self._t = FrequencyControl(frequency_seconds=self.freq_sec)

    def check(self, epochs: int, steps: int) -> bool:
        e = self._e.check(epochs) if (self.freq_epoch is not None and epochs) else False
        s = self._s.check(steps) if self.freq_step is not None else False
        t = self._t.check() if self.freq_sec is not None else False
        return e or s or t
