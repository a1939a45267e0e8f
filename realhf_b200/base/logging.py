"""Logging helpers (parity: `realhf/base/logging.py`: plain / colored / benchmark / system flavours)."""

import logging as _logging
import os
import sys

_FMT = "%(asctime)s.%(msecs)03d %(name)s %(levelname)s: %(message)s"
_DATE = "%Y%m%d-%H:%M:%S"
_COLORS = {"benchmark": "\033[36m", "system": "\033[35m", "colored": "\033[32m"}
_configured = False


class _ColorFormatter(_logging.Formatter):
    def __init__(self, color: str):
        super().__init__(_FMT, _DATE)
        self.color = color

    def format(self, record):
        s = super().format(record)
        return f"{self.color}{s}\033[0m" if sys.stderr.isatty() else s


def getLogger(name: str = None, type_: str = "plain") -> _logging.Logger:
    global _configured
    if not _configured:
        _logging.basicConfig(level=os.environ.get("REAL_LOG_LEVEL", "INFO"), format=_FMT, datefmt=_DATE, stream=sys.stderr)
        _configured = True
    lg = _logging.getLogger(name)
    if type_ in _COLORS and not any(isinstance(h.formatter, _ColorFormatter) for h in lg.handlers):
        h = _logging.StreamHandler(sys.stderr)
        h.setFormatter(_ColorFormatter(_COLORS[type_]))
        lg.addHandler(h)
        lg.propagate = False
    return lg
