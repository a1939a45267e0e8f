"""Minimal stand-in for `colorlog` (not in the offline image) so that the reference's `realhf.base.logging` imports.
Colours are dropped; the format string's `%(log_color)s` placeholder resolves to an empty string."""
import logging


class ColoredFormatter(logging.Formatter):
    def __init__(self, fmt=None, datefmt=None, style="%", log_colors=None, reset=True, secondary_log_colors=None, **kw):
        super().__init__(fmt, datefmt, style)

    def format(self, record):
        record.log_color = ""
        record.reset = ""
        return super().format(record)


class StreamHandler(logging.StreamHandler):
    pass
