from . import queue  # noqa: F401
