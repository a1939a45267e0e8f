"""`ray.util.queue.Queue` stand-in: a multiprocessing queue with Ray's method names (picklable into spawned tasks)."""

import multiprocessing as mp
import queue as _q

Empty = _q.Empty
Full = _q.Full


class Queue:
    def __init__(self, maxsize: int = 0, actor_options=None):
        self._q = mp.get_context("spawn").Queue(maxsize)

    def put(self, item, block=True, timeout=None):
        self._q.put(item, block, timeout)

    def get(self, block=True, timeout=None):
        return self._q.get(block, timeout)

    def put_nowait(self, item):
        self._q.put_nowait(item)

    def get_nowait(self):
        return self._q.get_nowait()

    def empty(self):
        return self._q.empty()

    def qsize(self):
        return self._q.qsize()

    def shutdown(self):
        self._q.close()
