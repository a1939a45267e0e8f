"""Running several coroutines of one worker together.  Parity: `realhf/base/asyncio_utils.py`
(setup_run_until_complete / teardown / raise_asyncio_exception).

The master worker is a set of long-lived coroutines (request pump, one DFG walker per MFC, control endpoint).  If one of them
dies, the others would wait for it forever: `gather_or_raise` turns the first exception into the caller's exception and
cancels the rest, which is what makes a failing MFC stop the trial instead of hanging it."""

from __future__ import annotations

import asyncio
from typing import Awaitable, Iterable, List, Optional


async def gather_or_raise(aws: Iterable[Awaitable], cancel_timeout: float = 5.0) -> List:
    """Like `asyncio.gather`, but the first exception cancels every other task (and waits for the cancellations) before it
    propagates."""
    tasks = [asyncio.ensure_future(a) for a in aws]
    try:
        done, pending = await asyncio.wait(tasks, return_when=asyncio.FIRST_EXCEPTION)
        err = next((t.exception() for t in done if not t.cancelled() and t.exception() is not None), None)
        if err is not None:
            await cancel_all(pending, cancel_timeout)
            raise err
        if pending:  # none failed yet: FIRST_EXCEPTION returned because everything finished
            await asyncio.wait(pending)
        return [t.result() for t in tasks]
    except asyncio.CancelledError:
        await cancel_all(tasks, cancel_timeout)
        raise


async def cancel_all(tasks: Iterable["asyncio.Future"], timeout: Optional[float] = 5.0):
    tasks = [t for t in tasks if not t.done()]
    for t in tasks:
        t.cancel()
    if tasks:
        await asyncio.wait(tasks, timeout=timeout)


def raise_first_exception(tasks: Iterable["asyncio.Future"]):
    """Poll-style check for worker loops that keep tasks running across iterations."""
    for t in tasks:
        if t.done() and not t.cancelled() and t.exception() is not None:
            raise t.exception()
