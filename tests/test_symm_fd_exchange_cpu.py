"""File-descriptor exchange used by the VMM symmetric memory (cuMemExportToShareableHandle fds travel between the ranks over
abstract unix sockets): every rank must end up with working duplicates of every peer's descriptors."""
import os
import sys

import pytest

pytestmark = pytest.mark.distributed
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _worker(rank, world):
    from realhf_b200.parallel.symm_mem import _exchange_fds
    r1, w1 = os.pipe()
    r2, w2 = os.pipe()
    os.write(w1, f"a{rank}".encode())
    os.write(w2, f"b{rank}".encode())
    got = []
    for rnd in range(2):  # two exchanges in a row (several symmetric buffers per process): names must not collide
        fds = _exchange_fds([r1, r2], None, tag="t")
        assert len(fds) == world and fds[rank] == [r1, r2]
        got.append([[int(f) for f in fds[p]] for p in range(world)])
    # read through the received duplicates of the LAST exchange: ring order so that every pipe is read exactly once
    peer = (rank + 1) % world
    return os.read(got[-1][peer][0], 16).decode(), os.read(got[-1][peer][1], 16).decode()


@pytest.mark.parametrize("world", [2, 3])
def test_fd_exchange_over_unix_sockets(world):
    from realhf_b200.base.testing import run_distributed
    res = run_distributed(_worker, world)
    for rank, (a, b) in enumerate(res):
        peer = (rank + 1) % world
        assert (a, b) == (f"a{peer}", f"b{peer}"), res
