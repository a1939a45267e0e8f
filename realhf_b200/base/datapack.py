"""Balanced partitioning / batching of variable-length sequences.

Native implementation in `ops/csrc/host/datapack.cpp` (the reference uses numba,
`realhf/base/datapack.py`); the Python functions below implement the *same algorithm* and are used
when the extension is not built, so both give identical partitions.
"""

from __future__ import annotations

import bisect
import heapq
import itertools
from typing import Any, List, Sequence, Tuple

import numpy as np

from realhf_b200.ops import host


def flat2d(arr: List[List[Any]]) -> List[Any]:
    return list(itertools.chain.from_iterable(arr))


def _greedy_left(pre, k, min_size, cap):
    n = len(pre) - 1
    start, ends = 0, []
    for p in range(k):
        must_leave = (k - 1 - p) * min_size
        lo_end = start + min_size
        if lo_end > n - must_leave or pre[lo_end] - pre[start] > cap:
            return None
        end = bisect.bisect_right(pre, pre[start] + cap, lo_end, n - must_leave + 1) - 1
        if p == k - 1:
            if end < n:
                return None
            end = n
        ends.append(end)
        start = end
    return ends


def _partition_balanced_py(nums: Sequence[int], k: int, min_size: int = 1) -> List[int]:
    n = len(nums)
    pre = [0] * (n + 1)
    for i, x in enumerate(nums):
        pre[i + 1] = pre[i] + int(x)
    lo, hi = 0, pre[n]
    while lo < hi:
        mid = lo + (hi - lo) // 2
        if _greedy_left(pre, k, min_size, mid) is not None:
            hi = mid
        else:
            lo = mid + 1
    cap = lo
    latest = _greedy_left(pre, k, min_size, cap)
    earliest = [0] * k
    end = n
    for p in range(k - 1, -1, -1):
        earliest[p] = end
        must_leave = p * min_size
        hi_start = end - min_size
        start = bisect.bisect_left(pre, pre[end] - cap, must_leave, hi_start + 1)
        start = min(start, hi_start)
        if p == 0:
            start = 0
        end = start
    bounds = [0] * (k + 1)
    bounds[k] = n
    prev = 0
    for p in range(k - 1):
        wlo = max(earliest[p], prev + min_size)
        whi = min(latest[p], n - (k - 1 - p) * min_size)
        whi = min(whi, bisect.bisect_right(pre, pre[prev] + cap) - 1)
        whi = max(whi, wlo)
        ideal = pre[n] * (p + 1) / k
        pos = bisect.bisect_left(pre, int(ideal), wlo, whi + 1)
        pos = min(pos, whi)
        if pos > wlo and abs(pre[pos - 1] - ideal) <= abs(pre[pos] - ideal):
            pos -= 1
        bounds[p + 1] = pos
        prev = pos
    return bounds


def partition_balanced(nums: Sequence[int], k: int, min_size: int = 1) -> List[int]:
    """k+1 boundaries of the contiguous k-way split with the smallest possible largest part."""
    nums = [int(x) for x in nums]
    h = host()
    if h is not None:
        return list(h.partition_balanced(nums, k, min_size))
    return _partition_balanced_py(nums, k, min_size)


def min_abs_diff_partition(arr, k: int, min_size: int = 1) -> List[Tuple[int, int]]:
    """[(start, end)] * k — name kept from the reference API."""
    arr = np.asarray(arr)
    if arr.ndim != 1:
        raise ValueError(f"the array to partition must be 1-D, got shape {arr.shape}")
    if len(arr) < k * min_size:
        raise ValueError(f"cannot split {len(arr)} items into {k} parts of at least {min_size}")
    b = partition_balanced(arr.tolist(), k, min_size)
    parts = [(b[i], b[i + 1]) for i in range(k)]
    assert all(e > s for s, e in parts), (arr, k, parts)
    return parts


def _reorder_py(seqlens: Sequence[int], n_seqs_per_batch: int):
    n = len(seqlens)
    n_bins = (n + n_seqs_per_batch - 1) // n_seqs_per_batch
    order = sorted(range(n), key=lambda i: -seqlens[i])
    heap = [(0, b) for b in range(n_bins)]
    heapq.heapify(heap)
    bins = [[] for _ in range(n_bins)]
    tokens = [0] * n_bins
    for idx in order:
        t, b = heapq.heappop(heap)
        bins[b].append(idx)
        tokens[b] = t + seqlens[idx]
        if len(bins[b]) < n_seqs_per_batch:
            heapq.heappush(heap, (tokens[b], b))
    bin_order = sorted(range(n_bins), key=lambda b: -tokens[b])
    out = [i for b in bin_order for i in bins[b]]
    return out, (max(tokens) - min(tokens) if n_bins else 0)


def reorder_to_balanced_batches(seqlens, n_seqs_per_batch: int) -> Tuple[np.ndarray, int]:
    """Permutation such that consecutive groups of `n_seqs_per_batch` have balanced token counts."""
    lens = [int(x) for x in seqlens]
    h = host()
    if h is not None:
        out, diff = h.reorder_to_balanced_batches(lens, n_seqs_per_batch)
    else:
        out, diff = _reorder_py(lens, n_seqs_per_batch)
    return np.asarray(out, dtype=np.int64), int(diff)


def merge_intervals(iv: List[Tuple[int, int]]) -> List[Tuple[int, int]]:
    h = host()
    if h is not None:
        return [tuple(x) for x in h.merge_intervals([(int(a), int(b)) for a, b in iv])]
    out: List[Tuple[int, int]] = []
    for a, b in iv:
        if out and out[-1][1] == a:
            out[-1] = (out[-1][0], b)
        else:
            out.append((a, b))
    return out
