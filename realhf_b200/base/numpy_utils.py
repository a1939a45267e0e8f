"""Shape helpers for dictionaries of arrays (parity: `realhf/base/numpy_utils.py`): used when several heterogeneous arrays are
packed along one axis into a single buffer (e.g. one shared-memory segment or one message) and cut apart again."""

from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np


def shape_leq(shape1: Sequence[int], shape2: Sequence[int]) -> bool:
    """Whether an array of `shape1` fits into one of `shape2` (same rank, every dimension <=)."""
    return len(shape1) == len(shape2) and all(a <= b for a, b in zip(shape1, shape2))


def shape_union(*shapes: Sequence[int]) -> Tuple[int, ...]:
    """The smallest shape every given shape fits into (all ranks must agree)."""
    if not shapes:
        return ()
    rank = len(shapes[0])
    if any(len(s) != rank for s in shapes):
        raise ValueError(f"shapes of different rank have no union: {shapes}")
    return tuple(max(s[d] for s in shapes) for d in range(rank))


def split_to_shapes(x: np.ndarray, shapes: Dict[str, Sequence[int]], axis: int = -1) -> Dict[str, np.ndarray]:
    """Cut `x` along `axis` into consecutive pieces and give each the trailing shape of its entry: the inverse of flattening every
    array's dimensions from `axis` on and concatenating them there (the leading dimensions are shared)."""
    axis = axis % x.ndim
    out, off = {}, 0
    for k, shp in shapes.items():
        n = int(np.prod(shp[axis:])) if len(shp) > axis else 1
        piece = np.take(x, np.arange(off, off + n), axis=axis)
        out[k] = piece.reshape(*x.shape[:axis], *shp[axis:])
        off += n
    if off != x.shape[axis]:
        raise ValueError(f"shapes consume {off} entries along axis {axis}, the array has {x.shape[axis]}")
    return out
