"""Seeding.  `set_random_seed(seed, offset)` seeds the process-wide generators with `seed + offset` (workers pass their index so
data order / rank-local dropout differ per rank) and remembers the un-offset experiment seed, from which `derive_seed` builds
streams that must be IDENTICAL on a chosen set of ranks (e.g. all tensor-parallel ranks of one model replica).

Parity: `realhf/base/seeding.py` + the purpose of the TP-aware RNG tracker in `impl/model/utils/random.py:76-286`."""

import random
import zlib

import numpy as np
import torch

_BASE_SEED = [1]


def set_random_seed(seed: int, offset: int = 0):
    _BASE_SEED[0] = int(seed)
    s = int(seed) + int(offset)
    random.seed(s)
    np.random.seed(s % (2 ** 32))
    torch.manual_seed(s)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(s)


def base_seed() -> int:
    return _BASE_SEED[0]


def derive_seed(*tags) -> int:
    """A 62-bit seed that depends only on the experiment seed and `tags` (strings / ints), not on the calling rank."""
    h = zlib.crc32(repr((_BASE_SEED[0],) + tuple(tags)).encode())
    h2 = zlib.crc32(repr(tuple(tags) + (_BASE_SEED[0], "x")).encode())
    return ((h << 31) ^ h2) & ((1 << 62) - 1)
