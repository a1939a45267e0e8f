"""Metadata-only sequence buffer of the master worker.

Parity: `realhf/system/buffer.py` (AsyncIOSequenceBuffer :117-347): slots hold per-sample metadata (which keys
exist, their sequence lengths), an MFC can take a batch once `n_seqs` samples have all its input keys and have not
been consumed by it yet, and a slot is freed when every MFC has read it.
"""

from __future__ import annotations

import asyncio
import dataclasses
from typing import Dict, Hashable, List, Set

from realhf_b200.api.data import SequenceSample
from realhf_b200.api.dfg import MFCDef


@dataclasses.dataclass
class _Slot:
    sample: SequenceSample          # metadata-only, one item
    birth: int
    consumed_by: Set[str] = dataclasses.field(default_factory=set)


class AsyncIOSequenceBuffer:
    def __init__(self, rpcs: List[MFCDef], max_size: int = 1_000_000):
        self.rpcs = rpcs
        self.max_size = max_size
        self._slots: Dict[Hashable, _Slot] = {}
        self._counter = 0
        self._cond = asyncio.Condition()

    @property
    def size(self) -> int:
        return len(self._slots)

    def __contains__(self, sample_id: Hashable) -> bool:
        return sample_id in self._slots

    def n_ready_for(self, rpc: MFCDef) -> int:
        keys = set(rpc.input_keys)
        return sum(1 for s in self._slots.values() if rpc.name not in s.consumed_by and keys.issubset(s.sample.keys))

    async def put_batch(self, samples: List[SequenceSample]):
        async with self._cond:
            if len(self._slots) + len(samples) > self.max_size:
                raise RuntimeError("sequence buffer overflow")
            for s in samples:
                assert s.bs == 1 and s.ids[0] not in self._slots, s.ids
                self._slots[s.ids[0]] = _Slot(s, self._counter)
                self._counter += 1
            self._cond.notify_all()

    async def amend_batch(self, ids: List[Hashable], new: List[SequenceSample]):
        async with self._cond:
            for i, s in zip(ids, new):
                self._slots[i].sample.update_(s)
            self._cond.notify_all()

    async def get_batch_for_rpc(self, rpc: MFCDef):
        """Oldest `n_seqs` samples that carry all input keys of `rpc` and were not consumed by it."""
        keys = set(rpc.input_keys)
        async with self._cond:
            while True:
                ready = [(s.birth, i) for i, s in self._slots.items()
                         if rpc.name not in s.consumed_by and keys.issubset(s.sample.keys)]
                if len(ready) >= rpc.n_seqs:
                    break
                await self._cond.wait()
            ready.sort()
            ids = [i for _, i in ready[: rpc.n_seqs]]
            for i in ids:
                self._slots[i].consumed_by.add(rpc.name)
            with SequenceSample.disable_validation():
                sample = SequenceSample.gather([self._slots[i].sample for i in ids], keys=rpc.input_keys)
            return ids, sample

    def pop_fully_consumed(self) -> List[Hashable]:
        """Ids whose every consumer MFC has read them (their tensors can be dropped on the workers)."""
        all_names = {r.name for r in self.rpcs}
        done = [i for i, s in self._slots.items() if s.consumed_by >= all_names]
        for i in done:
            del self._slots[i]
        return done
