"""Symmetric (peer-mapped) memory over CUDA IPC and the collectives built on it.

Every rank of a process group allocates a buffer of the same size; IPC handles are exchanged once over the control
plane (`all_gather_object`) and every peer's buffer is mapped into this process, so kernels can load / store peer
memory directly over NVLink / NVSwitch.  The reference ships the same handles for its dormant custom all-reduce
(`parallelism/model_parallel/custom_all_reduce.py:206-236`); here the substrate also carries parameter reallocation
(peer stores of the destination layout) and the fused TP GEMM kernels.

All ranks must be on one node and see all GPUs (no per-process CUDA_VISIBLE_DEVICES isolation).
"""

from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from realhf_b200.ops import lib


class SymmetricBuffer:
    """`nbytes` of data region + a signal pad on every rank of `group`, all mapped everywhere."""

    def __init__(self, nbytes: int, group=None, device=None):
        assert dist.is_initialized()
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.device(device if device is not None else ("cuda", torch.cuda.current_device()))
        self.nbytes = (nbytes + 4095) // 4096 * 4096
        self.pad_words = int(lib().symm_pad_words())
        # one raw cudaMalloc holds [data | pad], zeroed so that epoch counters start equal on all ranks
        total = self.nbytes + self.pad_words * 4
        self.local, handle = lib().symm_alloc(total, self.device.index)
        gathered: List = [None] * self.world
        dist.all_gather_object(gathered, handle, group=group)
        self.peers: List[torch.Tensor] = []
        for r, h in enumerate(gathered):
            # the peer buffer is mapped into THIS device's context (lazy peer access over NVLink)
            self.peers.append(self.local if r == self.rank else lib().symm_open(h, total, self.device.index))
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)
        self.data_ptrs = [int(p.data_ptr()) for p in self.peers]
        self.pad_ptrs = [int(p.data_ptr()) + self.nbytes for p in self.peers]
        self.counter_ptr = self.pad_ptrs[self.rank] + 4 * int(lib().symm_counter_word())

    def data(self, rank: Optional[int] = None, dtype=torch.uint8) -> torch.Tensor:
        t = self.peers[self.rank if rank is None else rank][: self.nbytes]
        return t.view(dtype)

    def pad(self, rank: Optional[int] = None) -> torch.Tensor:
        """int32 view of a rank's signal pad (words >= 960 are free for host-orchestrated protocols)."""
        return self.peers[self.rank if rank is None else rank][self.nbytes:].view(torch.int32)

    def barrier(self):
        lib().symm_barrier(self.local, self.data_ptrs, self.pad_ptrs, self.rank)

    def all_reduce(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, algo: Optional[int] = None) -> torch.Tensor:
        """Sum over the group.  Capturable in a CUDA graph (epochs advance on the device)."""
        x = x.contiguous()
        n = x.numel() * x.element_size()
        out = torch.empty_like(x) if out is None else out
        if algo is None:  # same crossover rule as the reference table: small or 2 ranks -> one-shot
            algo = 1 if (self.world == 2 or n <= (512 << 10 if self.world <= 4 else 256 << 10)) else 2
        need = n if algo == 1 else 2 * ((n + 1023) // 1024 * 1024)
        if need > self.nbytes or n % 16 != 0:
            dist.all_reduce(x, group=self.group)
            return x
        lib().symm_allreduce(x, out, self.data_ptrs, self.pad_ptrs, self.rank, algo)
        return out
