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


# ------------------------------------------------------------------------------------------------ VMM + multicast (NVLS)


def _exchange_fds(my_fds: List[int], group, tag: str) -> List[List[int]]:
    """All-to-all of POSIX file descriptors between the (same-host) ranks of `group` over abstract unix sockets.
    Returns, per rank, the list of that rank's fds as valid descriptors in THIS process (my own entry is `my_fds`)."""
    import os
    import socket
    import threading
    import uuid
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    tok = [uuid.uuid4().hex if rank == 0 else None]
    dist.broadcast_object_list(tok, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    name = lambda r: f"\0realhf_b200_symm_{tok[0]}_{tag}_{r}"
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(name(rank))
    srv.listen(world)

    def serve():
        for _ in range(world - 1):
            conn, _ = srv.accept()
            with conn:
                socket.send_fds(conn, [b"fds"], list(my_fds))
    th = threading.Thread(target=serve, daemon=True)
    th.start()
    dist.barrier(group=group)  # every listener is up
    out: List[List[int]] = []
    for r in range(world):
        if r == rank:
            out.append(list(my_fds))
            continue
        c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        c.connect(name(r))
        _, fds, _, _ = socket.recv_fds(c, 16, len(my_fds))
        c.close()
        assert len(fds) == len(my_fds), f"expected {len(my_fds)} descriptors from rank {r}, got {len(fds)}"
        out.append(list(fds))
    th.join(timeout=60)
    srv.close()
    dist.barrier(group=group)
    return out


def multicast_supported(device=None) -> bool:
    dev = torch.device(device if device is not None else ("cuda", torch.cuda.current_device()))
    try:
        return int(lib().vmm_multicast_supported(dev.index)) == 1
    except Exception:
        return False


class VmmSymmetricBuffer(SymmetricBuffer):
    """Symmetric buffer on the CUDA VMM API with an NVSwitch multicast mapping (`mc_ptr`).

    Same interface as `SymmetricBuffer` (peer tensors, data / pad pointers, barrier, all-reduce), plus:
      * `mc_ptr`: multicast address of the data region — `multimem.ld_reduce` sums the same offset over all ranks inside the
        switch, `multimem.st` writes all replicas (0 when the platform has no multicast support);
      * `nvls_all_reduce`, and the ZeRO building blocks `reduce_scatter_` / `adam_allgather_` / `all_gather_`.
    Physical memory comes from cuMemCreate (exportable as a file descriptor); descriptors travel over unix sockets."""

    def __init__(self, nbytes: int, group=None, device=None, multicast: bool = True):
        assert dist.is_initialized()
        L = lib()
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.device(device if device is not None else ("cuda", torch.cuda.current_device()))
        dev = self.device.index
        self.pad_words = int(L.symm_pad_words())
        want_mc = multicast and self.world > 1 and multicast_supported(self.device)
        flags = [want_mc]
        allf: List = [None] * self.world
        dist.all_gather_object(allf, flags, group=group)
        want_mc = all(f[0] for f in allf)
        gran = int(L.vmm_granularity(dev, self.world if want_mc else 1))
        assert gran > 0, "cuMemGetAllocationGranularity failed"
        self.nbytes = (nbytes + 4095) // 4096 * 4096
        self.total = (self.nbytes + self.pad_words * 4 + gran - 1) // gran * gran
        self.gran = gran
        ptr, handle, fd = (int(x) for x in L.vmm_alloc(self.total, gran, dev))
        self._handles = [(ptr, handle)]
        mc_handle = mc_fd = None
        if want_mc and self.rank == 0:
            mc_handle, mc_fd = (int(x) for x in L.mc_create(self.total, self.world))
        my_fds = [fd] + ([mc_fd] if mc_fd is not None else [])
        # rank 0 ships two descriptors (memory + multicast object), the others one: pad to a common count
        if want_mc and self.rank != 0:
            my_fds = [fd, fd]
        fds = _exchange_fds(my_fds, group, tag="vmm")  # the socket names carry a fresh token per exchange (broadcast by rank 0)
        ptrs = []
        for r in range(self.world):
            if r == self.rank:
                ptrs.append(ptr)
            else:
                p, h = (int(x) for x in L.vmm_import(fds[r][0], self.total, gran, dev))
                self._handles.append((p, h))
                ptrs.append(p)
        self.mc_ptr = 0
        if want_mc:
            if self.rank != 0:
                mc_handle = int(L.mc_import(fds[0][1]))
            L.mc_add_device(mc_handle, dev)
            dist.barrier(group=group)  # every device is part of the object before anybody binds memory
            self.mc_ptr = int(L.mc_bind_and_map(mc_handle, handle, self.total, gran, dev))
            self._mc_handle = mc_handle
        for f in {f for r in range(self.world) for f in fds[r]}:  # imported (or, for my own, delivered): the mappings keep the memory alive
            try:
                L.close_fd(f)
            except Exception:
                pass
        self.peers = [L.tensor_from_ptr(p, self.total, dev) for p in ptrs]
        self.local = self.peers[self.rank]
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)
        self.data_ptrs = [int(p) for p in ptrs]
        self.pad_ptrs = [int(p) + self.nbytes for p in ptrs]
        self.counter_ptr = self.pad_ptrs[self.rank] + 4 * int(L.symm_counter_word())
        self._flip = 0

    # ---- NVLS collectives
    _DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}

    def nvls_all_reduce(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, mode: Optional[int] = None) -> torch.Tensor:
        """Sum over the group through the switch.  mode 1: every rank pulls the whole reduced tensor (`multimem.ld_reduce`);
        mode 2: reduce my slice + `multimem.st` it to everybody.  Alternates between the two halves of the data region, so no
        trailing barrier is needed in mode 1.  Graph-capturable."""
        x = x.contiguous()
        n = x.numel() * x.element_size()
        out = torch.empty_like(x) if out is None else out
        half = (self.nbytes // 2) // 1024 * 1024
        if self.mc_ptr == 0 or n % 16 != 0 or (n > half if mode != 2 else 2 * n > half):
            return self.all_reduce(x, out)
        if mode is None:
            mode = 1 if n <= (256 << 10) else 2
        base = self._flip * half
        self._flip ^= 1
        if mode == 2 and 2 * ((n + 1023) // 1024 * 1024) > half:
            mode = 1
        off_out = base + (n + 1023) // 1024 * 1024
        lib().nvls_allreduce(x, out, n, self.data_ptrs, self.pad_ptrs, self.mc_ptr, base, off_out, self.rank, self._DT[x.dtype], mode, 0,
                             self.device.index)
        return out

    def reduce_scatter_(self, byte_off: int, nbytes: int, dtype, scale: float, stats: torch.Tensor):
        """In place on this rank's copy: data[off: off+nbytes] = scale * sum over ranks; accumulates [sumsq, nonfinite] in `stats`."""
        lib().nvls_reduce_scatter(self.data_ptrs, self.pad_ptrs, self.mc_ptr, byte_off, nbytes, self._DT[dtype], scale, stats, self.rank)

    def all_gather_(self, byte_off: int, nbytes: int, lead_barrier: bool = True):
        """Broadcast data[off: off+nbytes] of this rank to the same offset of every rank (`multimem.st`)."""
        lib().nvls_allgather(self.data_ptrs, self.pad_ptrs, self.mc_ptr, byte_off, nbytes, lead_barrier, self.rank, self.device.index)
