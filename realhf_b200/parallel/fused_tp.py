"""Fused tensor-parallel GEMM + collective kernels over symmetric memory.

The three places where a projection is adjacent to a TP collective:

  * `gemm_rs`  row-parallel linear with sequence parallelism: y = reduce_scatter_tokens(x @ W^T).  The tcgen05 GEMM's
    epilogue stores each output row directly into the owning rank's inbox slab over NVLink and bumps that rank's
    arrival counter while later tiles are still on the tensor cores; a tail kernel sums the slabs.
  * `ag_gemm`  column-parallel linear with sequence parallelism: y = all_gather_tokens(x) @ W^T.  The GEMM's TMA
    producer loads A row-blocks straight from the owning rank's staging buffer (peer memory), so no gathered copy of
    the activations is ever written.
  * `gemm_ar`  row-parallel linear without SP (decode): y = all_reduce(x @ W^T) with the one-shot / two-shot
    peer-memory all-reduce, capturable in CUDA graphs.

`FusedTP` is attached to a `ParallelContext` as `ctx.symm`; `parallel/tp.py` routes `col_linear` / `row_linear`
through it when shapes qualify, else falls back to GEMM + NCCL.  Backward passes use the mirrored fused op
(d(gemm_rs) = ag_gemm of the gradient and vice versa); weight gradients gather with NCCL (cold relative to the GEMMs).
All fused calls of one rank must be issued on one stream (inbox / staging regions are double-buffered by call parity).
"""

from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist

from realhf_b200.ops import functional as OF
from realhf_b200.ops import lib
from realhf_b200.parallel.symm_mem import SymmetricBuffer, VmmSymmetricBuffer, multicast_supported


class FusedTP:
    def __init__(self, ctx, max_tokens: int, max_features: int, device=None):
        """max_tokens: largest gathered token count T of a fused call; max_features: largest N (gemm_rs output width) / K (ag_gemm)."""
        self.ctx = ctx
        self.group = ctx.tp_group
        self.world = ctx.tp_size
        self.rank = ctx.tp_rank
        self.device = torch.device(device if device is not None else ("cuda", torch.cuda.current_device()))
        self.max_tokens, self.max_features = max_tokens, max_features
        inbox = max_tokens * max_features * 2                 # world slabs of [T/world, N] bf16
        stage = (max_tokens // self.world + 128) * max_features * 2
        self.inbox_off = [0, inbox]
        self.stage_off = [2 * inbox, 2 * inbox + stage]
        self.ar_off = 2 * inbox + 2 * stage
        ar_bytes = 64 << 20
        # VMM allocation with an NVSwitch multicast mapping when the platform has one (multimem.ld_reduce all-reduce fused with
        # the residual add + RMSNorm of the decode path); CUDA-IPC buffer with peer loads otherwise
        use_mc = os.environ.get("REAL_TP_NVLS", "1") == "1" and multicast_supported(self.device)
        flags = [None] * self.world
        dist.all_gather_object(flags, bool(use_mc), group=self.group)
        if all(flags):
            self.symm = VmmSymmetricBuffer(self.ar_off + ar_bytes, group=self.group, device=self.device)
        else:
            self.symm = SymmetricBuffer(self.ar_off + ar_bytes, group=self.group, device=self.device)
        self.nvls = getattr(self.symm, "mc_ptr", 0) != 0
        self._rs_calls = 0
        self._ag_calls = 0
        self.sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        counter0 = int(lib().symm_counter_word())
        self._counter_ptrs = [[p + 4 * (counter0 + par) for p in self.symm.pad_ptrs] for par in (0, 1)]
        self._calls_ptr = [self.symm.pad_ptrs[self.rank] + 4 * int(lib().symm_calls_word(par)) for par in (0, 1)]
        # a dedicated all-reduce view over the tail of the buffer
        self._ar_data_ptrs = [p + self.ar_off for p in self.symm.data_ptrs]
        self._ar_bytes = ar_bytes
        self._ar_calls = 0
        # all-gather -> GEMM: gathered activations are assembled in a local buffer by the copy engines on a side stream
        self._ag_stream = torch.cuda.Stream(self.device)
        self._ag_flags = torch.zeros(4096, dtype=torch.int32, device=self.device)
        self._ag_epoch = 0
        self._ag_bufs = {}
        self._rs_state = {}
        self._epoch_ring = [torch.zeros(1, dtype=torch.int32, device=self.device) for _ in range(64)]

    # ------------------------------------------------------------------ eligibility
    def _ok(self, x, w) -> bool:
        return (x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1
                and w.stride(1) == 1 and x.shape[1] % 8 == 0 and w.stride(0) % 8 == 0)

    def can_gemm_reduce(self, x, w) -> bool:
        return self._ok(x, w) and x.shape[0] % self.world == 0 and x.shape[0] <= self.max_tokens and w.shape[0] <= self.max_features \
            and w.shape[0] % 8 == 0

    def can_ag_gemm(self, x, w) -> bool:
        return self._ok(x, w) and x.shape[0] % 128 == 0 and x.shape[0] * self.world <= self.max_tokens and x.shape[1] <= self.max_features \
            and x.shape[0] * self.world > 128

    # ------------------------------------------------------------------ raw fused ops (no autograd)
    def _gemm_rs_raw(self, x, w, b_mn: bool):
        """y[T/t, N] = reduce_scatter_tokens(x @ w^T).  The CTA-pair GEMM computes the row-blocks owned by the other ranks
        FIRST (m-tile rotation) and bumps a completion counter per row-block; a side stream waits on those counters and
        ships every finished block into the owner's inbox slab with the copy engines while the tensor cores continue with
        the remaining blocks (own rows last); a 4-byte peer copy per source signals arrival, and the owner sums its own
        partial with the received slabs.  (v1 -- peer stores from the GEMM epilogue, 16 bytes per lane over NVLink, plus
        a reduce kernel -- is kept as `_gemm_rs_raw_v1`; it measured 25% slower than GEMM + NCCL reduce-scatter.)"""
        if os.environ.get("REAL_FUSED_RS_V1", "0") == "1":
            return self._gemm_rs_raw_v1(x, w, b_mn)
        world, me = self.world, self.rank
        T = x.shape[0]
        N = w.shape[1] if b_mn else w.shape[0]
        rows = T // world
        par = self._rs_calls & 1
        self._rs_calls += 1
        ep = self._rs_calls
        chunk = next((c for c in (1024, 512, 256, 128) if rows % c == 0), None)
        if chunk is None or (rows // chunk) * world > 2048:
            return self._gemm_rs_raw_v1(x, w, b_mn)
        n_chunk = rows // chunk
        key = (T, N, chunk)
        st = self._rs_state.get(key)
        if st is None:
            st = self._rs_state[key] = dict(done=torch.zeros(n_chunk * world, dtype=torch.int32, device=self.device), calls=0)
        st["calls"] += 1
        per_call = ((N + 255) // 256 if N > 128 else 1) * (chunk // 128) * 4   # epilogue-warp arrivals per row-block and call
        target = st["calls"] * per_call
        main = torch.cuda.current_stream(self.device)
        self.symm.barrier()  # every rank has consumed the inbox of the previous call before anyone overwrites it
        self._ag_stream.wait_stream(main)
        y = lib().gemm_gated(x, w, b_mn, None, 0, chunk, ((me + 1) % world) * rows, self.sms, st["done"])
        slab = rows * N * 2
        with torch.cuda.stream(self._ag_stream):
            for d in range(1, world):
                dst = (me + d) % world
                inbox = self.symm.data(dst)[self.inbox_off[par] + me * slab: self.inbox_off[par] + (me + 1) * slab].view(torch.bfloat16).view(rows, N)
                for c in range(n_chunk):
                    lib().spin_wait(st["done"], dst * n_chunk + c, target)
                    inbox[c * chunk:(c + 1) * chunk].copy_(y[dst * rows + c * chunk: dst * rows + (c + 1) * chunk], non_blocking=True)
                self.symm.pad(dst)[960 + par * 8 + me: 960 + par * 8 + me + 1].copy_(self._epoch_src_for(ep), non_blocking=True)
        out = y[me * rows:(me + 1) * rows]
        mine = self.symm.data()[self.inbox_off[par]: self.inbox_off[par] + world * slab].view(torch.bfloat16).view(world, rows, N)
        for d in range(1, world):
            src = (me + d) % world
            lib().spin_wait(self.symm.pad(), 960 + par * 8 + src, ep)
            out = out + mine[src]
        main.wait_stream(self._ag_stream)
        return out

    def _epoch_src_for(self, ep: int) -> torch.Tensor:
        """A device int32 holding `ep`, alive until the copy that reads it has run (ring of 64 scalars)."""
        t = self._epoch_ring[ep % 64]
        t.fill_(ep)
        return t

    def _gemm_rs_raw_v1(self, x, w, b_mn: bool):
        par = self._rs_calls & 1
        self._rs_calls += 1
        inbox = [p + self.inbox_off[par] for p in self.symm.data_ptrs]
        return lib().gemm_rs(x, w, b_mn, inbox, self._counter_ptrs[par], self._calls_ptr[par], self.rank, self.sms)

    def _ag_gemm_raw(self, x_local, w, b_mn: bool):
        """y = all_gather_tokens(x_local) @ w^T.  Every rank stages its slice in symmetric memory; the peers' slices are
        pulled over NVLink by the copy engines on a side stream, chunk by chunk, each chunk followed by a flag write; the
        CTA-pair GEMM starts immediately on the local rows (m-tile order rotated to them) and its TMA producers wait on the
        chunk flags before touching rows that are still in flight -- the transfer is hidden behind the tensor cores
        without spending SMs on communication.  (The first version let TMA read A straight from peer memory: every
        n-tile re-read its A block over NVLink and it ran 4x slower than all-gather + GEMM.)"""
        par = self._ag_calls & 1
        self._ag_calls += 1
        rows, K = x_local.shape
        world, me = self.world, self.rank
        stage = self.symm.data()[self.stage_off[par]: self.stage_off[par] + rows * K * 2].view(torch.bfloat16).view(rows, K)
        stage.copy_(x_local)
        self.symm.barrier()  # every rank's slice is staged (and visible) before anyone pulls it
        key = (rows * world, K)
        G = self._ag_bufs.get(key)
        if G is None:
            G = self._ag_bufs[key] = torch.empty(rows * world, K, dtype=torch.bfloat16, device=self.device)
        chunk = 1024 if rows % 1024 == 0 else (512 if rows % 512 == 0 else (256 if rows % 256 == 0 else 128))
        n_chunk = rows // chunk
        if n_chunk * world > self._ag_flags.numel():
            chunk, n_chunk = rows, 1
        self._ag_epoch += 1
        ep = self._ag_epoch
        main = torch.cuda.current_stream(self.device)
        flags = self._ag_flags
        # local rows: plain device copy on the compute stream, flagged ready before the GEMM is launched
        G[me * rows:(me + 1) * rows].copy_(x_local)
        flags[me * n_chunk:(me + 1) * n_chunk].fill_(ep)
        self._ag_stream.wait_stream(main)
        with torch.cuda.stream(self._ag_stream):
            for d in range(1, world):
                src = (me + d) % world
                peer = self.symm.data(src)[self.stage_off[par]: self.stage_off[par] + rows * K * 2].view(torch.bfloat16).view(rows, K)
                for c in range(n_chunk):
                    G[src * rows + c * chunk: src * rows + (c + 1) * chunk].copy_(peer[c * chunk:(c + 1) * chunk], non_blocking=True)
                    flags[src * n_chunk + c: src * n_chunk + c + 1].fill_(ep)
        y = lib().gemm_gated(G, w, b_mn, flags, ep, chunk, me * rows, self.sms, None)
        main.wait_stream(self._ag_stream)
        return y

    def all_reduce(self, x: torch.Tensor) -> torch.Tensor:
        n = x.numel() * x.element_size()
        if self.nvls and x.dtype in (torch.bfloat16, torch.float16) and x.dim() >= 1 and n % 16 == 0 and x.is_contiguous() \
                and (x.shape[-1] * x.element_size()) % 16 == 0 and x.shape[-1] % 8 == 0:
            x2 = x.reshape(-1, x.shape[-1])
            y = self.symm_out(x2.shape[0], x2.shape[1], x.dtype)
            if y is not None:
                y.copy_(x2)
                return self.all_reduce_symm(y).view(x.shape)
        algo = 1 if (self.world == 2 or n <= (512 << 10 if self.world <= 4 else 256 << 10)) else 2
        need = n if algo == 1 else 2 * ((n + 1023) // 1024 * 1024)
        if need > self._ar_bytes or n % 16 != 0 or not x.is_contiguous():
            dist.all_reduce(x, group=self.group)
            return x
        out = torch.empty_like(x)
        lib().symm_allreduce(x, out, self._ar_data_ptrs, self.symm.pad_ptrs, self.rank, algo)
        return out

    def symm_out(self, rows: int, cols: int, dtype) -> Optional[torch.Tensor]:
        """A [rows, cols] view inside this rank's all-reduce region (regions alternate per call) for a producer that
        writes its partial result straight into symmetric memory; pass it to `all_reduce_symm`."""
        nbytes = rows * cols * torch.tensor([], dtype=dtype).element_size()
        half = self._ar_bytes // 2
        if nbytes > half or nbytes % 16 != 0:
            return None
        off = self.ar_off + (self._ar_calls & 1) * half
        self._ar_calls += 1
        return self.symm.data()[off: off + nbytes].view(dtype).view(rows, cols)

    def all_reduce_symm(self, y_sym: torch.Tensor) -> torch.Tensor:
        """Sum over ranks of a tensor obtained from `symm_out` (one barrier, no staging copy).  With a multicast mapping the sum
        is formed inside the switch (`multimem.ld_reduce`), else by loads from every peer."""
        if self.nvls and y_sym.dtype in (torch.bfloat16, torch.float16):
            return self.ar_add_rmsnorm(y_sym, None, None)
        out = torch.empty(y_sym.shape, dtype=y_sym.dtype, device=y_sym.device)
        lib().symm_allreduce(y_sym, out, self.symm.data_ptrs, self.symm.pad_ptrs, self.rank, 3)
        return out

    def gemm_partial(self, x: torch.Tensor, w: torch.Tensor) -> Optional[torch.Tensor]:
        """This rank's partial product x @ w^T written straight into symmetric memory ([rows, N] view), to be reduced by
        `ar_add_rmsnorm` / `all_reduce_symm`; None when the shapes do not qualify."""
        from realhf_b200.ops import gemm as G
        x2 = x.reshape(-1, x.shape[-1])
        if not (G.supported(x2, w) and x2.stride(-1) == 1 and x2.stride(0) % 8 == 0 and x2.data_ptr() % 16 == 0):
            return None
        y_sym = self.symm_out(x2.shape[0], w.shape[0], x.dtype)
        if y_sym is None:
            return None
        G.gemm(x2, w, out=y_sym)
        return y_sym

    def ar_add_rmsnorm(self, y_sym: torch.Tensor, residual: Optional[torch.Tensor], w: Optional[torch.Tensor], eps: float = 1e-5,
                       w_offset: float = 0.0):
        """ONE kernel for the layer boundary of tensor-parallel decode: x_new = sum_ranks(y_sym) + residual,
        h = rmsnorm(x_new) * (w + w_offset).  Returns (h, x_new), or x_new alone when `w` is None."""
        assert self.nvls
        rows, H = y_sym.shape
        off = y_sym.data_ptr() - self.symm.data_ptrs[self.rank]
        res = lib().nvls_ar_add_rmsnorm(self.symm.data_ptrs, self.symm.pad_ptrs, self.symm.mc_ptr, off, rows, H, residual, w, eps, w_offset,
                                        self.rank, y_sym)
        return (res[0], res[1]) if w is not None else res[0]

    def align_parity(self):
        """Symmetric output regions alternate per call and a region may only be rewritten two calls later (the barrier of the
        call in between proves every peer finished reading it).  A CUDA graph replays a fixed region sequence, so a captured
        step must use an even number of regions: pad with one barrier if it did not."""
        if self._ar_calls & 1:
            self.symm.barrier()
            self._ar_calls += 1

    # ------------------------------------------------------------------ autograd-aware entry points used by parallel/tp.py
    def gemm_rs(self, x, w):
        return _GemmRS.apply(x, w, self)

    def ag_gemm(self, x_local, w, bias=None):
        y = _AGGemm.apply(x_local, w, self)
        return y if bias is None else y + bias

    def gemm_ar(self, x, w):
        return _GemmAR.apply(x, w, self)


def _gather_tokens(t: torch.Tensor, f: FusedTP) -> torch.Tensor:
    out = torch.empty((t.shape[0] * f.world, *t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=f.group)
    return out


class _GemmRS(torch.autograd.Function):
    """y_local[T/t, N] = reduce_scatter(x[T, K/t] @ w[N, K/t]^T)"""

    @staticmethod
    def forward(ctx, x, w, f: FusedTP):
        ctx.save_for_backward(x, w)
        ctx.f = f
        return f._gemm_rs_raw(x.contiguous() if x.stride(1) != 1 else x, w, False)

    @staticmethod
    def backward(ctx, dy_local):
        x, w = ctx.saved_tensors
        f: FusedTP = ctx.f
        dy_local = dy_local.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            # dX[T, K/t] = all_gather(dY)[T, N] @ w[N, K/t]  (B given as [K_red=N, N_out=K/t] -> MN-major)
            if dy_local.shape[0] % 128 == 0 and dy_local.shape[0] * f.world <= f.max_tokens and dy_local.shape[1] <= f.max_features:
                dx = f._ag_gemm_raw(dy_local, w, True)
            else:
                dx = _gather_tokens(dy_local, f) @ w
        if ctx.needs_input_grad[1]:
            from realhf_b200.ops import gemm as G
            dy = _gather_tokens(dy_local, f)
            dw = G.gemm(dy, x, a_mn=True, b_mn=True)
        return dx, dw, None


class _AGGemm(torch.autograd.Function):
    """y[T, N/t] = all_gather(x_local[T/t, K]) @ w[N/t, K]^T"""

    @staticmethod
    def forward(ctx, x_local, w, f: FusedTP):
        ctx.save_for_backward(x_local, w)
        ctx.f = f
        return f._ag_gemm_raw(x_local, w, False)

    @staticmethod
    def backward(ctx, dy):
        x_local, w = ctx.saved_tensors
        f: FusedTP = ctx.f
        dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            # dX_local[T/t, K] = reduce_scatter(dY[T, N/t] @ w[N/t, K])
            dx = f._gemm_rs_raw(dy, w, True)
        if ctx.needs_input_grad[1]:
            from realhf_b200.ops import gemm as G
            xg = _gather_tokens(x_local, f)
            dw = G.gemm(dy, xg, a_mn=True, b_mn=True)
        return dx, dw, None


class _GemmAR(torch.autograd.Function):
    """y[T, N] = all_reduce(x[T, K/t] @ w[N, K/t]^T) (no sequence parallelism; decode path)."""

    @staticmethod
    def forward(ctx, x, w, f: FusedTP):
        ctx.save_for_backward(x, w)
        from realhf_b200.ops import gemm as G
        x2 = x.reshape(-1, x.shape[-1])
        if G.supported(x2, w) and x2.stride(-1) == 1 and x2.stride(0) % 8 == 0 and x2.data_ptr() % 16 == 0:
            y_sym = f.symm_out(x2.shape[0], w.shape[0], x.dtype)
            if y_sym is not None:  # the GEMM writes its partial sums straight into symmetric memory
                G.gemm(x2, w, out=y_sym)
                return f.all_reduce_symm(y_sym).view(*x.shape[:-1], w.shape[0])
        y = OF.linear(x, w)
        return f.all_reduce(y)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        return OF.linear(dy, w.t().contiguous()) if ctx.needs_input_grad[0] else None, \
            (dy.reshape(-1, dy.shape[-1]).t() @ x.reshape(-1, x.shape[-1])) if ctx.needs_input_grad[1] else None, None
