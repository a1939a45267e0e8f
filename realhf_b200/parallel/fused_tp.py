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

from typing import Optional

import torch
import torch.distributed as dist

from realhf_b200.ops import functional as OF
from realhf_b200.ops import lib
from realhf_b200.parallel.symm_mem import SymmetricBuffer


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
        self.symm = SymmetricBuffer(self.ar_off + ar_bytes, group=self.group, device=self.device)
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

    # ------------------------------------------------------------------ eligibility
    def _ok(self, x, w) -> bool:
        return (x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1
                and w.stride(1) == 1 and x.shape[1] % 8 == 0 and w.stride(0) % 8 == 0)

    def can_gemm_reduce(self, x, w) -> bool:
        return self._ok(x, w) and x.shape[0] % self.world == 0 and x.shape[0] <= self.max_tokens and w.shape[0] <= self.max_features \
            and w.shape[0] % 8 == 0

    def can_ag_gemm(self, x, w) -> bool:
        return self._ok(x, w) and x.shape[0] % 128 == 0 and x.shape[0] * self.world <= self.max_tokens and x.shape[1] <= self.max_features

    # ------------------------------------------------------------------ raw fused ops (no autograd)
    def _gemm_rs_raw(self, x, w, b_mn: bool):
        par = self._rs_calls & 1
        self._rs_calls += 1
        inbox = [p + self.inbox_off[par] for p in self.symm.data_ptrs]
        return lib().gemm_rs(x, w, b_mn, inbox, self._counter_ptrs[par], self._calls_ptr[par], self.rank, self.sms)

    def _ag_gemm_raw(self, x_local, w, b_mn: bool):
        par = self._ag_calls & 1
        self._ag_calls += 1
        rows, K = x_local.shape
        stage = self.symm.data()[self.stage_off[par]: self.stage_off[par] + rows * K * 2].view(torch.bfloat16).view(rows, K)
        stage.copy_(x_local)
        self.symm.barrier()  # every rank's slice is staged (and visible) before anyone's TMA reads it
        peers = [p + self.stage_off[par] for p in self.symm.data_ptrs]
        return lib().ag_gemm(peers, rows, K, w, b_mn, self.rank, self.sms)

    def all_reduce(self, x: torch.Tensor) -> torch.Tensor:
        n = x.numel() * x.element_size()
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
        """Sum over ranks of a tensor obtained from `symm_out` (one barrier, no staging copy)."""
        out = torch.empty(y_sym.shape, dtype=y_sym.dtype, device=y_sym.device)
        lib().symm_allreduce(y_sym, out, self.symm.data_ptrs, self.symm.pad_ptrs, self.rank, 3)
        return out

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
