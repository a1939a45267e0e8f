"""tcgen05 GEMM front-end: `gemm()` raw call and `linear()` with a custom autograd (fwd / dgrad / wgrad all on
the same kernel, no transposed copies; wgrad can accumulate into an fp32 main-grad buffer)."""

from __future__ import annotations

import os
from typing import Optional

import torch

from realhf_b200.ops import lib

_NUM_SMS = {}


def _sms(dev) -> int:
    i = torch.device(dev).index or 0
    if i not in _NUM_SMS:
        _NUM_SMS[i] = torch.cuda.get_device_properties(i).multi_processor_count
    return _NUM_SMS[i]


_SK_WS = {}
_SK_ENABLED = os.environ.get("REAL_GEMM_STREAMK", "1") != "0"


def streamk_workspace(dev, create: bool = True):
    """Per-device persistent stream-K workspace (fp32 partial slots + self-resetting flags).  Calls that share it must
    be stream-ordered; never allocated while a CUDA graph is being captured (returns None then)."""
    i = torch.device(dev).index
    i = torch.cuda.current_device() if i is None else i
    if i not in _SK_WS:
        if not create or torch.cuda.is_current_stream_capturing():
            return None
        n = _sms(i)
        _SK_WS[i] = (torch.empty(2 * n * 128 * 256, dtype=torch.float32, device=f"cuda:{i}"),
                     torch.zeros(8192, dtype=torch.int32, device=f"cuda:{i}"))
    return _SK_WS[i]


def gemm_streamk(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None, out_dtype=None, bn: int = 0,
                 out: Optional[torch.Tensor] = None,
                 split: int = 0, dbg: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Decode-shaped y = a @ b^T (M <= 128): stream-K over every SM (see gemm_tcgen05.cu)."""
    ws, flags = streamk_workspace(a.device)
    return lib().gemm_streamk(a, b, out, bias, ws, flags, out_dtype, bn, split, _sms(a.device), dbg)


def gemm(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
         a_mn: bool = False, b_mn: bool = False, accumulate: bool = False, out_dtype=None, bn: int = 0, mc: int = -1) -> torch.Tensor:
    """D = A x B.  A is [M,K] (or [K,M] if a_mn), B is [N,K] (or [K,N] if b_mn)."""
    if (_SK_ENABLED and not (a_mn or b_mn or accumulate) and bn == 0 and mc < 0 and a.shape[0] <= 128
            and 256 <= b.shape[0] <= 256 * _sms(a.device) and a.shape[1] >= 256):
        wsf = streamk_workspace(a.device)
        if wsf is not None:
            return lib().gemm_streamk(a, b, out, bias, wsf[0], wsf[1], out_dtype, 0, 0, _sms(a.device), None)
    return lib().gemm(a, b, out, bias, a_mn, b_mn, accumulate, out_dtype, bn, _sms(a.device), mc)


def supported(x: torch.Tensor, w: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and w.dtype == x.dtype and x.shape[-1] % 8 == 0
            and w.shape[0] % 8 == 0 and w.stride(-1) == 1 and w.stride(0) % 8 == 0)


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias):
        x2 = x.reshape(-1, x.shape[-1])
        if x2.stride(-1) != 1 or x2.stride(0) % 8 != 0 or x2.data_ptr() % 16 != 0:
            x2 = x2.contiguous()
        ctx.save_for_backward(x2, w)
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        y = gemm(x2, w, bias=bias)
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        dx, dw = _linear_backward(x2, w, dy, ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.x_shape)
        db = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.reshape(-1, dy.shape[-1]).sum(0)
        return dx, dw, db


def _linear_backward(x2, w, dy, need_dx: bool, need_dw: bool, x_shape):
    """dgrad and wgrad of y = x2 @ w^T on the tcgen05 kernel (no transposed copies); the weight gradient accumulates straight
    into the flat gradient buffer when the parameter lives in one (returns dw = None then)."""
    dy2 = dy.reshape(-1, dy.shape[-1])
    if dy2.stride(-1) != 1 or dy2.stride(0) % 8 != 0 or dy2.data_ptr() % 16 != 0:
        dy2 = dy2.contiguous()
    dx = dw = None
    if need_dx:
        dx = gemm(dy2, w, b_mn=True).view(x_shape)                     # [T,N] x [N,K] -> [T,K]
    if need_dw:
        main_grad = getattr(w, "main_grad", None)
        if main_grad is None and getattr(w, "_grad_in_flat_buffer", False) and w.grad is not None and w.grad.dtype == dy2.dtype:
            main_grad = w.grad                                          # same-dtype flat gradient buffer: no temp dW + add
        if main_grad is not None:                                       # accumulate straight into the flat grad bucket
            gemm(dy2, x2, out=main_grad.view(w.shape), a_mn=True, b_mn=True, accumulate=True)
        else:
            dw = gemm(dy2, x2, a_mn=True, b_mn=True)                   # [T,N]^T x [T,K] -> [N,K]
    return dx, dw


class _GatedLinear(torch.autograd.Function):
    """act(x @ Wg^T) * (x @ Wu^T) for the fused [Wg; Wu] weight with the gated activation in the GEMM epilogue
    (`csrc/gemm_2cta.cu`, kGlu): the [T, 2F] projection is not re-read by a separate activation kernel; it is written once
    only when a backward pass will need it."""

    @staticmethod
    def forward(ctx, x, w, kind: int):
        x2 = x.reshape(-1, x.shape[-1])
        if x2.stride(-1) != 1 or x2.stride(0) % 8 != 0 or x2.data_ptr() % 16 != 0:
            x2 = x2.contiguous()
        need_raw = any(ctx.needs_input_grad[:2])
        res = lib().gemm_glu(x2, w, kind, need_raw, _sms(x.device))
        if need_raw:
            ctx.save_for_backward(x2, w, res[1])
        ctx.kind, ctx.x_shape = kind, x.shape
        return res[0].view(*x.shape[:-1], w.shape[0] // 2)

    @staticmethod
    def backward(ctx, dact):
        x2, w, gu = ctx.saved_tensors
        dgu = lib().gated_act_bwd(gu, dact.reshape(-1, dact.shape[-1]).contiguous(), ctx.kind)
        dx, dw = _linear_backward(x2, w, dgu, ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.x_shape)
        return dx, dw, None


def gated_linear_supported(x: torch.Tensor, w: torch.Tensor) -> bool:
    n_rows = x.numel() // max(1, x.shape[-1])
    return (supported(x, w) and n_rows > 128 and w.shape[0] % 16 == 0 and (w.shape[0] // 2) % 8 == 0 and w.is_contiguous()
            and os.environ.get("REAL_GEMM_GLU", "1") == "1")


def gated_linear(x: torch.Tensor, w: torch.Tensor, kind: int) -> torch.Tensor:
    return _GatedLinear.apply(x, w, kind)


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    if not supported(x, w):
        return torch.nn.functional.linear(x, w, bias)
    return _Linear.apply(x, w, bias)


def pad_groups(offsets: torch.Tensor, n_rows: int, block: int = 64):
    """Row map that pads every group's block of a group-sorted [n_rows, *] tensor to a multiple of `block` rows, computed on
    the device (no host read of the group sizes).  Returns (dest [n_rows] int64: padded position of every source row,
    offsets_pad [G + 1] int32: padded block starts, n_pad: static upper bound of padded rows, a multiple of `block`).
    Rows outside [offsets[0], offsets[G]) belong to no group (an expert-parallel rank that only owns a sub-range of the sorted
    rows; the unused tail of a receive buffer): their `dest` is n_pad, one row past the padded tensor -- callers allocate
    n_pad + 1 rows and hand the first n_pad to the kernel, so stale data never lands in a group's zero padding."""
    G = offsets.numel() - 1
    off = offsets.long()
    counts = off[1:] - off[:-1]
    padded = (counts + block - 1) // block * block
    off_pad = torch.zeros(G + 1, dtype=torch.long, device=offsets.device)
    off_pad[1:] = padded.cumsum(0)
    rows = torch.arange(n_rows, device=offsets.device)
    g = torch.bucketize(rows, off[1:], right=True).clamp_(max=G - 1)
    n_pad = (n_rows + G * (block - 1) + block - 1) // block * block
    dest = rows + (off_pad[:-1] - off[:-1])[g]
    dest = torch.where((rows >= off[0]) & (rows < off[G]), dest, torch.full_like(dest, n_pad))
    return dest, off_pad.int(), n_pad


def grouped_wgrad_ref(dy, x, offsets, G):
    """fp32 reference: dW[g] = dy[rows of g]^T @ x[rows of g]."""
    off = offsets.tolist()
    return torch.stack([dy[off[g]:off[g + 1]].float().t() @ x[off[g]:off[g + 1]].float() for g in range(G)])


def grouped_wgrad(dy: torch.Tensor, x: torch.Tensor, offsets: torch.Tensor, G: int, out: Optional[torch.Tensor] = None,
                  accumulate: bool = False) -> torch.Tensor:
    """dW [G, M, N] for group-sorted dy [T, M], x [T, N] in ONE launch (`csrc/gemm_grouped_wgrad.cu`): both tensors are
    scattered into zero-padded buffers whose group blocks start at multiples of 64 rows, so that no K block of the kernel
    straddles two experts.  No host synchronisation."""
    T = dy.shape[0]
    dest, off_pad, n_pad = pad_groups(offsets, T)
    # one spare row at the end swallows the rows that belong to no group (see pad_groups)
    dy_p = torch.zeros(n_pad + 1, dy.shape[1], dtype=dy.dtype, device=dy.device).index_copy_(0, dest, dy)[:n_pad]
    x_p = torch.zeros(n_pad + 1, x.shape[1], dtype=x.dtype, device=x.device).index_copy_(0, dest, x)[:n_pad]
    if out is None:
        out = torch.empty(G, dy.shape[1], x.shape[1], dtype=dy.dtype, device=dy.device)
        accumulate = False
    lib().gemm_grouped_wgrad(dy_p, x_p, out, off_pad, accumulate, _sms(dy.device))
    return out


def _grouped_wgrad_enabled() -> bool:
    import os
    return os.environ.get("REAL_MOE_GROUPED_WGRAD", "1") == "1"  # single-launch kernel (validated on B200); 0 = per-expert loop


class _GroupedLinear(torch.autograd.Function):
    """y[rows of group g] = x[rows of group g] @ w[g]^T for rows sorted by group (MoE experts), one kernel launch."""

    @staticmethod
    def forward(ctx, x, w, offsets):
        ctx.save_for_backward(x, w, offsets)
        return lib().gemm_grouped(x, w, offsets, False, _sms(x.device))

    @staticmethod
    def backward(ctx, dy):
        x, w, offsets = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = lib().gemm_grouped(dy, w, offsets, True, _sms(x.device))        # [rows_g, N] x [N, K]
        if ctx.needs_input_grad[1] and _grouped_wgrad_enabled():
            dw = grouped_wgrad(dy, x, offsets, w.shape[0])
        elif ctx.needs_input_grad[1]:
            # per-group wgrad: the reduction runs over a data-dependent row range, so it stays one GEMM per group
            dw = torch.zeros_like(w)
            off = offsets.tolist()
            for g in range(w.shape[0]):
                a, b = off[g], off[g + 1]
                if b - a >= 8 and (b - a) % 8 == 0:
                    gemm(dy[a:b], x[a:b], out=dw[g], a_mn=True, b_mn=True)
                elif b > a:
                    dw[g] = (dy[a:b].float().t() @ x[a:b].float()).to(w.dtype)
        return dx, dw, None


def grouped_linear(x: torch.Tensor, w: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
    """x [M, K] sorted by group, w [G, N, K], offsets int32 [G + 1] (device) -> [M, N]."""
    return _GroupedLinear.apply(x, w, offsets)


def grouped_supported(x: torch.Tensor, w: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and w.dim() == 3 and w.is_contiguous()
            and x.dim() == 2 and x.stride(1) == 1 and x.stride(0) % 8 == 0 and w.shape[1] % 8 == 0 and w.shape[2] % 8 == 0
            and w.shape[0] <= 256)
