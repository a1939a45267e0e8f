"""tcgen05 GEMM front-end: `gemm()` raw call and `linear()` with a custom autograd (fwd / dgrad / wgrad all on
the same kernel, no transposed copies; wgrad can accumulate into an fp32 main-grad buffer)."""

from __future__ import annotations

from typing import Optional

import torch

from realhf_b200.ops import lib

_NUM_SMS = {}


def _sms(dev) -> int:
    i = torch.device(dev).index or 0
    if i not in _NUM_SMS:
        _NUM_SMS[i] = torch.cuda.get_device_properties(i).multi_processor_count
    return _NUM_SMS[i]


def gemm(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
         a_mn: bool = False, b_mn: bool = False, accumulate: bool = False, out_dtype=None, bn: int = 0, mc: int = -1) -> torch.Tensor:
    """D = A x B.  A is [M,K] (or [K,M] if a_mn), B is [N,K] (or [K,N] if b_mn)."""
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
        dy2 = dy.reshape(-1, dy.shape[-1])
        if dy2.stride(-1) != 1 or dy2.stride(0) % 8 != 0 or dy2.data_ptr() % 16 != 0:
            dy2 = dy2.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = gemm(dy2, w, b_mn=True).view(ctx.x_shape)            # [T,N] x [N,K] -> [T,K]
        if ctx.needs_input_grad[1]:
            main_grad = getattr(w, "main_grad", None)
            if main_grad is not None:                                   # accumulate straight into the flat grad bucket
                gemm(dy2, x2, out=main_grad.view(w.shape), a_mn=True, b_mn=True, accumulate=True)
                dw = None
            else:
                dw = gemm(dy2, x2, a_mn=True, b_mn=True)               # [T,N]^T x [T,K] -> [N,K]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy2.sum(0)
        return dx, dw, db


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    if not supported(x, w):
        return torch.nn.functional.linear(x, w, bias)
    return _Linear.apply(x, w, bias)
