"""W8A8 (e4m3) linear layers for generation.

The decode step of a 7B model at <= 128 sequences per GPU reads every weight once per token: its floor is weight bytes /
HBM bandwidth.  Storing the weights as e4m3 with one fp32 scale per output channel halves that floor; activations are
quantised per token on the fly (`csrc/quant.cu`) and the product runs on `tcgen05.mma.kind::f8f6f4` in the same stream-K
kernel as the bf16 path (`csrc/gemm_tcgen05.cu`, `kFmt == 2`), both scales applied in the epilogue:

    y[m, n] = (sum_k qx[m, k] * qw[n, k]) * sx[m] * sw[n]

The reference generates in the training dtype only (`realhf/impl/model/nn/real_llm_generate.py`); this path is opt-in
(`GenerationHyperparameters.fp8_weights` / `REAL_GEN_FP8=1`) because sampled tokens and their log-probs drift from the
bf16 policy by the quantisation error (tests bound it).  Training and all inference MFCs stay bf16.
"""

from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from realhf_b200.ops import lib
from realhf_b200.ops.gemm import _sms, streamk_workspace

E4M3_MAX = 448.0


def emulate() -> bool:
    """`REAL_FP8_EMULATE=1`: tensors that are not on a CUDA device take the PyTorch implementation of every piece (same
    quantisation rule, fp32 matmul of the dequantised operands).  Lets the CPU suite run the whole W8A8 decode path -- which
    weights are quantised, where biases and norm offsets enter -- for every model family."""
    import os
    return os.environ.get("REAL_FP8_EMULATE", "0") == "1"


def quantize_rows(x: torch.Tensor, q_out: Optional[torch.Tensor] = None, scale_out: Optional[torch.Tensor] = None
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """x [M, K] -> (q uint8 [M, K] holding e4m3 bytes, scale fp32 [M]) with x ~= q * scale[:, None]."""
    if not x.is_cuda:
        q, s = quantize_rows_ref(x)
        if q_out is not None:
            q_out.copy_(q)
            scale_out.copy_(s)
            return q_out, scale_out
        return q, s
    q, s = lib().quant_rows_e4m3(x, q_out, scale_out)
    return q, s


def quantize_rows_ref(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Plain PyTorch version of `quantize_rows` (tests, CPU)."""
    xf = x.float()
    amax = xf.abs().amax(dim=1)
    # tensor / tensor: a true IEEE division like the kernel's `__fdiv_rn` (tensor / python-scalar multiplies by a rounded reciprocal,
    # which lands 1 ulp away about half of the time and then moves exact rounding ties of x / s)
    s = torch.where(amax > 0, amax / torch.full_like(amax, E4M3_MAX), torch.ones_like(amax))
    q = (xf / s[:, None]).clamp_(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn)
    return q.view(torch.uint8), s


def dequantize(q: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    return q.view(torch.float8_e4m3fn).float() * s[:, None]


def gemm_fp8(qx: torch.Tensor, sx: torch.Tensor, qw: torch.Tensor, sw: torch.Tensor, bias: Optional[torch.Tensor] = None,
             out_dtype=torch.bfloat16, out: Optional[torch.Tensor] = None, bn: int = 0, split: int = 0) -> torch.Tensor:
    if not qx.is_cuda:
        y = dequantize(qx, sx[: qx.shape[0]]) @ dequantize(qw, sw).t()
        if bias is not None:
            y = y + bias.float()
        return y.to(out_dtype)
    ws, flags = streamk_workspace(qx.device)
    return lib().gemm_streamk_fp8(qx, qw, sx, sw, out, bias, ws, flags, out_dtype, bn, split, _sms(qx.device))


class Fp8Linear:
    """One quantised weight matrix [N, K] (+ optional bias).  `__call__(x)` quantises x per row and multiplies."""

    __slots__ = ("qw", "sw", "bias", "N", "K")

    def __init__(self, w: torch.Tensor, bias: Optional[torch.Tensor] = None):
        assert w.dim() == 2
        self.N, self.K = w.shape
        self.qw, self.sw = quantize_weight(w)
        self.bias = bias

    def requantize(self, w: torch.Tensor):
        """Refresh the e4m3 copy in place after the weight changed (same buffers: a kept CUDA graph stays valid)."""
        assert tuple(w.shape) == (self.N, self.K)
        quantize_weight(w, out=(self.qw, self.sw))

    def __call__(self, x: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
        lead = x.shape[:-1]
        x2 = x.reshape(-1, self.K)
        if x2.stride(-1) != 1 or x2.stride(0) % 8 or x2.data_ptr() % 16:
            x2 = x2.contiguous()
        qx, sx = quantize_rows(x2)
        y = gemm_fp8(qx, sx, self.qw, self.sw, self.bias if bias is None else bias, out_dtype=x.dtype)
        return y.view(*lead, self.N)

    def gemm_q(self, qx: torch.Tensor, sx: torch.Tensor, bias: Optional[torch.Tensor] = None, out_dtype=torch.bfloat16) -> torch.Tensor:
        """Input already quantised by a fused producer (`add_rmsnorm_quant`, `gated_act_quant`)."""
        return gemm_fp8(qx, sx, self.qw, self.sw, self.bias if bias is None else bias, out_dtype=out_dtype)

    def nbytes(self) -> int:
        return self.qw.numel() + 4 * self.sw.numel()


def quantize_weight(w: torch.Tensor, chunk_rows: int = 8192, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-output-channel e4m3 copy of a weight [N, K], quantised in row chunks (bounded transient memory)."""
    N, K = w.shape
    if out is not None:
        q, s = out
    else:
        q = torch.empty(N, K, dtype=torch.uint8, device=w.device)
        s = torch.empty(N, dtype=torch.float32, device=w.device)
    for lo in range(0, N, chunk_rows):
        hi = min(N, lo + chunk_rows)
        quantize_rows(w[lo:hi], q[lo:hi], s[lo:hi])
    return q, s


def gated_act_quant(gu: torch.Tensor, kind: str) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
    """[M, 2F] = [gate | up] -> e4m3(act(gate) * up) in one kernel; None when the shape / activation has no fused kernel."""
    from realhf_b200.ops.functional import _ACT_KIND
    k = _ACT_KIND.get(kind)
    F = gu.shape[-1] // 2
    if not gu.is_cuda:
        from realhf_b200.ops.functional import gated_act
        return quantize_rows_ref(gated_act(gu, kind))
    if k is None or gu.dim() != 2 or F % 8 or F > 16384 or gu.stride(-1) != 1 or gu.stride(0) % 8 or gu.dtype not in (torch.bfloat16, torch.float16):
        return None
    q, s = lib().gated_act_quant_e4m3(gu, k)
    return q, s


def add_rmsnorm_quant(d: Optional[torch.Tensor], x: torch.Tensor, w: torch.Tensor, eps: float, w_offset: float = 0.0):
    """Residual add + RMSNorm + e4m3 quantisation of the normalised rows in one kernel.
    Returns (q, scale, new residual stream) or None when the shape has no fused kernel (H % 8, H > 8192, dtype)."""
    H = x.shape[-1]
    if not x.is_cuda:
        from realhf_b200.ops.functional import add_rmsnorm, rmsnorm
        h, r = (rmsnorm(x, w, eps, w_offset), x) if d is None else add_rmsnorm(d, x, w, eps, w_offset)
        q, s = quantize_rows_ref(h)
        return q, s, r
    if H % 8 or H > 8192 or x.dtype not in (torch.bfloat16, torch.float16) or not x.is_contiguous() or (d is not None and not d.is_contiguous()):
        return None
    if d is None:
        q, s = lib().add_rmsnorm_quant_e4m3(x, None, w, eps, w_offset)
        return q, s, x
    q, s, r = lib().add_rmsnorm_quant_e4m3(d, x, w, eps, w_offset)
    return q, s, r


def supported(w: torch.Tensor, max_rows: int = 128) -> bool:
    """Shapes the fp8 stream-K kernel takes: K a multiple of 16 and <= 16384 (row quantiser), N >= 256."""
    if not w.is_cuda:
        return bool(emulate() and w.dim() == 2)
    return bool(w.dim() == 2 and w.shape[1] % 16 == 0 and w.shape[1] <= 16384 and w.shape[0] >= 256
                and w.dtype in (torch.bfloat16, torch.float16))


def quantize_named(weights: Dict[str, torch.Tensor]) -> Dict[str, "Fp8Linear"]:
    return {k: Fp8Linear(w) for k, w in weights.items() if supported(w)}
