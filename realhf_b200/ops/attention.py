"""Attention entry points: packed varlen causal attention (train / inference / prefill) and
single-token decode attention over a dense KV cache.

CUDA tensors: decode runs the split-KV sm_100a kernel in `csrc/attn_decode.cu` (with fused RoPE and
in-place KV append); varlen forward runs `csrc/attn_fwd_tcgen05.cu` and backward `csrc/attn_bwd_tcgen05.cu`
(TMA + tcgen05 + TMEM; head dim 64 / 128, validated on B200 against the fp32 reference and the library kernel:
`profiles/attention_validation.log`).  The flash-attn library kernel is only the fallback for what the own
kernels do not cover (attention dropout, other head dims) and the A/B arm: `REAL_ATTN=flash` /
`REAL_ATTN_BWD=flash` select it explicitly; both produce / consume the same LSE layout.
CPU tensors: plain PyTorch reference (also the numerics oracle for the tests).
"""

from __future__ import annotations

import math
from typing import Optional

import torch

import os

from realhf_b200.ops import lib, use_native


def attn_impl() -> str:
    """`tcgen05` (own forward kernel, default) or `flash` (library kernel, A/B runs only)."""
    return os.environ.get("REAL_ATTN", "tcgen05")


def _own_fwd_ok(hd: int, dropout_p: float) -> bool:
    return attn_impl() == "tcgen05" and hd in (64, 128) and dropout_p == 0.0


def _own_bwd_ok(hd: int, dropout_p: float) -> bool:
    """`REAL_ATTN_BWD=tcgen05`: own backward (`csrc/attn_bwd_tcgen05.cu`), independent of the forward switch (both produce /
    consume the same LSE layout).  Default; `REAL_ATTN_BWD=flash` selects the library kernel."""
    return os.environ.get("REAL_ATTN_BWD", "tcgen05") == "tcgen05" and hd in (64, 128) and dropout_p == 0.0


def varlen_attention_ref(q, k, v, cu_seqlens, scale: float, causal: bool = True, sliding_window: Optional[int] = None):
    """q [T,nq,hd], k/v [T,nkv,hd] packed; fp32 math; returns [T,nq,hd] in q.dtype."""
    T, nq, hd = q.shape
    nkv = k.shape[1]
    rep = nq // nkv
    out = torch.empty_like(q)
    cu = cu_seqlens.tolist()
    for i in range(len(cu) - 1):
        s, e = cu[i], cu[i + 1]
        if e == s:
            continue
        qi = q[s:e].float().transpose(0, 1)                          # [nq, L, hd]
        ki = k[s:e].float().transpose(0, 1).repeat_interleave(rep, 0)
        vi = v[s:e].float().transpose(0, 1).repeat_interleave(rep, 0)
        att = (qi @ ki.transpose(1, 2)) * scale
        L = e - s
        if causal:
            idx = torch.arange(L, device=q.device)
            mask = idx[None, :] > idx[:, None]
            if sliding_window:
                mask |= idx[None, :] <= idx[:, None] - sliding_window
            att = att.masked_fill(mask, float("-inf"))
        out[s:e] = (torch.softmax(att, dim=-1) @ vi).transpose(0, 1).to(q.dtype)
    return out


def varlen_attention(q, k, v, cu_seqlens, max_seqlen: int, scale: Optional[float] = None, causal: bool = True,
                     dropout_p: float = 0.0):
    """Packed variable-length causal attention with GQA.  q [T,nq,hd]; k,v [T,nkv,hd]."""
    scale = scale if scale is not None else 1.0 / math.sqrt(q.shape[-1])
    if use_native(q) and q.dtype in (torch.bfloat16, torch.float16):
        from flash_attn import flash_attn_varlen_func
        cu = cu_seqlens.int()
        if _own_fwd_ok(q.shape[-1], dropout_p) and not torch.is_grad_enabled() and all(
                t.stride(-1) == 1 and t.stride(1) == t.shape[-1] for t in (q, k, v)):
            return lib().attn_fwd(q, k, v, cu, max_seqlen, scale, causal)[0]
        return flash_attn_varlen_func(q, k, v, cu, cu, max_seqlen, max_seqlen, dropout_p=dropout_p,
                                      softmax_scale=scale, causal=causal)
    return varlen_attention_ref(q, k, v, cu_seqlens, scale, causal)


class _PackedQKVAttention(torch.autograd.Function):
    """Varlen attention on the fused projection output qkv [T, (nq + 2 nkv) * hd] (q heads | k heads | v heads).

    Slicing q / k / v out of `qkv` in autograd costs three zero-filled [T, 3H] buffers, three slice copies and two adds per
    layer in the backward pass (0.5 GB each at 20k tokens: ~3% of a 7B PPO step).  Here the backward kernel writes dq / dk / dv
    straight into views of ONE gradient buffer that is returned as d(qkv)."""

    @staticmethod
    def forward(ctx, qkv, cu, max_seqlen, nq, nkv, hd, scale, causal, dropout_p):
        from flash_attn.flash_attn_interface import _wrapped_flash_attn_varlen_forward
        T = qkv.shape[0]
        q = qkv[:, : nq * hd].view(T, nq, hd)
        k = qkv[:, nq * hd:(nq + nkv) * hd].view(T, nkv, hd)
        v = qkv[:, (nq + nkv) * hd:].view(T, nkv, hd)
        if _own_fwd_ok(hd, dropout_p):
            out, lse = lib().attn_fwd(q, k, v, cu, max_seqlen, scale, causal)
            ctx.save_for_backward(qkv, out, lse, cu, torch.empty(2, dtype=torch.int64, device=qkv.device))
            ctx.meta = (max_seqlen, nq, nkv, hd, scale, causal, dropout_p)
            return out
        out, lse, _, rng = _wrapped_flash_attn_varlen_forward(q, k, v, cu, cu, max_seqlen, max_seqlen, dropout_p, scale, causal=causal,
                                                              window_size_left=-1, window_size_right=-1, softcap=0.0, alibi_slopes=None,
                                                              return_softmax=False, block_table=None)
        ctx.save_for_backward(qkv, out, lse, cu, rng)
        ctx.meta = (max_seqlen, nq, nkv, hd, scale, causal, dropout_p)
        return out

    @staticmethod
    def backward(ctx, dout):
        from flash_attn.flash_attn_interface import _wrapped_flash_attn_varlen_backward
        qkv, out, lse, cu, rng = ctx.saved_tensors
        max_seqlen, nq, nkv, hd, scale, causal, dropout_p = ctx.meta
        T = qkv.shape[0]
        q = qkv[:, : nq * hd].view(T, nq, hd)
        k = qkv[:, nq * hd:(nq + nkv) * hd].view(T, nkv, hd)
        v = qkv[:, (nq + nkv) * hd:].view(T, nkv, hd)
        dqkv = torch.empty_like(qkv)
        dq = dqkv[:, : nq * hd].view(T, nq, hd)
        dk = dqkv[:, nq * hd:(nq + nkv) * hd].view(T, nkv, hd)
        dv = dqkv[:, (nq + nkv) * hd:].view(T, nkv, hd)
        if _own_bwd_ok(hd, dropout_p):
            lib().attn_bwd(dout.contiguous().view(T, nq, hd), q, k, v, out.view(T, nq, hd), lse, dq, dk, dv, cu, max_seqlen, scale, causal)
            return dqkv, None, None, None, None, None, None, None, None
        _wrapped_flash_attn_varlen_backward(dout.contiguous(), q, k, v, out, lse, dq, dk, dv, cu, cu, max_seqlen, max_seqlen, dropout_p,
                                            scale, causal, -1, -1, 0.0, None, False, rng_state=rng)
        return dqkv, None, None, None, None, None, None, None, None


def varlen_attention_qkv(qkv, cu_seqlens, max_seqlen: int, nq: int, nkv: int, hd: int, scale: Optional[float] = None,
                         causal: bool = True, dropout_p: float = 0.0):
    """Packed varlen attention taking the fused [T, (nq + 2 nkv) * hd] projection; returns [T, nq, hd]."""
    scale = scale if scale is not None else 1.0 / math.sqrt(hd)
    T = qkv.shape[0]
    if use_native(qkv) and qkv.dtype in (torch.bfloat16, torch.float16) and qkv.stride(-1) == 1 and hd % 8 == 0:
        return _PackedQKVAttention.apply(qkv, cu_seqlens.int(), max_seqlen, nq, nkv, hd, scale, causal, dropout_p)
    q = qkv[:, : nq * hd].view(T, nq, hd)
    k = qkv[:, nq * hd:(nq + nkv) * hd].view(T, nkv, hd)
    v = qkv[:, (nq + nkv) * hd:].view(T, nkv, hd)
    return varlen_attention(q, k, v, cu_seqlens, max_seqlen, scale, causal, dropout_p)


def decode_attention_ref(q, k_cache, v_cache, cache_lens, scale: float):
    """q [B,nq,hd]; caches [B,S,nkv,hd]; cache_lens [B] = number of valid positions (incl. the new token)."""
    B, nq, hd = q.shape
    nkv = k_cache.shape[2]
    rep = nq // nkv
    S = k_cache.shape[1]
    kk = k_cache.float().permute(0, 2, 1, 3).repeat_interleave(rep, 1)   # [B,nq,S,hd]
    vv = v_cache.float().permute(0, 2, 1, 3).repeat_interleave(rep, 1)
    att = torch.einsum("bhd,bhsd->bhs", q.float(), kk) * scale
    mask = torch.arange(S, device=q.device)[None, :] >= cache_lens[:, None].to(q.device)
    att = att.masked_fill(mask[:, None, :], float("-inf"))
    return torch.einsum("bhs,bhsd->bhd", torch.softmax(att, -1), vv).to(q.dtype)


def decode_attention(qkv, k_cache, v_cache, cache_lens, n_q: int, n_kv: int, hd: int, scale: Optional[float] = None,
                     cos=None, sin=None, rot_dim: Optional[int] = None, interleaved: bool = False):
    """One decode step for B sequences.

    qkv: [B, (n_q+2*n_kv)*hd] fused projection of the new token.  `cache_lens[b]` is the number of tokens
    already in the cache (= position of the new token).  The op applies RoPE to q and k (if cos/sin given),
    writes the new k,v at position cache_lens[b] of the caches *in place*, and returns attention over
    positions [0, cache_lens[b]] as [B, n_q*hd].  Does not advance cache_lens.
    """
    scale = scale if scale is not None else 1.0 / math.sqrt(hd)
    B = qkv.shape[0]
    if use_native(qkv) and qkv.dtype in (torch.bfloat16, torch.float16) and hd in (64, 128) \
            and k_cache.dtype == qkv.dtype:
        return lib().decode_attention(qkv, k_cache, v_cache, cache_lens.int(), n_q, n_kv, hd, scale, cos, sin,
                                      rot_dim or hd, interleaved)
    # reference path
    from realhf_b200.ops.functional import rope_ref
    x = qkv
    if cos is not None:
        x = rope_ref(qkv, cos, sin, cache_lens, n_q + n_kv, hd, rot_dim or hd, interleaved)
    q = x[:, : n_q * hd].reshape(B, n_q, hd)
    k = x[:, n_q * hd:(n_q + n_kv) * hd].reshape(B, n_kv, hd)
    v = x[:, (n_q + n_kv) * hd:].reshape(B, n_kv, hd)
    idx = torch.arange(B, device=qkv.device)
    pos = cache_lens.long()
    k_cache[idx, pos] = k.to(k_cache.dtype)
    v_cache[idx, pos] = v.to(v_cache.dtype)
    return decode_attention_ref(q, k_cache, v_cache, cache_lens + 1, scale).reshape(B, n_q * hd)
