"""Python face of the kernels: every op = (PyTorch reference for CPU tensors, sm_100a kernel for CUDA tensors).

The `*_ref` functions are the numerics oracle used by `tests/` (fp32 PyTorch of the same op).
"""

from __future__ import annotations

import os

from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F

from realhf_b200.ops import lib, use_native

# ------------------------------------------------------------------------------------------------ GAE


def gae_1d_misalign_ref(rewards, values, cu_seqlens, bootstrap, gamma, lam):
    """Serial reference.  rewards [sum L], values [sum (L+1)], cu_seqlens over rewards, bootstrap [bs]."""
    bs = cu_seqlens.numel() - 1
    adv = torch.zeros_like(rewards)
    cu = cu_seqlens.tolist()
    for i in range(bs):
        r0, r1 = cu[i], cu[i + 1]
        v = values[r0 + i: r1 + i + 1].clone()
        if not bool(bootstrap[i]):
            v[-1] = 0
        last = 0.0
        for t in reversed(range(r1 - r0)):
            delta = rewards[r0 + t] + gamma * v[t + 1] - v[t]
            last = delta + gamma * lam * last
            adv[r0 + t] = last
    idx = torch.arange(rewards.numel(), device=rewards.device)
    seq_of = torch.bucketize(idx, cu_seqlens[1:].to(idx.dtype), right=True)
    ret = adv + values[idx + seq_of]
    return adv, ret


def gae_1d_misalign(rewards, values, cu_seqlens, bootstrap, gamma: float, lam: float):
    """GAE over packed sequences (reference: cugae `gae_1d_nolp_misalign`)."""
    if use_native(rewards):
        a, r = lib().gae_1d_misalign(rewards.float().contiguous(), values.float().contiguous(),
                                     cu_seqlens.int().contiguous(), bootstrap.bool().contiguous(), gamma, lam)
        return a, r
    return gae_1d_misalign_ref(rewards.float(), values.float(), cu_seqlens, bootstrap, gamma, lam)


def ppo_rewards_gae_ref(logp, ref_logp, scores, values, cu_seqlens, no_eos, gamma, lam, kl_ctl, clip_reward):
    kl = -kl_ctl * (logp - ref_logp)
    tot = kl.clone()
    sc = scores.clamp(-clip_reward, clip_reward)
    ends = (cu_seqlens[1:] - 1).long()
    tot[ends] += torch.where(no_eos.bool(), torch.zeros_like(sc), sc)
    adv, ret = gae_1d_misalign_ref(tot, values, cu_seqlens, no_eos, gamma, lam)
    return adv, ret, kl, tot


def ppo_rewards_gae(logp, ref_logp, scores, values, cu_seqlens, no_eos, gamma, lam, kl_ctl, clip_reward):
    """KL-shaped reward + terminal score + GAE in one launch.  Returns (adv, returns, kl_rewards, rewards)."""
    if use_native(logp):
        return tuple(lib().ppo_rewards_gae(logp.float().contiguous(), ref_logp.float().contiguous(),
                                           scores.float().contiguous(), values.float().contiguous(),
                                           cu_seqlens.int().contiguous(), no_eos.bool().contiguous(),
                                           gamma, lam, kl_ctl, clip_reward))
    return ppo_rewards_gae_ref(logp.float(), ref_logp.float(), scores.float(), values.float(), cu_seqlens, no_eos,
                               gamma, lam, kl_ctl, clip_reward)


def gae_2d_ref(rewards, values, dones, truncs, gamma, lam, mode: str):
    T = rewards.shape[1]
    nd, nt = 1 - dones.float(), 1 - truncs.float()
    delta = rewards + gamma * values[:, 1:] * nd[:, 1:] - values[:, :-1]
    m = gamma * lam * nd[:, 1:] * nt[:, 1:]
    adv = torch.zeros_like(rewards)
    gae = torch.zeros_like(rewards[:, 0])
    for t in reversed(range(T)):
        if mode == "olp":
            gae = delta[:, t] * nt[:, t + 1] + m[:, t] * gae
        else:
            gae = delta[:, t] + m[:, t] * gae
        adv[:, t] = gae
    return adv, adv + values[:, :-1]


def gae_2d(rewards, values, dones, truncs, gamma, lam, mode: str = "olp"):
    """Padded [bs,T] GAE with done/truncate flags (reference: `gae_2d_olp` / `gae_2d_nolp`)."""
    if use_native(rewards):
        return tuple(lib().gae_2d(rewards.float().contiguous(), values.float().contiguous(), dones.bool().contiguous(),
                                  truncs.bool().contiguous(), gamma, lam, 0 if mode == "olp" else 1))
    return gae_2d_ref(rewards.float(), values.float(), dones, truncs, gamma, lam, mode)


# ------------------------------------------------------------------------------------------------ norms


def rmsnorm_ref(x, w, eps, w_offset=0.0, residual=None):
    if residual is not None:
        x = (x.float() + residual.float()).to(x.dtype)
    xf = x.float()
    y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps) * (w.float() + w_offset)
    y = y.to(x.dtype)
    return (y, x) if residual is not None else y


class _RMSNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, eps, w_offset):
        y, rstd = lib().rmsnorm_fwd(x.contiguous(), None, w, eps, w_offset)
        ctx.save_for_backward(x, w, rstd)
        ctx.w_offset = w_offset
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, rstd = ctx.saved_tensors
        dx, dw = lib().rmsnorm_bwd(x.contiguous(), w, dy.contiguous(), rstd, ctx.w_offset, None)
        return dx, dw, None, None


class _AddRMSNorm(torch.autograd.Function):
    """(h, x_new) = (rmsnorm(x + d), x + d) in one kernel, with a one-kernel backward: the gradient of the normalised branch
    and the gradient arriving on the residual stream are summed inside the RMSNorm backward kernel.  Replaces, per residual
    connection of a training step, one eager add in the forward pass and the autograd accumulation add in the backward pass."""

    @staticmethod
    def forward(ctx, d, x, w, eps, w_offset):
        y, rstd, res = lib().rmsnorm_fwd(d.contiguous(), x.contiguous(), w, eps, w_offset)
        ctx.save_for_backward(res, w, rstd)
        ctx.w_offset = w_offset
        ctx.mark_non_differentiable(rstd)
        return y, res

    @staticmethod
    def backward(ctx, dy, dres):
        res, w, rstd = ctx.saved_tensors
        if dy is None:  # the normalised output was not used
            return dres, dres, None, None, None
        dsum, dw = lib().rmsnorm_bwd(res, w, dy.contiguous(), rstd, ctx.w_offset, dres.contiguous() if dres is not None else None)
        return dsum, dsum, dw, None, None


def rmsnorm(x, w, eps: float, w_offset: float = 0.0):
    """y = x * rsqrt(mean(x^2)+eps) * (w + w_offset); w_offset=1 is the Gemma flavour."""
    if use_native(x) and x.dtype in (torch.bfloat16, torch.float16, torch.float32) and x.shape[-1] % 8 == 0 \
            and x.shape[-1] <= 8192:
        return _RMSNorm.apply(x, w, eps, w_offset)
    return rmsnorm_ref(x, w, eps, w_offset)


def add_rmsnorm(x, residual, w, eps: float, w_offset: float = 0.0):
    """(residual + x) -> new residual, and its RMSNorm: one kernel forward, one kernel backward."""
    if use_native(x) and x.shape[-1] % 8 == 0 and x.shape[-1] <= 8192 and x.dtype in (torch.bfloat16, torch.float16, torch.float32) \
            and x.dtype == residual.dtype == w.dtype:
        if not torch.is_grad_enabled():
            y, _, res = lib().rmsnorm_fwd(x.contiguous(), residual.contiguous(), w, eps, w_offset)
            return y, res
        return _AddRMSNorm.apply(x, residual, w, eps, w_offset)
    res = x + residual
    return rmsnorm(res, w, eps, w_offset), res


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        y, mean, rstd = lib().layernorm_fwd(x.contiguous(), w, b, eps)
        ctx.save_for_backward(x, w, mean, rstd)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, mean, rstd = ctx.saved_tensors
        dx, dw, db = lib().layernorm_bwd(x.contiguous(), w, dy.contiguous(), mean, rstd)
        return dx, dw, (db if ctx.has_bias else None), None


def layer_norm(x, w, b, eps: float):
    """LayerNorm over the last dimension.  CUDA tensors run `csrc/layernorm.cu` (validated on B200 against the fp32 reference);
    `REAL_LAYERNORM=torch` selects `F.layer_norm` for A/B runs."""
    import os
    if os.environ.get("REAL_LAYERNORM", "native") == "native" and use_native(x) and x.dtype == w.dtype \
            and x.dtype in (torch.bfloat16, torch.float16, torch.float32) and x.shape[-1] % 8 == 0 and x.shape[-1] <= 8192:
        return _LayerNorm.apply(x, w, b, eps)
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), w, b, eps)


# ------------------------------------------------------------------------------------------------ rope


def rope_tables(max_pos: int, rot_dim: int, base: float, device, scaling: Optional[float] = None,
                scaling_type: Optional[str] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 cos/sin tables [max_pos, rot_dim/2] (linear and dynamic-NTK scaling, as HF)."""
    if scaling_type == "dynamic" and scaling is not None:
        base = base * scaling ** (rot_dim / (rot_dim - 2))
    inv_freq = 1.0 / (base ** (torch.arange(0, rot_dim, 2, dtype=torch.float32, device=device) / rot_dim))
    t = torch.arange(max_pos, dtype=torch.float32, device=device)
    if scaling_type == "linear" and scaling is not None:
        t = t / scaling
    freqs = torch.outer(t, inv_freq)
    return freqs.cos().contiguous(), freqs.sin().contiguous()


def rope_ref(x, cos, sin, pos, n_heads, hd, rot_dim, interleaved=False, inverse=False):
    """x [T, row]; rotates heads [0,n_heads) in the leading n_heads*hd columns.  Returns a new tensor."""
    T = x.shape[0]
    out = x.clone()
    xh = x[:, : n_heads * hd].reshape(T, n_heads, hd).float()
    c = cos[pos.long()].unsqueeze(1)
    s = sin[pos.long()].unsqueeze(1)
    if inverse:
        s = -s
    rot = xh[..., :rot_dim]
    if interleaved:
        x0, x1 = rot[..., 0::2], rot[..., 1::2]
        r = torch.stack([x0 * c - x1 * s, x1 * c + x0 * s], dim=-1).flatten(-2)
    else:
        x0, x1 = rot[..., : rot_dim // 2], rot[..., rot_dim // 2:]
        r = torch.cat([x0 * c - x1 * s, x1 * c + x0 * s], dim=-1)
    xh = torch.cat([r, xh[..., rot_dim:]], dim=-1)
    out[:, : n_heads * hd] = xh.reshape(T, n_heads * hd).to(x.dtype)
    return out


class _Rope(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cos, sin, pos, n_heads, hd, rot_dim, interleaved):
        ctx.save_for_backward(cos, sin, pos)
        ctx.cfg = (n_heads, hd, rot_dim, interleaved)
        lib().rope_inplace(x, cos, sin, pos, n_heads, hd, rot_dim, interleaved, False)
        ctx.mark_dirty(x)
        return x

    @staticmethod
    def backward(ctx, dx):
        cos, sin, pos = ctx.saved_tensors
        n_heads, hd, rot_dim, interleaved = ctx.cfg
        dx = dx.contiguous().clone() if not dx.is_contiguous() else dx.clone()
        lib().rope_inplace(dx, cos, sin, pos, n_heads, hd, rot_dim, interleaved, True)
        return dx, None, None, None, None, None, None, None


def apply_rope(x, cos, sin, pos, n_heads: int, hd: int, rot_dim: Optional[int] = None, interleaved: bool = False):
    """Rotate the first `n_heads` heads of every row of x [T, row] (fused-QKV layout). pos: int32 [T]."""
    rot_dim = rot_dim or hd
    if use_native(x) and x.dtype in (torch.bfloat16, torch.float16) and rot_dim % 16 == 0 and x.stride(0) % 8 == 0:
        if x.requires_grad and x.is_leaf:
            x = x.clone()
        if not x.requires_grad:
            lib().rope_inplace(x, cos, sin, pos.int(), n_heads, hd, rot_dim, interleaved, False)
            return x
        # autograd path: operate on a fresh buffer (the producer may need its output for backward)
        return _Rope.apply(x.clone(), cos, sin, pos.int(), n_heads, hd, rot_dim, interleaved)
    return rope_ref(x, cos, sin, pos, n_heads, hd, rot_dim, interleaved)


# ------------------------------------------------------------------------------------------------ gated act

_ACT_KIND = {"silu": 0, "swiglu": 0, "gelu_pytorch_tanh": 1, "gelu_new": 1, "gelu_tanh": 1}


def gated_act_ref(gu, kind: str):
    F_ = gu.shape[-1] // 2
    g, u = gu[..., :F_], gu[..., F_:]
    if _ACT_KIND[kind] == 0:
        return F.silu(g) * u
    return F.gelu(g, approximate="tanh") * u


class _GatedAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gu, kind):
        ctx.save_for_backward(gu)
        ctx.kind = kind
        return lib().gated_act_fwd(gu.contiguous(), kind)

    @staticmethod
    def backward(ctx, dout):
        (gu,) = ctx.saved_tensors
        return lib().gated_act_bwd(gu.contiguous(), dout.contiguous(), ctx.kind), None


def gated_act(gu, kind: str = "silu"):
    """act(gate) * up on the fused [.., 2F] gate|up projection."""
    if use_native(gu) and gu.dtype in (torch.bfloat16, torch.float16) and (gu.shape[-1] // 2) % 8 == 0:
        return _GatedAct.apply(gu, _ACT_KIND[kind])
    return gated_act_ref(gu, kind)


def gated_linear(x, w_gate_up, kind: str = "silu"):
    """act(x @ Wg^T) * (x @ Wu^T) with the fused [Wg; Wu] weight.  CUDA tensors with the tcgen05 GEMM enabled run the
    projection and the gated activation as ONE kernel (activation in the epilogue); everything else composes the two ops."""
    if use_native(x) and x.dtype in (torch.bfloat16, torch.float16) and kind in _ACT_KIND and gemm_impl() is not None:
        from realhf_b200.ops import gemm as G
        if gemm_impl() is G.linear and G.gated_linear_supported(x, w_gate_up):
            return G.gated_linear(x, w_gate_up, _ACT_KIND[kind])
    return gated_act(linear(x, w_gate_up), kind)


# ------------------------------------------------------------------------------------------------ logprob


def pack_mask_bits(mask: torch.Tensor) -> torch.Tensor:
    """bool [T, V] (True = filtered out) -> uint8 [T, ceil(V/8)], bit j%8 of byte j//8."""
    T, V = mask.shape
    pad = (-V) % 8
    m = F.pad(mask.to(torch.uint8), (0, pad)).view(T, -1, 8)
    w = (2 ** torch.arange(8, device=mask.device, dtype=torch.int32)).view(1, 1, 8)
    return (m.int() * w).sum(-1).to(torch.uint8)


def unpack_mask_bits(bits: torch.Tensor, V: int) -> torch.Tensor:
    w = (2 ** torch.arange(8, device=bits.device, dtype=torch.int32)).view(1, 1, 8)
    return ((bits.int().unsqueeze(-1) & w) != 0).flatten(1)[:, :V]


def logprob_from_logits_ref(logits, labels, mask_bits=None, inv_temp=1.0):
    x = logits.float() * inv_temp
    if mask_bits is not None:
        x = x.masked_fill(unpack_mask_bits(mask_bits, x.shape[1]), float("-inf"))
    lse = torch.logsumexp(x, dim=-1)
    tgt = x.gather(-1, labels.long().unsqueeze(-1)).squeeze(-1)
    return tgt - lse, lse


def logprob_from_logits(logits, labels, mask_bits=None, inv_temp: float = 1.0):
    """(log p(label), logsumexp) per row; no autograd (see `lm_head_logprobs` for the trainable path)."""
    if use_native(logits):
        lp, lse = lib().logprob_fwd(logits, labels.long().contiguous(), mask_bits, inv_temp, 0, False)
        return lp, lse
    return logprob_from_logits_ref(logits, labels, mask_bits, inv_temp)


class _LMHeadLogProb(torch.autograd.Function):
    """log p(label | hidden) through the LM head, chunked over tokens so [T, V] never materialises.

    backward recomputes each logits chunk, turns it into d logits in place and feeds the two GEMMs.
    """

    @staticmethod
    def forward(ctx, hidden, weight, labels, mask_bits, inv_temp, chunk):
        T = hidden.shape[0]
        logp = torch.empty(T, dtype=torch.float32, device=hidden.device)
        lse = torch.empty(T, dtype=torch.float32, device=hidden.device)
        for s in range(0, T, chunk):
            e = min(T, s + chunk)
            logits = linear(hidden[s:e], weight)
            mb = mask_bits[s:e] if mask_bits is not None else None
            if use_native(logits):
                a, b = lib().logprob_fwd(logits, labels[s:e], mb, inv_temp, 0, False)
            else:
                a, b = logprob_from_logits_ref(logits, labels[s:e], mb, inv_temp)
            logp[s:e], lse[s:e] = a, b
        ctx.save_for_backward(hidden, weight, labels, lse, mask_bits if mask_bits is not None else torch.empty(0))
        ctx.has_mask = mask_bits is not None
        ctx.inv_temp, ctx.chunk = inv_temp, chunk
        return logp

    @staticmethod
    def backward(ctx, dlogp):
        hidden, weight, labels, lse, mask_bits = ctx.saved_tensors
        mask_bits = mask_bits if ctx.has_mask else None
        T = hidden.shape[0]
        dlogp = dlogp.float().contiguous()
        dh = torch.empty_like(hidden) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            mg = getattr(weight, "main_grad", None)
            if mg is not None and mg.dtype == torch.float32 and _native_gemm(hidden, weight) is not None:
                dw = mg  # same contract as ops.gemm._Linear: wgrad accumulates into the fp32 main-grad view
            else:
                dw = torch.zeros(weight.shape, dtype=torch.float32, device=weight.device)
        for s in range(0, T, ctx.chunk):
            e = min(T, s + ctx.chunk)
            logits = linear(hidden[s:e], weight)
            mb = mask_bits[s:e] if mask_bits is not None else None
            if use_native(logits):
                lib().logprob_bwd_(logits, labels[s:e], mb, lse[s:e], dlogp[s:e], ctx.inv_temp, 0)
                dlogits = logits
            else:
                x = logits.float() * ctx.inv_temp
                if mb is not None:
                    x = x.masked_fill(unpack_mask_bits(mb, x.shape[1]), float("-inf"))
                p = torch.exp(x - lse[s:e].unsqueeze(-1))
                onehot = F.one_hot(labels[s:e].long(), x.shape[1]).to(p.dtype)
                dlogits = ((onehot - p) * (dlogp[s:e] * ctx.inv_temp).unsqueeze(-1)).to(logits.dtype)
                if mb is not None:
                    dlogits = dlogits.masked_fill(unpack_mask_bits(mb, x.shape[1]), 0)
            G = _native_gemm(dlogits, weight)
            if dh is not None:
                if G is not None:
                    G.gemm(dlogits, weight, out=dh[s:e], b_mn=True)                    # [t,V] x [V,H]
                else:
                    dh[s:e] = dlogits @ weight
            if dw is not None:
                if G is not None:
                    G.gemm(dlogits, hidden[s:e], out=dw, a_mn=True, b_mn=True, accumulate=True)  # [t,V]^T x [t,H] += into fp32
                else:
                    dw += (dlogits.t() @ hidden[s:e]).float()
        if dw is not None and dw is getattr(weight, "main_grad", None):
            dw = None  # accumulated straight into the flat gradient bucket
        return dh, (dw.to(weight.dtype) if dw is not None else None), None, None, None, None


def _native_gemm(x, w):
    """The tcgen05 GEMM module when it is installed as the projection matmul and the operands qualify, else None."""
    if not x.is_cuda or gemm_impl() is None:
        return None
    from realhf_b200.ops import gemm as G
    ok = (x.dtype in (torch.bfloat16, torch.float16) and w.dtype == x.dtype and x.dim() == 2 and x.stride(-1) == 1 and w.stride(-1) == 1
          and x.stride(0) % 8 == 0 and w.stride(0) % 8 == 0 and x.shape[0] % 8 == 0 and x.shape[1] % 8 == 0 and w.shape[0] % 8 == 0
          and x.data_ptr() % 16 == 0 and w.data_ptr() % 16 == 0)
    return G if ok else None


def lm_head_logprobs(hidden, weight, labels, mask_bits=None, temperature: float = 1.0, chunk: int = 8192):
    """Fused LM-head + log-prob gather: hidden [T,H], weight [V,H], labels [T] -> fp32 [T]."""
    return _LMHeadLogProb.apply(hidden, weight, labels.long().contiguous(), mask_bits, 1.0 / temperature, chunk)


def lm_head_logprobs_ref(hidden, weight, labels, mask_bits=None, temperature: float = 1.0):
    logits = hidden.float() @ weight.float().t()
    return logprob_from_logits_ref(logits, labels, mask_bits, 1.0 / temperature)[0]


# ------------------------------------------------------------------------------------------------ GEMM

_GEMM_IMPL = {"fn": None}  # None: auto (tcgen05 on CUDA unless REAL_GEMM=cublas); False: library matmul; callable: that


def set_gemm_impl(fn):
    """Choose the matmul behind every projection: a callable (`ops.gemm.linear`), None for the default (the tcgen05 GEMM on
    CUDA tensors), or False to force the library matmul (cuBLAS; used by the A/B benchmarks)."""
    _GEMM_IMPL["fn"] = fn


def gemm_impl():
    """The active projection matmul for CUDA tensors, or None when the library path is selected."""
    fn = _GEMM_IMPL["fn"]
    if fn is None:
        if os.environ.get("REAL_GEMM", "tcgen05") == "cublas":
            return None
        from realhf_b200.ops import gemm as G
        fn = _GEMM_IMPL["fn"] = G.linear
    return fn or None


def linear(x, w, bias=None):
    """y = x @ w.T (+ bias), w is [out, in] (torch layout)."""
    if x.is_cuda:
        fn = gemm_impl()
        if fn is not None:
            return fn(x, w, bias)
    return F.linear(x, w, bias)


# ------------------------------------------------------------------------------------------------ optimizer


def adamw_ref(p, g, m, v, master, lr, b1, b2, eps, wd, step, scale=1.0):
    g = g.float() * scale
    w = master if master is not None else p.float()
    m.copy_((b1 * m.float() + (1 - b1) * g).to(m.dtype))
    v.copy_((b2 * v.float() + (1 - b2) * g * g).to(v.dtype))
    mh = m.float() / (1 - b1 ** step)
    vh = v.float() / (1 - b2 ** step)
    w_new = w - lr * (mh / (vh.sqrt() + eps) + wd * w)
    if master is not None:
        master.copy_(w_new)
    p.copy_(w_new.to(p.dtype))


def adamw_step(p, g, m, v, master, lr, b1, b2, eps, wd, step, scale: Optional[torch.Tensor] = None,
               skip: Optional[torch.Tensor] = None, stochastic: bool = False, seed: int = 0):
    """In-place AdamW on flat buffers.  scale/skip are device scalars (no host sync)."""
    if use_native(p):
        lib().adamw_step(p, g, m, v, master, lr, b1, b2, eps, wd, step, scale, skip, stochastic, seed)
        return
    if skip is not None and int(skip.item()) != 0:
        return
    adamw_ref(p, g, m, v, master, lr, b1, b2, eps, wd, step, float(scale.item()) if scale is not None else 1.0)


def sumsq_accum(g: torch.Tensor, out2: torch.Tensor):
    """out2[0] += sum(g^2), out2[1] += #non-finite."""
    if use_native(g):
        lib().sumsq_accum(g, out2)
        return
    gf = g.float()
    out2[0] += torch.nan_to_num(gf, nan=0.0, posinf=0.0, neginf=0.0).pow(2).sum()
    out2[1] += (~torch.isfinite(gf)).sum().float()


# ------------------------------------------------------------------------------------------------ segment copy


def segment_copy_ref(src_flat_bytes, dst_flat_bytes, src_off, dst_off, lens, eta=1.0, dtype=None):
    for so, do, ln in zip(src_off, dst_off, lens):
        if eta == 1.0:
            dst_flat_bytes[do: do + ln] = src_flat_bytes[so: so + ln]
        else:
            s = src_flat_bytes[so: so + ln].view(dtype).float()
            d = dst_flat_bytes[do: do + ln].view(dtype).float()
            dst_flat_bytes[do: do + ln] = (eta * s + (1 - eta) * d).to(dtype).view(torch.uint8)


class SegmentPlan:
    """A reusable list of (src_byte_off, dst_byte_off, n_bytes) segments, uploaded once."""

    def __init__(self, src_off, dst_off, lens, device):
        """Offsets / lengths in bytes: python lists or integer numpy arrays (reallocation plans of 7B models have millions of
        segments; everything here is vectorised)."""
        import numpy as np
        so, do, ln = (np.asarray(x, dtype=np.int64).reshape(-1) for x in (src_off, dst_off, lens))
        self.n = int(ln.shape[0])
        self._host = (so, do, ln)
        cum = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(ln, out=cum[1:])
        self.total = int(cum[-1])
        self.device = torch.device(device)
        if self.device.type == "cuda":
            self.src_off = torch.from_numpy(so.copy()).to(device)
            self.dst_off = torch.from_numpy(do.copy()).to(device)
            self.cum = torch.from_numpy(cum).to(device)

    @property
    def src_off_h(self) -> List[int]:
        return self._host[0].tolist()

    @property
    def dst_off_h(self) -> List[int]:
        return self._host[1].tolist()

    @property
    def lens_h(self) -> List[int]:
        return self._host[2].tolist()

    def run(self, src: torch.Tensor, dst: Optional[torch.Tensor] = None, dst_ptr: int = 0, eta: float = 1.0):
        """Copy every segment from `src` into `dst` (or the raw — possibly peer — address `dst_ptr`)."""
        if self.total == 0:
            return
        if src.is_cuda:
            lib().segment_copy(src, dst if dst is not None else src, dst_ptr, self.src_off, self.dst_off, self.cum,
                               self.total, eta, eta != 1.0)
        else:
            segment_copy_ref(src.view(torch.uint8).view(-1), dst.view(torch.uint8).view(-1), self.src_off_h,
                             self.dst_off_h, self.lens_h, eta, src.dtype)


def slice_intervals(src: torch.Tensor, intervals: List[Tuple[int, int]]) -> torch.Tensor:
    """Pack 1-D element intervals [a,b) of a flat tensor into a new contiguous tensor."""
    es = src.element_size()
    lens = [(b - a) * es for a, b in intervals]
    dst_off, acc = [], 0
    for ln in lens:
        dst_off.append(acc)
        acc += ln
    out = torch.empty(acc // es, dtype=src.dtype, device=src.device)
    SegmentPlan([a * es for a, _ in intervals], dst_off, lens, src.device).run(src, out)
    return out


def set_intervals(src: torch.Tensor, dst: torch.Tensor, intervals: List[Tuple[int, int]]):
    """Scatter contiguous `src` into 1-D element intervals of flat `dst`."""
    es = dst.element_size()
    lens = [(b - a) * es for a, b in intervals]
    src_off, acc = [], 0
    for ln in lens:
        src_off.append(acc)
        acc += ln
    assert acc == src.numel() * es
    SegmentPlan(src_off, [a * es for a, _ in intervals], lens, dst.device).run(src, dst)
