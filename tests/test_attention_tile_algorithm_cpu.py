"""Tile-level emulation of the tcgen05 attention kernels (csrc/attn_fwd_tcgen05.cu, attn_bwd_tcgen05.cu) in PyTorch.

The kernels could not be run when they were written, so their ALGORITHM -- tile ranges, mask predicates, the deferred fold of
`T_j = P_j V_j` into the running output, the LSE convention, the transposed dKV pass with per-column statistics, the 64-row
streamed blocks, the GQA head loop -- is mirrored here statement by statement (same names, same predicates, bf16 rounding of
P / dS where the kernels round) and checked against the plain fp32 reference and its autograd.  A mistake in any of those
shows up here; what this cannot check is the hardware protocol (barriers, descriptors, swizzles)."""
import math

import pytest
import torch

from realhf_b200.ops.attention import varlen_attention_ref

LOG2E = 1.4426950408889634
kBQ = kBKV = 128   # forward tiles
kR, kX = 128, 64   # backward: resident rows, streamed rows


def _bf16(x):
    return x.to(torch.bfloat16).float()


def emulate_fwd(q, k, v, cu, scale, causal):
    """One "CTA" per (sequence, head, 128-row q tile); returns out [T,nq,D] and lse [nq,T]."""
    T, nq, D = q.shape
    nkv = k.shape[1]
    out = torch.zeros(T, nq, D)
    lse = torch.zeros(nq, T)
    sl2 = scale * LOG2E
    for seq in range(len(cu) - 1):
        tok0, L = cu[seq], cu[seq + 1] - cu[seq]
        for head in range(nq):
            hk = head // (nq // nkv)
            for qt in range((L + kBQ - 1) // kBQ):
                q0 = qt * kBQ
                kv_len = min(L, q0 + kBQ) if causal else L
                n_kv = (kv_len + kBKV - 1) // kBKV
                rows = torch.arange(kBQ)
                q_pos = q0 + rows
                # TMA box: rows past the end of the packed tensor are zero-filled, rows of the next sequence are real data
                Q = torch.zeros(kBQ, D)
                n_in = min(kBQ, T - (tok0 + q0))
                Q[:n_in] = q[tok0 + q0: tok0 + q0 + n_in, head]
                m_run = torch.full((kBQ,), -math.inf)
                l_run = torch.zeros(kBQ)
                alpha_prev = torch.zeros(kBQ)
                O = torch.zeros(kBQ, D)
                T_tiles = []
                for j in range(n_kv):
                    kv0 = j * kBKV
                    K = torch.zeros(kBKV, D)
                    V = torch.zeros(kBKV, D)
                    n_in = min(kBKV, T - (tok0 + kv0))
                    K[:n_in] = k[tok0 + kv0: tok0 + kv0 + n_in, hk]
                    V[:n_in] = v[tok0 + kv0: tok0 + kv0 + n_in, hk]
                    S = Q @ K.t()
                    need_mask = (kv0 + kBKV > L) or (causal and kv0 + kBKV - 1 > q0)
                    lim = (torch.minimum(torch.tensor(L), q_pos + 1) if causal else torch.full((kBQ,), L)) - kv0
                    cols = torch.arange(kBKV)
                    masked = need_mask & (cols[None, :] >= lim[:, None])
                    mx = S.masked_fill(masked, -math.inf).max(dim=1).values
                    m_new = torch.maximum(m_run, mx)
                    m_eff = torch.where(torch.isinf(m_new) & (m_new < 0), torch.zeros_like(m_new), m_new)
                    alpha = torch.exp2((m_run - m_eff) * sl2)
                    P = torch.exp2(S * sl2 - (m_eff * sl2)[:, None]).masked_fill(masked, 0.0)
                    l_run = l_run * alpha + P.sum(1)
                    m_run = m_new
                    T_tiles.append(_bf16(P) @ V)          # P is rounded to bf16 before the second MMA
                    if j > 0:                             # fold of T_{j-1} is deferred by one iteration
                        O = O * alpha_prev[:, None] + T_tiles[j - 1]
                    alpha_prev = alpha
                O = O * alpha_prev[:, None] + T_tiles[n_kv - 1]
                ok = q_pos < L
                out[tok0 + q_pos[ok], head] = (O / l_run[:, None])[ok]
                lse[head, tok0 + q_pos[ok]] = (m_run * scale + torch.log(l_run))[ok]
    return out, lse


def _delta(out, dout):
    return (out * dout).sum(-1).t().contiguous()  # [nq, T]


def _load_rows(x, row0, n, head, T):
    """TMA box of n rows starting at row0 of x[:, head]: zero fill past the end of the packed tensor."""
    D = x.shape[-1]
    tile = torch.zeros(n, D)
    n_in = max(0, min(n, T - row0))
    tile[:n_in] = x[row0: row0 + n_in, head]
    return tile


def emulate_bwd(q, k, v, out, dout, lse, cu, scale, causal):
    T, nq, D = q.shape
    nkv = k.shape[1]
    G = nq // nkv
    delta = _delta(out, dout)
    sl2 = scale * LOG2E
    dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    for seq in range(len(cu) - 1):
        tok0, L = cu[seq], cu[seq + 1] - cu[seq]
        nblk_all = (L + kX - 1) // kX
        for tile in range((L + kR - 1) // kR):
            r0 = tile * kR
            r_pos = r0 + torch.arange(kR)
            # ---------------- dKV pass: lanes are keys
            for head in range(nkv):
                Kt, Vt = _load_rows(k, tok0 + r0, kR, head, T), _load_rows(v, tok0 + r0, kR, head, T)
                blk0 = r0 // kX if causal else 0
                nblk = nblk_all - blk0
                accV, accK = torch.zeros(kR, D), torch.zeros(kR, D)
                for t in range(nblk * G):
                    h, blk = head * G + t // nblk, blk0 + t % nblk
                    x0 = blk * kX
                    Qx, dOx = _load_rows(q, tok0 + x0, kX, h, T), _load_rows(dout, tok0 + x0, kX, h, T)
                    qpos = x0 + torch.arange(kX)
                    inb = qpos < L
                    lse_c = torch.full((kX,), math.inf)
                    del_c = torch.zeros(kX)
                    lse_c[inb] = lse[h, tok0 + qpos[inb]] * LOG2E
                    del_c[inb] = delta[h, tok0 + qpos[inb]]
                    St, dPt = Kt @ Qx.t(), Vt @ dOx.t()
                    Pt = torch.exp2(St * sl2 - lse_c[None, :])
                    need_mask = causal and (r0 + kR - 1 > x0)
                    if need_mask:
                        Pt = Pt.masked_fill(~(r_pos[:, None] <= qpos[None, :]), 0.0)
                    dSt = Pt * (dPt - del_c[None, :]) * scale
                    accV += _bf16(Pt) @ dOx
                    accK += _bf16(dSt) @ Qx
                ok = r_pos < L
                dv[tok0 + r_pos[ok], head] = accV[ok]
                dk[tok0 + r_pos[ok], head] = accK[ok]
            # ---------------- dQ pass: lanes are queries
            for head in range(nq):
                hk = head // G
                Qt, dOt = _load_rows(q, tok0 + r0, kR, head, T), _load_rows(dout, tok0 + r0, kR, head, T)
                ok = r_pos < L
                lse_r = torch.full((kR,), math.inf)
                del_r = torch.zeros(kR)
                lse_r[ok] = lse[head, tok0 + r_pos[ok]] * LOG2E
                del_r[ok] = delta[head, tok0 + r_pos[ok]]
                blk1 = (min(L, r0 + kR) + kX - 1) // kX if causal else nblk_all
                acc = torch.zeros(kR, D)
                for t in range(blk1):
                    x0 = t * kX
                    Kx, Vx = _load_rows(k, tok0 + x0, kX, hk, T), _load_rows(v, tok0 + x0, kX, hk, T)
                    xpos = x0 + torch.arange(kX)
                    S, dP = Qt @ Kx.t(), dOt @ Vx.t()
                    P = torch.exp2(S * sl2 - lse_r[:, None])
                    need_mask = (x0 + kX > L) or (causal and x0 + kX - 1 > r0)
                    if need_mask:
                        vis = (xpos[None, :] < L) & ((xpos[None, :] <= r_pos[:, None]) if causal else torch.ones(kR, kX, dtype=torch.bool))
                        P = P.masked_fill(~vis, 0.0)
                    dS = P * (dP - del_r[:, None]) * scale
                    acc += _bf16(dS) @ Kx
                dq[tok0 + r_pos[ok], head] = acc[ok]
    return dq, dk, dv


CASES = [((1, 37, 128, 129, 300, 5), 4, 4, 64, True), ((200, 64, 257), 4, 2, 64, True), ((130, 70), 2, 2, 128, False),
         ((1, 2, 3), 2, 1, 64, True)]


@pytest.mark.parametrize("lens,nq,nkv,hd,causal", CASES)
def test_forward_tile_algorithm_matches_reference(lens, nq, nkv, hd, causal):
    torch.manual_seed(0)
    T = sum(lens)
    cu = [0]
    for l in lens:
        cu.append(cu[-1] + l)
    q, k, v = _bf16(torch.randn(T, nq, hd)), _bf16(torch.randn(T, nkv, hd)), _bf16(torch.randn(T, nkv, hd))
    scale = 1.0 / math.sqrt(hd)
    out, lse = emulate_fwd(q, k, v, cu, scale, causal)
    ref = varlen_attention_ref(q, k, v, torch.tensor(cu), scale, causal)
    torch.testing.assert_close(out, ref, atol=2e-2, rtol=2e-2)
    rep = nq // nkv
    for s, e in zip(cu[:-1], cu[1:]):
        att = torch.einsum("qhd,khd->hqk", q[s:e], k[s:e].repeat_interleave(rep, 1)) * scale
        if causal:
            att = att.masked_fill(torch.ones(e - s, e - s, dtype=torch.bool).triu(1), -math.inf)
        torch.testing.assert_close(lse[:, s:e], torch.logsumexp(att, -1), atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("lens,nq,nkv,hd,causal", CASES)
def test_backward_tile_algorithm_matches_autograd(lens, nq, nkv, hd, causal):
    torch.manual_seed(1)
    T = sum(lens)
    cu = [0]
    for l in lens:
        cu.append(cu[-1] + l)
    q, k, v = (_bf16(torch.randn(T, h, hd)).requires_grad_(True) for h in (nq, nkv, nkv))
    dout = _bf16(torch.randn(T, nq, hd))
    scale = 1.0 / math.sqrt(hd)
    ref = varlen_attention_ref(q, k, v, torch.tensor(cu), scale, causal)
    ref.backward(dout)
    with torch.no_grad():
        out, lse = emulate_fwd(q, k, v, cu, scale, causal)
        dq, dk, dv = emulate_bwd(q.detach(), k.detach(), v.detach(), out, dout, lse, cu, scale, causal)
    for name, got, want in (("dq", dq, q.grad), ("dk", dk, k.grad), ("dv", dv, v.grad)):
        torch.testing.assert_close(got, want, atol=3e-2, rtol=3e-2, msg=lambda m, n=name: f"{n}: {m}")
