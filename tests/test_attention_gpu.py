"""Decode attention kernel (fused RoPE + KV append + split-KV softmax) vs the PyTorch reference."""
import pytest
import torch

from realhf_b200.ops import attention as A
from realhf_b200.ops import functional as OF

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("cfg", [(128, 32, 32, 128, 640), (3, 32, 8, 128, 2048), (5, 8, 1, 64, 700), (16, 16, 16, 64, 300),
                                 (2, 32, 4, 128, 4096)])
@pytest.mark.parametrize("rope", [None, "half", "interleaved"])
@pytest.mark.parametrize("layout", ["bshd", "bhsd"])
def test_decode_attention(cfg, rope, layout):
    B, nq, nkv, hd, S = cfg
    torch.manual_seed(0)
    if layout == "bshd":
        kc = torch.randn(B, S, nkv, hd, device=DEV, dtype=torch.bfloat16)
        vc = torch.randn(B, S, nkv, hd, device=DEV, dtype=torch.bfloat16)
    else:
        kc = torch.randn(B, nkv, S, hd, device=DEV, dtype=torch.bfloat16).permute(0, 2, 1, 3)
        vc = torch.randn(B, nkv, S, hd, device=DEV, dtype=torch.bfloat16).permute(0, 2, 1, 3)
    lens = torch.randint(0, S - 1, (B,), device=DEV, dtype=torch.int32)
    lens[0] = 0
    lens[-1] = S - 1
    qkv = torch.randn(B, (nq + 2 * nkv) * hd, device=DEV, dtype=torch.bfloat16)
    cos = sin = None
    if rope:
        cos, sin = OF.rope_tables(S, hd, 10000.0, DEV)
    kc_ref, vc_ref = kc.clone(), vc.clone()
    out = A.decode_attention(qkv, kc, vc, lens, nq, nkv, hd, None, cos, sin, hd, rope == "interleaved")
    # reference: same op through the CPU/PyTorch path
    x = qkv
    if rope:
        x = OF.rope_ref(qkv, cos, sin, lens, nq + nkv, hd, hd, rope == "interleaved")
    q = x[:, : nq * hd].reshape(B, nq, hd)
    k = x[:, nq * hd:(nq + nkv) * hd].reshape(B, nkv, hd)
    v = x[:, (nq + nkv) * hd:].reshape(B, nkv, hd)
    idx = torch.arange(B, device=DEV)
    kc_ref[idx, lens.long()] = k
    vc_ref[idx, lens.long()] = v
    ref = A.decode_attention_ref(q, kc_ref, vc_ref, lens + 1, hd ** -0.5).reshape(B, nq * hd)
    torch.testing.assert_close(out.float(), ref.float(), atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(kc.float(), kc_ref.float(), atol=1e-2, rtol=1e-2)
    assert torch.equal(vc, vc_ref)


@pytest.mark.parametrize("heads", [(8, 8), (8, 2)])
def test_packed_qkv_attention_forward_backward(heads):
    """Fused-QKV varlen attention (one d(qkv) buffer written by the backward kernel) vs the fp32 reference on sliced q/k/v."""
    nq, nkv = heads
    hd = 64
    torch.manual_seed(1)
    lens = [37, 128, 5, 200]
    T = sum(lens)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), device=DEV, dtype=torch.int32)
    qkv = (torch.randn(T, (nq + 2 * nkv) * hd, device=DEV) * 0.5).to(torch.bfloat16).requires_grad_(True)
    out = A.varlen_attention_qkv(qkv, cu, max(lens), nq, nkv, hd)
    dout = torch.randn_like(out)
    out.backward(dout)
    g = qkv.grad.clone()
    x = qkv.detach().float().requires_grad_(True)
    q = x[:, : nq * hd].view(T, nq, hd)
    k = x[:, nq * hd:(nq + nkv) * hd].view(T, nkv, hd)
    v = x[:, (nq + nkv) * hd:].view(T, nkv, hd)
    ref = A.varlen_attention_ref(q, k, v, cu, hd ** -0.5, True)
    ref.backward(dout.float())
    torch.testing.assert_close(out.float(), ref, atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(g.float(), x.grad, atol=3e-2, rtol=3e-2)


@pytest.mark.parametrize("hd,nq,nkv,causal", [(128, 8, 8, True), (128, 8, 2, True), (64, 4, 4, True), (128, 4, 4, False)])
def test_attn_fwd_tcgen05_matches_reference(hd, nq, nkv, causal):
    import math

    import numpy as np

    from realhf_b200.ops import lib
    torch.manual_seed(0)
    lens = [1, 37, 128, 129, 300, 640, 5]
    T = sum(lens)
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device=DEV)
    qkv = torch.randn(T, (nq + 2 * nkv) * hd, device=DEV, dtype=torch.bfloat16)
    q = qkv[:, : nq * hd].view(T, nq, hd)
    k = qkv[:, nq * hd:(nq + nkv) * hd].view(T, nkv, hd)
    v = qkv[:, (nq + nkv) * hd:].view(T, nkv, hd)
    scale = 1.0 / math.sqrt(hd)
    out, lse = lib().attn_fwd(q, k, v, cu, max(lens), scale, causal)
    ref = A.varlen_attention_ref(q.float(), k.float(), v.float(), cu, scale, causal)
    torch.testing.assert_close(out.float(), ref, atol=2e-2, rtol=2e-2)
    # LSE [nq, T]: natural log of the softmax denominator with the scale applied
    rep = nq // nkv
    for s, e in zip(cu[:-1].tolist(), cu[1:].tolist()):
        att = torch.einsum("qhd,khd->hqk", q[s:e].float(), k[s:e].float().repeat_interleave(rep, 1)) * scale
        if causal:
            att = att.masked_fill(torch.ones(e - s, e - s, device=DEV, dtype=torch.bool).triu(1), float("-inf"))
        torch.testing.assert_close(lse[:, s:e], torch.logsumexp(att, -1), atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("hd,nq,nkv,causal", [(128, 8, 8, True), (128, 8, 2, True), (64, 4, 4, True), (128, 4, 4, False)])
def test_attn_bwd_tcgen05_matches_autograd_of_the_reference(hd, nq, nkv, causal):
    """dq / dk / dv written into the three column ranges of one d(qkv) buffer vs autograd through the fp32 reference."""
    import math

    import numpy as np

    from realhf_b200.ops import lib
    torch.manual_seed(0)
    lens = [1, 37, 128, 129, 300, 640, 5]
    T = sum(lens)
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device=DEV)
    W = (nq + 2 * nkv) * hd
    qkv = torch.randn(T, W, device=DEV, dtype=torch.bfloat16)
    dout = torch.randn(T, nq, hd, device=DEV, dtype=torch.bfloat16)
    scale = 1.0 / math.sqrt(hd)

    def views(t):
        return (t[:, : nq * hd].view(T, nq, hd), t[:, nq * hd:(nq + nkv) * hd].view(T, nkv, hd), t[:, (nq + nkv) * hd:].view(T, nkv, hd))
    # reference: fp32 autograd, forward statistics from the same math
    x = qkv.float().requires_grad_(True)
    qf, kf, vf = views(x)
    ref = A.varlen_attention_ref(qf, kf, vf, cu, scale, causal)
    ref.backward(dout.float())
    q, k, v = views(qkv)
    rep = nq // nkv
    lse = torch.empty(nq, T, device=DEV, dtype=torch.float32)
    for s, e in zip(cu[:-1].tolist(), cu[1:].tolist()):
        att = torch.einsum("qhd,khd->hqk", q[s:e].float(), k[s:e].float().repeat_interleave(rep, 1)) * scale
        if causal:
            att = att.masked_fill(torch.ones(e - s, e - s, device=DEV, dtype=torch.bool).triu(1), float("-inf"))
        lse[:, s:e] = torch.logsumexp(att, -1)
    dqkv = torch.full_like(qkv, float("nan"))   # every element must be written
    dq, dk, dv = views(dqkv)
    lib().attn_bwd(dout, q, k, v, ref.detach().to(torch.bfloat16), lse, dq, dk, dv, cu, max(lens), scale, causal)
    assert torch.isfinite(dqkv.float()).all()
    torch.testing.assert_close(dqkv.float(), x.grad, atol=6e-2, rtol=6e-2)
