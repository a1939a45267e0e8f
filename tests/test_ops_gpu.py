"""Numerics of every sm_100a kernel against a plain PyTorch fp32 reference of the same op."""
import pytest
import torch

from realhf_b200.ops import functional as OF

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _packed_case(bs=37, max_len=700, seed=0):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, max_len, (bs,), generator=g)
    cu = torch.zeros(bs + 1, dtype=torch.int32)
    cu[1:] = lens.cumsum(0)
    total = int(cu[-1])
    return lens, cu, total, g


def test_gae_1d_matches_reference():
    lens, cu, total, g = _packed_case()
    bs = lens.numel()
    rewards = torch.randn(total, generator=g)
    values = torch.randn(total + bs, generator=g)
    boot = torch.rand(bs, generator=g) > 0.5
    a_ref, r_ref = OF.gae_1d_misalign_ref(rewards, values, cu, boot, 0.99, 0.95)
    a, r = OF.gae_1d_misalign(rewards.to(DEV), values.to(DEV), cu.to(DEV), boot.to(DEV), 0.99, 0.95)
    torch.testing.assert_close(a.cpu(), a_ref, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(r.cpu(), r_ref, atol=1e-4, rtol=1e-4)


def test_ppo_rewards_gae_fused():
    lens, cu, total, g = _packed_case(bs=64, max_len=600, seed=1)
    bs = lens.numel()
    logp, ref = torch.randn(total, generator=g), torch.randn(total, generator=g)
    scores = torch.randn(bs, generator=g) * 10
    values = torch.randn(total + bs, generator=g)
    no_eos = torch.rand(bs, generator=g) > 0.7
    ref_out = OF.ppo_rewards_gae_ref(logp, ref, scores, values, cu, no_eos, 1.0, 0.95, 0.1, 5.0)
    out = OF.ppo_rewards_gae(logp.to(DEV), ref.to(DEV), scores.to(DEV), values.to(DEV), cu.to(DEV), no_eos.to(DEV),
                             1.0, 0.95, 0.1, 5.0)
    for o, r in zip(out, ref_out):
        torch.testing.assert_close(o.cpu(), r, atol=2e-4, rtol=1e-4)


@pytest.mark.parametrize("mode", ["olp", "nolp"])
def test_gae_2d(mode):
    g = torch.Generator().manual_seed(2)
    bs, T = 33, 257
    rewards, values = torch.randn(bs, T, generator=g), torch.randn(bs, T + 1, generator=g)
    dones = torch.rand(bs, T + 1, generator=g) > 0.95
    truncs = (torch.rand(bs, T + 1, generator=g) > 0.5) & dones
    a_ref, r_ref = OF.gae_2d_ref(rewards, values, dones, truncs, 0.98, 0.9, mode)
    a, r = OF.gae_2d(rewards.to(DEV), values.to(DEV), dones.to(DEV), truncs.to(DEV), 0.98, 0.9, mode)
    torch.testing.assert_close(a.cpu(), a_ref, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(r.cpu(), r_ref, atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("H", [256, 4096, 5120, 8192])
def test_rmsnorm_fwd_bwd(dtype, H):
    torch.manual_seed(0)
    x = torch.randn(300, H, device=DEV, dtype=dtype, requires_grad=True)
    w = (torch.randn(H, device=DEV) * 0.1 + 1).to(dtype).requires_grad_(True)
    y = OF.rmsnorm(x, w, 1e-5, 0.0)
    dy = torch.randn_like(y)
    y.backward(dy)
    xr, wr = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    yr = OF.rmsnorm_ref(xr, wr, 1e-5, 0.0)
    yr.backward(dy.float())
    tol = dict(atol=2e-2, rtol=2e-2) if dtype != torch.float32 else dict(atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(y.float(), yr, **tol)
    torch.testing.assert_close(x.grad.float(), xr.grad, **tol)
    torch.testing.assert_close(w.grad.float(), wr.grad, atol=0.3 if dtype != torch.float32 else 1e-3, rtol=3e-2)


def test_add_rmsnorm_gemma_offset():
    torch.manual_seed(0)
    x = torch.randn(128, 2048, device=DEV, dtype=torch.bfloat16)
    res = torch.randn_like(x)
    w = torch.randn(2048, device=DEV, dtype=torch.bfloat16) * 0.1
    with torch.no_grad():
        y, r = OF.add_rmsnorm(x, res, w, 1e-6, 1.0)
    yr, rr = OF.rmsnorm_ref(x, w, 1e-6, 1.0, residual=res)
    torch.testing.assert_close(r, rr)
    torch.testing.assert_close(y.float(), yr.float(), atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("w_off", [0.0, 1.0])
def test_add_rmsnorm_training_forward_and_fused_backward(dtype, w_off):
    """(h, x_new) = (rmsnorm(x + d), x + d) with autograd: both outputs are used downstream (the norm feeds the block, the sum
    continues as the residual stream), so the backward kernel must add the residual-stream gradient to the norm gradient."""
    torch.manual_seed(0)
    H = 4096
    d = torch.randn(257, H, device=DEV, dtype=dtype, requires_grad=True)
    x = torch.randn(257, H, device=DEV, dtype=dtype, requires_grad=True)
    w = (torch.randn(H, device=DEV) * 0.1 + (0.0 if w_off else 1.0)).to(dtype).requires_grad_(True)
    h, xn = OF.add_rmsnorm(d, x, w, 1e-5, w_off)
    gh, gx = torch.randn_like(h), torch.randn_like(xn)
    torch.autograd.backward([h, xn], [gh, gx])
    dr, xr, wr = (t.detach().float().requires_grad_(True) for t in (d, x, w))
    s = (dr + xr) if dtype == torch.float32 else (dr + xr).to(dtype).float() + 0 * (dr + xr)  # the stream is stored in `dtype`
    sr = dr + xr
    hr = sr * torch.rsqrt(sr.pow(2).mean(-1, keepdim=True) + 1e-5) * (wr + w_off)
    torch.autograd.backward([hr, sr], [gh.float(), gx.float()])
    tol = dict(atol=3e-2, rtol=3e-2) if dtype != torch.float32 else dict(atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(h.float(), hr, **tol)
    torch.testing.assert_close(xn.float(), sr, **tol)
    torch.testing.assert_close(d.grad.float(), dr.grad, **tol)
    torch.testing.assert_close(x.grad.float(), xr.grad, **tol)
    torch.testing.assert_close(w.grad.float(), wr.grad, atol=0.4 if dtype != torch.float32 else 2e-3, rtol=3e-2)
    # only the residual output used (e.g. a stage boundary): gradient passes straight through
    d2, x2 = d.detach().clone().requires_grad_(True), x.detach().clone().requires_grad_(True)
    _, xn2 = OF.add_rmsnorm(d2, x2, w.detach().requires_grad_(True), 1e-5, w_off)
    xn2.backward(gx)
    torch.testing.assert_close(d2.grad, gx)
    torch.testing.assert_close(x2.grad, gx)


@pytest.mark.parametrize("interleaved", [False, True])
def test_rope_inplace_and_grad(interleaved):
    torch.manual_seed(0)
    T, nq, nkv, hd = 513, 8, 2, 128
    row = (nq + 2 * nkv) * hd
    qkv = torch.randn(T, row, device=DEV, dtype=torch.bfloat16)
    pos = torch.randint(0, 2048, (T,), device=DEV, dtype=torch.int32)
    cos, sin = OF.rope_tables(2048, hd, 10000.0, DEV)
    ref = OF.rope_ref(qkv, cos, sin, pos, nq + nkv, hd, hd, interleaved)
    out = OF.apply_rope(qkv.clone(), cos, sin, pos, nq + nkv, hd, hd, interleaved)
    torch.testing.assert_close(out.float(), ref.float(), atol=2e-2, rtol=2e-2)
    assert torch.equal(out[:, (nq + nkv) * hd:], qkv[:, (nq + nkv) * hd:])
    # gradient = inverse rotation
    x = qkv.clone().requires_grad_(True)
    y = OF.apply_rope(x * 1.0, cos, sin, pos, nq + nkv, hd, hd, interleaved)
    dy = torch.randn_like(y)
    y.backward(dy)
    gref = OF.rope_ref(dy, cos, sin, pos, nq + nkv, hd, hd, interleaved, inverse=True)
    torch.testing.assert_close(x.grad.float(), gref.float(), atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("kind", ["silu", "gelu_pytorch_tanh"])
def test_gated_act(kind):
    torch.manual_seed(0)
    gu = torch.randn(777, 2 * 1408, device=DEV, dtype=torch.bfloat16, requires_grad=True)
    out = OF.gated_act(gu, kind)
    d = torch.randn_like(out)
    out.backward(d)
    gr = gu.detach().float().requires_grad_(True)
    outr = OF.gated_act_ref(gr, kind)
    outr.backward(d.float())
    torch.testing.assert_close(out.float(), outr, atol=3e-2, rtol=3e-2)
    torch.testing.assert_close(gu.grad.float(), gr.grad, atol=3e-2, rtol=3e-2)


@pytest.mark.parametrize("V", [32000, 50257])
@pytest.mark.parametrize("masked", [False, True])
def test_logprob_from_logits(V, masked):
    torch.manual_seed(0)
    T = 257
    logits = (torch.randn(T, V, device=DEV) * 3).to(torch.bfloat16)
    labels = torch.randint(0, V, (T,), device=DEV)
    bits = None
    if masked:
        m = torch.rand(T, V, device=DEV) > 0.3
        m[torch.arange(T), labels] = False
        bits = OF.pack_mask_bits(m)
        assert torch.equal(OF.unpack_mask_bits(bits, V), m)
    lp, lse = OF.logprob_from_logits(logits, labels, bits, 1 / 0.7)
    lpr, lser = OF.logprob_from_logits_ref(logits, labels, bits, 1 / 0.7)
    torch.testing.assert_close(lp, lpr, atol=2e-3, rtol=1e-3)
    torch.testing.assert_close(lse, lser, atol=2e-3, rtol=1e-3)


def test_lm_head_logprobs_fwd_bwd():
    torch.manual_seed(0)
    T, H, V = 1000, 512, 32000
    h = (torch.randn(T, H, device=DEV) * 0.5).to(torch.bfloat16).requires_grad_(True)
    w = (torch.randn(V, H, device=DEV) * 0.05).to(torch.bfloat16).requires_grad_(True)
    labels = torch.randint(0, V, (T,), device=DEV)
    lp = OF.lm_head_logprobs(h, w, labels, chunk=384)
    d = torch.randn(T, device=DEV)
    lp.backward(d)
    hr, wr = h.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    lpr = OF.lm_head_logprobs_ref(hr, wr, labels)
    lpr.backward(d)
    torch.testing.assert_close(lp, lpr, atol=5e-2, rtol=2e-2)
    torch.testing.assert_close(h.grad.float(), hr.grad, atol=5e-2, rtol=5e-2)
    torch.testing.assert_close(w.grad.float(), wr.grad, atol=5e-2, rtol=5e-2)


@pytest.mark.parametrize("cfg", [("bf16", "bf16", "fp32", True), ("bf16", "bf16", "bf16", False), ("fp32", "fp32", "fp32", False),
                                 ("bf16", "fp32", "fp32", True)])
def test_adamw_matches_reference(cfg):
    dt = {"bf16": torch.bfloat16, "fp32": torch.float32}
    pd, gd, sd, use_master = dt[cfg[0]], dt[cfg[1]], dt[cfg[2]], cfg[3]
    torch.manual_seed(0)
    n = 1_000_003
    p = torch.randn(n, device=DEV).to(pd)
    g = (torch.randn(n, device=DEV) * 1e-2).to(gd)
    m, v = torch.zeros(n, device=DEV, dtype=sd), torch.zeros(n, device=DEV, dtype=sd)
    master = p.float().clone() if use_master else None
    pr, mr, vr = p.clone(), m.clone(), v.clone()
    masterr = master.clone() if use_master else None
    scale = torch.tensor(0.5, device=DEV)
    skip = torch.zeros(1, dtype=torch.int32, device=DEV)
    for step in (1, 2, 3):
        OF.adamw_step(p, g, m, v, master, 1e-3, 0.9, 0.95, 1e-5, 0.05, step, scale, skip)
        OF.adamw_ref(pr, g, mr, vr, masterr, 1e-3, 0.9, 0.95, 1e-5, 0.05, step, 0.5)
    tol = dict(atol=1e-5, rtol=1e-4) if pd == torch.float32 else dict(atol=1e-2, rtol=1e-2)
    torch.testing.assert_close(p.float(), pr.float(), **tol)
    if use_master:
        torch.testing.assert_close(master, masterr, atol=1e-5, rtol=1e-4)
    # skip flag suppresses the update
    skip.fill_(1)
    before = p.clone()
    OF.adamw_step(p, g, m, v, master, 1e-3, 0.9, 0.95, 1e-5, 0.05, 4, scale, skip)
    assert torch.equal(before, p)


def test_sumsq_and_nonfinite():
    g = torch.randn(3_000_001, device=DEV).to(torch.bfloat16)
    out = torch.zeros(2, device=DEV)
    OF.sumsq_accum(g, out)
    torch.testing.assert_close(out[0], g.float().pow(2).sum(), rtol=1e-3, atol=1)
    assert out[1].item() == 0
    g[12345] = float("inf")
    out.zero_()
    OF.sumsq_accum(g, out)
    assert out[1].item() == 1


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_interval_ops_match_python(dtype):
    torch.manual_seed(0)
    n = 20_000_000
    src = torch.randn(n, device=DEV).to(dtype)
    starts = torch.randperm(n // 1000)[:5000].sort().values * 1000
    lens = torch.randint(1, 999, (5000,))
    iv = [(int(s), int(s + l)) for s, l in zip(starts, lens)]
    perm = torch.randperm(len(iv)).tolist()
    iv = [iv[i] for i in perm]
    packed = OF.slice_intervals(src, iv)
    ref = torch.cat([src[a:b] for a, b in iv])
    assert torch.equal(packed, ref)
    dst = torch.zeros_like(src)
    OF.set_intervals(packed, dst, iv)
    dref = torch.zeros_like(src)
    for a, b in iv:
        dref[a:b] = src[a:b]
    assert torch.equal(dst, dref)


def test_segment_copy_ema():
    src = torch.randn(1 << 20, device=DEV).to(torch.bfloat16)
    dst = torch.randn(1 << 20, device=DEV).to(torch.bfloat16)
    ref = (0.3 * src.float() + 0.7 * dst.float()).to(torch.bfloat16)
    plan = OF.SegmentPlan([0], [0], [src.numel() * 2], DEV)
    plan.run(src, dst, eta=0.3)
    torch.testing.assert_close(dst.float(), ref.float(), atol=1e-2, rtol=1e-2)


@pytest.mark.parametrize("H,dtype", [(768, torch.bfloat16), (1024, torch.float16), (4096, torch.bfloat16), (1600, torch.float32)])
def test_layernorm_native_matches_reference(H, dtype, monkeypatch):
    from realhf_b200.ops import functional as OF
    monkeypatch.setenv("REAL_LAYERNORM", "native")
    torch.manual_seed(0)
    x = (torch.randn(517, H, device="cuda") * 2 + 0.5).to(dtype).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(H, device="cuda")).to(dtype).requires_grad_(True)
    b = (0.1 * torch.randn(H, device="cuda")).to(dtype).requires_grad_(True)
    dy = torch.randn(517, H, device="cuda").to(dtype)
    y = OF.layer_norm(x, w, b, 1e-5)
    y.backward(dy)
    xr, wr, br = (t.detach().float().requires_grad_(True) for t in (x, w, b))
    yr = torch.nn.functional.layer_norm(xr, (H,), wr, br, 1e-5)
    yr.backward(dy.float())
    tol = dict(atol=2e-2, rtol=2e-2) if dtype != torch.float32 else dict(atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(y.float(), yr, **tol)
    torch.testing.assert_close(x.grad.float(), xr.grad, **tol)
    torch.testing.assert_close(w.grad.float(), wr.grad, atol=tol["atol"] * 20, rtol=tol["rtol"])
    torch.testing.assert_close(b.grad.float(), br.grad, atol=tol["atol"] * 20, rtol=tol["rtol"])
