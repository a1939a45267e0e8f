"""tcgen05 GEMM numerics vs fp32 PyTorch for every operand-major combination used by a linear layer."""
import pytest
import torch

from realhf_b200.ops import gemm as G

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    B = b.float() if b_mn else b.float().t()
    return A @ B


@pytest.mark.parametrize("bn", [0, 256, 128, 64, 32])
@pytest.mark.parametrize("shape", [(128, 256, 64), (256, 512, 256), (1000, 1032, 520), (4096, 4096, 4096), (77, 40, 72)])
def test_gemm_kmajor(shape, bn):
    M, N, K = shape
    torch.manual_seed(0)
    a = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    b = torch.randn(N, K, device=DEV, dtype=torch.bfloat16)
    c = G.gemm(a, b, bn=bn)
    ref = _ref(a, b, False, False)
    torch.testing.assert_close(c.float(), ref, atol=K ** 0.5 * 0.05, rtol=2e-2)


@pytest.mark.parametrize("majors", [(False, True), (True, True), (True, False)])
@pytest.mark.parametrize("shape", [(128, 256, 64), (512, 384, 256), (1000, 1032, 520), (4096, 11008, 2048)])
def test_gemm_mn_major(shape, majors):
    M, N, K = shape
    a_mn, b_mn = majors
    torch.manual_seed(1)
    a = torch.randn((K, M) if a_mn else (M, K), device=DEV, dtype=torch.bfloat16)
    b = torch.randn((K, N) if b_mn else (N, K), device=DEV, dtype=torch.bfloat16)
    c = G.gemm(a, b, a_mn=a_mn, b_mn=b_mn)
    ref = _ref(a, b, a_mn, b_mn)
    torch.testing.assert_close(c.float(), ref, atol=K ** 0.5 * 0.05, rtol=2e-2)


def test_gemm_bias_fp32_out_accumulate_fp16():
    torch.manual_seed(2)
    M, N, K = 300, 520, 264
    a = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    b = torch.randn(N, K, device=DEV, dtype=torch.bfloat16)
    bias = torch.randn(N, device=DEV, dtype=torch.bfloat16)
    c = G.gemm(a, b, bias=bias)
    torch.testing.assert_close(c.float(), _ref(a, b, False, False) + bias.float(), atol=1.0, rtol=2e-2)
    acc = torch.randn(M, N, device=DEV, dtype=torch.float32)
    acc0 = acc.clone()
    G.gemm(a, b, out=acc, accumulate=True)
    torch.testing.assert_close(acc, acc0 + _ref(a, b, False, False), atol=1e-2, rtol=1e-3)
    ah, bh = a.half(), b.half()
    ch = G.gemm(ah, bh)
    torch.testing.assert_close(ch.float(), _ref(ah, bh, False, False), atol=0.5, rtol=1e-2)


def test_linear_autograd_matches_torch():
    torch.manual_seed(3)
    T, K, N = 777, 512, 1024
    x = torch.randn(T, K, device=DEV, dtype=torch.bfloat16, requires_grad=True)
    w = (torch.randn(N, K, device=DEV) * 0.05).to(torch.bfloat16).requires_grad_(True)
    bias = torch.randn(N, device=DEV, dtype=torch.bfloat16, requires_grad=True)
    y = G.linear(x, w, bias)
    dy = torch.randn_like(y)
    y.backward(dy)
    xr, wr, br = (t.detach().float().requires_grad_(True) for t in (x, w, bias))
    yr = torch.nn.functional.linear(xr, wr, br)
    yr.backward(dy.float())
    torch.testing.assert_close(y.float(), yr, atol=0.1, rtol=2e-2)
    torch.testing.assert_close(x.grad.float(), xr.grad, atol=0.1, rtol=2e-2)
    torch.testing.assert_close(w.grad.float(), wr.grad, atol=0.5, rtol=2e-2)
    torch.testing.assert_close(bias.grad.float(), br.grad, atol=0.5, rtol=2e-2)


def test_gemm_strided_operand():
    torch.manual_seed(4)
    big = torch.randn(640, 3 * 512, device=DEV, dtype=torch.bfloat16)
    a = big[:, 512:1024]  # row pitch 1536, offset 1024 bytes
    b = torch.randn(256, 512, device=DEV, dtype=torch.bfloat16)
    c = G.gemm(a, b)
    torch.testing.assert_close(c.float(), a.float() @ b.float().t(), atol=1.5, rtol=2e-2)


@pytest.mark.parametrize("shape", [(128, 4096, 4096), (64, 12288, 4096), (128, 22016, 4096), (1, 4096, 512), (100, 11008, 1024),
                                   (128, 32000, 4096), (7, 264, 136)])
def test_gemm_small_m_multicast(shape):
    """M <= 128 takes the cluster / TMA-multicast variant (A tile shared by 4 CTAs)."""
    M, N, K = shape
    torch.manual_seed(5)
    a = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    b = torch.randn(N, K, device=DEV, dtype=torch.bfloat16)
    for _ in range(3):  # repeated launches exercise barrier phase carry-over
        c = G.gemm(a, b)
    torch.testing.assert_close(c.float(), a.float() @ b.float().t(), atol=K ** 0.5 * 0.05, rtol=2e-2)


@pytest.mark.parametrize("shape", [(128, 4096, 4096), (64, 4096, 11008), (17, 12288, 4096), (128, 22016, 4096), (1, 32000, 4096),
                                   (100, 1000, 520), (128, 256, 256), (8, 4104, 264)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_streamk_decode_gemm(shape, dtype):
    """Stream-K small-M GEMM vs fp32 reference; repeated calls check that the flags self-reset."""
    M, N, K = shape
    torch.manual_seed(M + N + K)
    a = (torch.randn(M, K, device=DEV) * 0.5).to(dtype)
    b = (torch.randn(N, K, device=DEV) * 0.05).to(dtype)
    bias = torch.randn(N, device=DEV).to(dtype)
    ref = a.float() @ b.float().t()
    for bn, split in ((0, 0), (128, 0), (48, 0), (96, 1), (256, 3), (32, 2)):
        for it in range(3):
            y = G.gemm_streamk(a, b, bn=bn, split=split)
        torch.testing.assert_close(y.float(), ref, atol=2e-2 * (K / 4096) ** 0.5 + 1e-2, rtol=2e-2)
    yb = G.gemm_streamk(a, b, bias=bias)
    torch.testing.assert_close(yb.float(), ref + bias.float(), atol=2e-2 * (K / 4096) ** 0.5 + 2e-2, rtol=2e-2)
    y32 = G.gemm_streamk(a, b, out_dtype=torch.float32)
    torch.testing.assert_close(y32, ref, atol=5e-3 * (K / 4096) ** 0.5 + 1e-3, rtol=1e-3)
    ws, flags = G.streamk_workspace(a.device)
    assert int(flags.abs().sum()) == 0
    # the default route for small M is the stream-K kernel, also under CUDA graph capture
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        G.gemm(a, b)
        with torch.cuda.graph(g):
            yg = G.gemm(a, b)
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    torch.testing.assert_close(yg.float(), ref, atol=2e-2 * (K / 4096) ** 0.5 + 1e-2, rtol=2e-2)


@pytest.mark.parametrize("majors", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("shape", [(256, 256, 64), (512, 384, 256), (1000, 1032, 520), (4096, 4096, 1024), (2048, 11008, 512), (304, 128, 136)])
@pytest.mark.parametrize("bn", [0, 128, 256])
def test_gemm_cta_pair(shape, majors, bn):
    """cta_group::2 kernel (256 x BN tiles on SM pairs), every operand-major combination, ragged edges."""
    M, N, K = shape
    a_mn, b_mn = majors
    torch.manual_seed(3)
    a = torch.randn((K, M) if a_mn else (M, K), device=DEV, dtype=torch.bfloat16)
    b = torch.randn((K, N) if b_mn else (N, K), device=DEV, dtype=torch.bfloat16)
    ref = _ref(a, b, a_mn, b_mn)
    for _ in range(2):
        c = G.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bn=bn, mc=2)
    torch.testing.assert_close(c.float(), ref, atol=K ** 0.5 * 0.05, rtol=2e-2)


def test_gemm_cta_pair_epilogues():
    torch.manual_seed(4)
    M, N, K = 600, 520, 264
    a = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    b = torch.randn(N, K, device=DEV, dtype=torch.bfloat16)
    bias = torch.randn(N, device=DEV, dtype=torch.bfloat16)
    c = G.gemm(a, b, bias=bias, mc=2)
    torch.testing.assert_close(c.float(), _ref(a, b, False, False) + bias.float(), atol=1.0, rtol=2e-2)
    acc = torch.randn(M, N, device=DEV, dtype=torch.float32)
    acc0 = acc.clone()
    G.gemm(a, b, out=acc, accumulate=True, mc=2)
    torch.testing.assert_close(acc, acc0 + _ref(a, b, False, False), atol=1e-2, rtol=1e-3)
    ah, bh = a.half(), b.half()
    torch.testing.assert_close(G.gemm(ah, bh, mc=2).float(), _ref(ah, bh, False, False), atol=0.5, rtol=1e-2)


def test_linear_wgrad_accumulates_into_flat_grad_view():
    """With a same-dtype flat gradient buffer the wgrad GEMM adds into the parameter's .grad view (two micro-batches)."""
    torch.manual_seed(5)
    T, K, N = 512, 256, 384
    w = (torch.randn(N, K, device=DEV) * 0.05).to(torch.bfloat16).requires_grad_(True)
    flat = torch.zeros(N * K, device=DEV, dtype=torch.bfloat16)
    w.grad = flat.view(N, K)
    w._grad_in_flat_buffer = True
    ref = torch.zeros(N, K, device=DEV)
    for _ in range(2):
        x = torch.randn(T, K, device=DEV, dtype=torch.bfloat16)
        dy = torch.randn(T, N, device=DEV, dtype=torch.bfloat16) * 0.1
        G.linear(x, w).backward(dy)
        ref += dy.float().t() @ x.float()
    assert w.grad.data_ptr() == flat.data_ptr()
    torch.testing.assert_close(flat.view(N, K).float(), ref, atol=0.2, rtol=3e-2)


@pytest.mark.parametrize("counts", [[130, 0, 257, 64, 5, 1024], [8, 8, 8, 8], [0, 0, 300, 0], [513]])
def test_grouped_gemm_moe(counts):
    """One-launch grouped GEMM over expert row groups (empty and ragged groups) incl. its backward."""
    torch.manual_seed(7)
    E, K, N = len(counts), 256, 384
    M = sum(counts)
    x = (torch.randn(M, K, device=DEV) * 0.5).to(torch.bfloat16).requires_grad_(True)
    w = (torch.randn(E, N, K, device=DEV) * 0.05).to(torch.bfloat16).requires_grad_(True)
    off = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), device=DEV, dtype=torch.int32)
    y = G.grouped_linear(x, w, off)
    dy = (torch.randn(M, N, device=DEV) * 0.1).to(torch.bfloat16)
    y.backward(dy)
    xr, wr = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    outs, o = [], 0
    for e, n in enumerate(counts):
        outs.append(xr[o:o + n] @ wr[e].t())
        o += n
    yr = torch.cat(outs)
    yr.backward(dy.float())
    torch.testing.assert_close(y.float(), yr, atol=0.1, rtol=2e-2)
    torch.testing.assert_close(x.grad.float(), xr.grad, atol=0.05, rtol=3e-2)
    torch.testing.assert_close(w.grad.float(), wr.grad, atol=0.3, rtol=3e-2)


@pytest.mark.parametrize("counts,M,N", [([300, 0, 64, 1000, 5, 77], 256, 512), ([4096, 4096], 1024, 384), ([0, 0, 9], 128, 128)])
def test_grouped_wgrad_single_launch(counts, M, N):
    from realhf_b200.ops import gemm as G
    torch.manual_seed(0)
    T, ng = sum(counts), len(counts)
    offsets = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), dtype=torch.int32, device="cuda")
    dy = torch.randn(T, M, device="cuda", dtype=torch.bfloat16)
    x = torch.randn(T, N, device="cuda", dtype=torch.bfloat16)
    ref = G.grouped_wgrad_ref(dy, x, offsets, ng)
    out = G.grouped_wgrad(dy, x, offsets, ng)
    torch.testing.assert_close(out.float(), ref, atol=2e-2 * max(counts) ** 0.5, rtol=2e-2)
    acc = torch.ones(ng, M, N, device="cuda", dtype=torch.float32)
    G.grouped_wgrad(dy, x, offsets, ng, out=acc, accumulate=True)
    torch.testing.assert_close(acc, ref + 1.0, atol=2e-2 * max(counts) ** 0.5, rtol=2e-2)


@pytest.mark.parametrize("shape", [(300, 1024, 512), (4096, 4096, 11008), (129, 512, 136)])
@pytest.mark.parametrize("kind", ["silu", "gelu_pytorch_tanh"])
def test_gated_linear_fused_epilogue_matches_reference(shape, kind):
    """SwiGLU / GeGLU in the epilogue of the CTA-pair GEMM (gate rows from CTA 0, up rows from CTA 1 of the fused weight):
    forward equals gated_act(x @ W^T), backward (through the saved raw projections) equals autograd of the fp32 reference."""
    from realhf_b200.ops import functional as OF
    from realhf_b200.ops import gemm as G
    OF.set_gemm_impl(G.linear)
    M, K, F_ = shape
    torch.manual_seed(0)
    x = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16).requires_grad_(True)
    w = (torch.randn(2 * F_, K, device="cuda") * (1.0 / K ** 0.5)).to(torch.bfloat16).requires_grad_(True)
    assert G.gated_linear_supported(x, w)
    a = OF.gated_linear(x, w, kind)
    da = torch.randn_like(a)
    a.backward(da)
    xr, wr = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    gu = (xr @ wr.t()).to(torch.bfloat16).float() + 0 * (xr @ wr.t())  # the projection is rounded to bf16 before the activation
    gur = xr @ wr.t()
    ar = OF.gated_act_ref(gur, kind)
    ar.backward(da.float())
    torch.testing.assert_close(a.float(), ar, atol=3e-2, rtol=3e-2)
    torch.testing.assert_close(x.grad.float(), xr.grad, atol=6e-2, rtol=6e-2)
    torch.testing.assert_close(w.grad.float(), wr.grad, atol=0.25 if M > 1000 else 8e-2, rtol=6e-2)
    # inference: no raw projection is written
    with torch.no_grad():
        a2 = OF.gated_linear(x.detach(), w.detach(), kind)
    torch.testing.assert_close(a2, a.detach())
