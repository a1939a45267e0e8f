"""A small tour of the sm_100a kernels for `compute-sanitizer` (tiny shapes: the tools slow kernels down 10-100x):

    compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_kernels.py > profiles/sanitizer_racecheck.log
    compute-sanitizer --tool memcheck  python scripts/sanitize_kernels.py > profiles/sanitizer_memcheck.log

Covers: tile GEMM, CTA-pair GEMM (+ GLU epilogue), small-M stream-K GEMM, RMSNorm fwd / fused add fwd+bwd, segment copy, sampling,
GAE, AdamW + sum-of-squares, decode attention, varlen attention forward / backward, single-rank all-reduce barrier path.
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realhf_b200.ops import functional as OF  # noqa: E402
from realhf_b200.ops import gemm as G  # noqa: E402
from realhf_b200.ops import lib  # noqa: E402

OF.set_gemm_impl(G.linear)
dev = "cuda"
torch.manual_seed(0)
bf = torch.bfloat16
which = set(sys.argv[1:]) or {"gemm", "norm", "misc", "attn"}
done = []
if "gemm" in which:
    a = torch.randn(200, 256, device=dev).to(bf)
    w = torch.randn(384, 256, device=dev).to(bf)
    y = G.gemm(a, w)                                   # tile / pair kernel
    torch.testing.assert_close(y.float(), a.float() @ w.float().t(), atol=0.5, rtol=5e-2)
    a2 = torch.randn(520, 512, device=dev).to(bf)
    w2 = torch.randn(2 * 256, 512, device=dev).to(bf)
    act = lib().gemm_glu(a2, w2, 0, True, 148)[0]       # CTA-pair GLU epilogue
    a3 = torch.randn(16, 1024, device=dev).to(bf)
    w3 = torch.randn(1024, 1024, device=dev).to(bf)
    y3 = G.gemm(a3, w3)                                 # stream-K small-M
    torch.testing.assert_close(y3.float(), a3.float() @ w3.float().t(), atol=1.0, rtol=5e-2)
    done.append("gemm")
if "norm" in which:
    x = torch.randn(33, 1024, device=dev).to(bf).requires_grad_(True)
    d = torch.randn(33, 1024, device=dev).to(bf).requires_grad_(True)
    wn = torch.ones(1024, device=dev).to(bf).requires_grad_(True)
    h, xn = OF.add_rmsnorm(d, x, wn, 1e-5, 0.0)
    (h.float().sum() + xn.float().sum()).backward()
    done.append("norm")
if "misc" in which:
    logits = torch.randn(8, 32000, device=dev)
    lib().sample(logits, None, 50, 0.9, 1.0, 2, True, False, 0, 1, 0, True)
    p = torch.randn(4096, device=dev).to(bf)
    g = torch.randn(4096, device=dev).to(bf)
    m = torch.zeros(4096, device=dev).to(bf)
    v = torch.zeros(4096, device=dev).to(bf)
    OF.adamw_step(p, g, m, v, None, 1e-3, 0.9, 0.95, 1e-5, 0.0, 1, None, None, stochastic=True, seed=3)
    st = torch.zeros(2, device=dev)
    OF.sumsq_accum(g, st)
    gu = torch.randn(64, 512, device=dev).to(bf)
    OF.gated_act(gu, "silu")
    done.append("misc")
if "attn" in which:
    nq = nkv = 4
    hd = 128
    lens = [70, 130, 9]
    T = sum(lens)
    cu = torch.tensor([0, 70, 200, 209], dtype=torch.int32, device=dev)
    qkv = torch.randn(T, (nq + 2 * nkv) * hd, device=dev).to(bf)
    q = qkv[:, : nq * hd].view(T, nq, hd)
    k = qkv[:, nq * hd:(nq + nkv) * hd].view(T, nkv, hd)
    vv = qkv[:, (nq + nkv) * hd:].view(T, nkv, hd)
    out, lse = lib().attn_fwd(q, k, vv, cu, 130, 1.0 / math.sqrt(hd), True)
    dqkv = torch.empty_like(qkv)
    lib().attn_bwd(torch.randn_like(out), q, k, vv, out, lse, dqkv[:, : nq * hd].view(T, nq, hd), dqkv[:, nq * hd:(nq + nkv) * hd].view(T, nkv, hd),
                   dqkv[:, (nq + nkv) * hd:].view(T, nkv, hd), cu, 130, 1.0 / math.sqrt(hd), True)
    done.append("attn")
torch.cuda.synchronize()
print("sanitize tour finished:", done)
