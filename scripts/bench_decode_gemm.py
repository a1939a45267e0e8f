"""Decode-shaped GEMMs with HBM-cold weights: cycle through 16+ distinct weight matrices (> L2) per shape."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realhf_b200.ops import gemm as G

def bench(fns, n_rounds=5):
    """All calls of one round are captured in a CUDA graph (as in the decode loop), so host launch cost is excluded."""
    for f in fns: f()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g):
            for f in fns: f()
    torch.cuda.current_stream().wait_stream(st)
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n_rounds):
        g.replay()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / (n_rounds * len(fns)) * 1e3

for M in ((16, 32) if os.environ.get("SMALL_B") else (64, 128)):
    for (N, K, name) in [(12288, 4096, "qkv"), (4096, 4096, "o"), (22016, 4096, "gate_up"), (4096, 11008, "down"), (32000, 4096, "head")]:
        copies = max(4, int(2.5e9 // (N * K * 2)))
        ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) for _ in range(copies)]
        x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        row = dict(M=M, N=N, K=K, name=name, copies=copies, roofline_us=round(N * K * 2 / 6.58e6, 1))
        row["cublas_us"] = round(bench([lambda w=w: x @ w.t() for w in ws]), 1)
        row["smallm_auto_us"] = round(bench([lambda w=w: G.gemm(x, w) for w in ws]), 1)
        for bn, sp in ((256, 0), (128, 0), (160, 1), (96, 1), (224, 1), (192, 2)):
            row[f"smallm_bn{bn}_s{sp}_us"] = round(bench([lambda w=w: G.gemm_streamk(x, w, bn=bn, split=sp) for w in ws]), 1)
        for mc in ((0, 1) if os.environ.get("SWEEP_MC") else (0,)):
            for bn in (128, 256):
                try:
                    row[f"tc_bn{bn}_mc{mc}_us"] = round(bench([lambda w=w: G.gemm(x, w, bn=bn, mc=mc) for w in ws]), 1)
                except Exception as ex:
                    row[f"tc_bn{bn}_mc{mc}_us"] = str(ex)[:40]
        print(json.dumps(row), flush=True)
        del ws
