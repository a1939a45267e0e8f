"""Decode-shaped GEMMs with HBM-cold weights: cycle through 16+ distinct weight matrices (> L2) per shape."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realhf_b200.ops import gemm as G

def bench(fns, n_rounds=3):
    for f in fns: f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n_rounds):
        for f in fns: f()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / (n_rounds * len(fns)) * 1e3

for M in (64, 128):
    for (N, K, name) in [(12288, 4096, "qkv"), (4096, 4096, "o"), (22016, 4096, "gate_up"), (4096, 11008, "down"), (32000, 4096, "head")]:
        copies = max(4, int(2.5e9 // (N * K * 2)))
        ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) for _ in range(copies)]
        x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        row = dict(M=M, N=N, K=K, name=name, copies=copies, roofline_us=round(N * K * 2 / 6.58e6, 1))
        row["cublas_us"] = round(bench([lambda w=w: x @ w.t() for w in ws]), 1)
        for mc in (0, 1):
            os.environ["REAL_GEMM_MULTICAST_FORCE"] = str(mc)
            for bn in (32, 64, 128, 256):
                try:
                    row[f"tc_bn{bn}_mc{mc}_us"] = round(bench([lambda w=w: G.gemm(x, w, bn=bn, mc=mc) for w in ws]), 1)
                except Exception as ex:
                    row[f"tc_bn{bn}_mc{mc}_us"] = str(ex)[:40]
        print(json.dumps(row), flush=True)
        del ws
