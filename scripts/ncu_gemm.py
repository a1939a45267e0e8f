"""Tiny driver for ncu captures: `python scripts/ncu_gemm.py big|small|decode_attn`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realhf_b200.ops import gemm as G

which = sys.argv[1] if len(sys.argv) > 1 else "big"
torch.manual_seed(0)
if which == "big":
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    for _ in range(6):
        G.gemm(a, b)
elif which == "small":
    ws = [torch.randn(4096, 4096, device="cuda", dtype=torch.bfloat16) for _ in range(8)]
    x = torch.randn(128, 4096, device="cuda", dtype=torch.bfloat16)
    for w in ws:
        G.gemm(x, w)
torch.cuda.synchronize()
