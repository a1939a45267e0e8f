"""Tiny driver for an ncu capture of the decode attention kernel (B=128, 32 heads, ctx 512)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realhf_b200.ops import attention as A
from realhf_b200.ops import functional as OF
B, nq, nkv, hd, S = 128, 32, 32, 128, 640
kc = torch.randn(B, nkv, S, hd, device="cuda", dtype=torch.bfloat16).permute(0, 2, 1, 3)
vc = torch.randn(B, nkv, S, hd, device="cuda", dtype=torch.bfloat16).permute(0, 2, 1, 3)
lens = torch.full((B,), 512, device="cuda", dtype=torch.int32)
qkv = torch.randn(B, (nq + 2 * nkv) * hd, device="cuda", dtype=torch.bfloat16)
cos, sin = OF.rope_tables(S, hd, 10000.0, "cuda")
for _ in range(5):
    A.decode_attention(qkv, kc, vc, lens, nq, nkv, hd, None, cos, sin, hd, False)
torch.cuda.synchronize()
