"""Tiny driver for ncu captures of the tcgen05 attention kernels: `python scripts/ncu_attn.py fwd|bwd [seqs] [len]`.

    ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -c 1 -o gpurun_out/attn_fwd \
        python scripts/ncu_attn.py fwd
    ncu -i gpurun_out/attn_fwd.ncu-rep --page raw --csv > profiles/ncu_attn_fwd_raw.csv      (read here, after the run)
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realhf_b200.ops import lib  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
seqs = int(sys.argv[2]) if len(sys.argv) > 2 else 32
L = int(sys.argv[3]) if len(sys.argv) > 3 else 640
nq = nkv = 32
hd = 128
torch.manual_seed(0)
T = seqs * L
cu = torch.arange(0, T + 1, L, dtype=torch.int32, device="cuda")
qkv = torch.randn(T, (nq + 2 * nkv) * hd, device="cuda", dtype=torch.bfloat16)
q = qkv[:, : nq * hd].view(T, nq, hd)
k = qkv[:, nq * hd:(nq + nkv) * hd].view(T, nkv, hd)
v = qkv[:, (nq + nkv) * hd:].view(T, nkv, hd)
scale = 1.0 / math.sqrt(hd)
for _ in range(3):
    out, lse = lib().attn_fwd(q, k, v, cu, L, scale, True)
if which == "bwd":
    dout = torch.randn_like(out)
    dqkv = torch.empty_like(qkv)
    dq = dqkv[:, : nq * hd].view(T, nq, hd)
    dk = dqkv[:, nq * hd:(nq + nkv) * hd].view(T, nkv, hd)
    dv = dqkv[:, (nq + nkv) * hd:].view(T, nkv, hd)
    for _ in range(3):
        lib().attn_bwd(dout, q, k, v, out, lse, dq, dk, dv, cu, L, scale, True)
torch.cuda.synchronize()
