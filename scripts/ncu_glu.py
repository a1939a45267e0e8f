"""Driver for an ncu capture of the gate|up GEMM with the SwiGLU epilogue (LLaMA-7B shapes, 20480 tokens):
    ncu --set full --clock-control none --import-source on -k regex:gemm_2cta_kernel -c 1 -o gpurun_out/gemm_glu python scripts/ncu_glu.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realhf_b200.ops import gemm as G  # noqa: E402
from realhf_b200.ops import lib  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 20480
x = (torch.randn(T, 4096, device="cuda") * 0.3).to(torch.bfloat16)
w = (torch.randn(2 * 11008, 4096, device="cuda") * 0.02).to(torch.bfloat16)
for _ in range(3):
    out = lib().gemm_glu(x, w, 0, False, G._sms(x.device))
torch.cuda.synchronize()
