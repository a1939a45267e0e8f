import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realhf_b200.ops import attention as A
from realhf_b200.ops import functional as OF
B, nq, nkv, hd, S = 128, 32, 32, 128, 640
kc = torch.randn(B, nkv, S, hd, device="cuda", dtype=torch.bfloat16).permute(0, 2, 1, 3)
vc = torch.randn(B, nkv, S, hd, device="cuda", dtype=torch.bfloat16).permute(0, 2, 1, 3)
qkv = torch.randn(B, (nq + 2 * nkv) * hd, device="cuda", dtype=torch.bfloat16)
cos, sin = OF.rope_tables(S, hd, 10000.0, "cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for L in (128, 384, 639):
    lens = torch.full((B,), L, device="cuda", dtype=torch.int32)
    ts = []
    for i in range(13):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); A.decode_attention(qkv, kc, vc, lens, nq, nkv, hd, None, cos, sin, hd, False); e.record()
        torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    t = sorted(ts[3:])[len(ts[3:]) // 2]
    byts = 2 * B * nkv * (L + 1) * hd * 2
    print(json.dumps(dict(op="decode_attention", B=B, ctx=L + 1, ms=round(t, 4), GBps=round(byts / t / 1e6, 1), l2="flushed")))
