"""Varlen attention forward: own tcgen05 kernel vs the flash-attn library kernel (CUDA-event timed, device only).

    python scripts/bench_attn.py [--seqs 32 --len 640 --nq 32 --nkv 32 --hd 128 --iters 20]

Prints one JSON line per implementation: ms, TFLOP/s (causal FLOPs = 2 * nq * hd * sum L^2) and max |diff| vs the library.
"""
import argparse
import json
import math

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs", type=int, default=32)
    ap.add_argument("--len", type=int, default=640)
    ap.add_argument("--nq", type=int, default=32)
    ap.add_argument("--nkv", type=int, default=32)
    ap.add_argument("--hd", type=int, default=128)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    from flash_attn import flash_attn_varlen_func

    from realhf_b200.ops import lib
    dev = "cuda"
    torch.manual_seed(0)
    lens = [a.len] * a.seqs
    T = sum(lens)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    qkv = torch.randn(T, (a.nq + 2 * a.nkv) * a.hd, device=dev, dtype=torch.bfloat16)
    q = qkv[:, : a.nq * a.hd].view(T, a.nq, a.hd)
    k = qkv[:, a.nq * a.hd:(a.nq + a.nkv) * a.hd].view(T, a.nkv, a.hd)
    v = qkv[:, (a.nq + a.nkv) * a.hd:].view(T, a.nkv, a.hd)
    scale = 1.0 / math.sqrt(a.hd)
    flops = 2.0 * a.nq * a.hd * sum(l * l for l in lens)  # QK^T + PV, causal half
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    ref = flash_attn_varlen_func(q, k, v, cu, cu, a.len, a.len, softmax_scale=scale, causal=True)
    ms = timed(lambda: flash_attn_varlen_func(q, k, v, cu, cu, a.len, a.len, softmax_scale=scale, causal=True))
    print(json.dumps({"impl": "flash-attn", "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1)}))
    out = lib().attn_fwd(q, k, v, cu, a.len, scale, True)[0]
    ms = timed(lambda: lib().attn_fwd(q, k, v, cu, a.len, scale, True))
    print(json.dumps({"impl": "tcgen05", "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1),
                      "max_abs_diff_vs_lib": float((out.float() - ref.float()).abs().max())}))


if __name__ == "__main__":
    main()
