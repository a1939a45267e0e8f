import torch, time, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realhf_b200.ops import gemm as G
def bench(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
res = []
for (M, N, K) in [(8192, 8192, 8192), (16384, 4096, 4096), (16384, 22016, 4096), (16384, 4096, 11008), (16384, 12288, 4096), (16384, 32000, 4096)]:
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16); b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    t_ref = bench(lambda: a @ b.t())
    row = dict(M=M, N=N, K=K, cublas_ms=round(t_ref, 4), cublas_tflops=round(2 * M * N * K / t_ref / 1e9, 1))
    if M > 128:
        for bn in (256, 128):
            t = bench(lambda: G.gemm(a, b, bn=bn, mc=2))
            row[f"pair_bn{bn}_ms"] = round(t, 4); row[f"pair_bn{bn}_tflops"] = round(2 * M * N * K / t / 1e9, 1)
    for bn in ([0] if M > 128 else [0]):
        try:
            t = bench(lambda: G.gemm(a, b, bn=bn, mc=0))
            row[f"ours_bn{bn}_ms"] = round(t, 4); row[f"ours_bn{bn}_tflops"] = round(2 * M * N * K / t / 1e9, 1)
        except Exception as ex:
            row[f"ours_bn{bn}"] = str(ex)[:80]
    print(json.dumps(row)); res.append(row)
# dgrad / wgrad majors
M, N, K = 16384, 4096, 4096
dy = torch.randn(M, N, device="cuda", dtype=torch.bfloat16); w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16); x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
t = bench(lambda: G.gemm(dy, w, b_mn=True, mc=0)); tr = bench(lambda: dy @ w); t2 = bench(lambda: G.gemm(dy, w, b_mn=True, mc=2))
print(json.dumps(dict(kind="dgrad", ours_ms=round(t, 4), pair_ms=round(t2, 4), cublas_ms=round(tr, 4), ours_tflops=round(2*M*N*K/t/1e9, 1), pair_tflops=round(2*M*N*K/t2/1e9, 1), cublas_tflops=round(2*M*N*K/tr/1e9, 1))))
t = bench(lambda: G.gemm(dy, x, a_mn=True, b_mn=True, mc=0)); tr = bench(lambda: dy.t() @ x); t2 = bench(lambda: G.gemm(dy, x, a_mn=True, b_mn=True, mc=2))
print(json.dumps(dict(kind="wgrad", ours_ms=round(t, 4), pair_ms=round(t2, 4), cublas_ms=round(tr, 4), ours_tflops=round(2*M*N*K/t/1e9, 1), pair_tflops=round(2*M*N*K/t2/1e9, 1), cublas_tflops=round(2*M*N*K/tr/1e9, 1))))
