import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realhf_b200.ops import gemm as G
for (M, N, K) in [(128, 4096, 4096), (64, 4096, 4096), (128, 22016, 4096)]:
    ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) for _ in range(8)]
    x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    for bn in (256, 128):
        dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
        for w in ws: G.gemm_streamk(x, w, bn=bn)
        torch.cuda.synchronize()
        G.gemm_streamk(x, ws[0], bn=bn, dbg=dbg)
        torch.cuda.synchronize()
        d = dbg.view(148, 8).cpu()
        t0 = d[:, 0].min()
        names = ["start", "first_acc_ready", "partials_posted", "bar0_passed", "red0_done", "bar1_passed", "red1_done"]
        print(M, N, K, "bn", bn)
        for i, n in enumerate(names):
            col = d[:, i]; col = col[col > 0]
            if len(col): print(f"  {n:18s} min {int(col.min()-t0)/1e3:7.2f} us  median {int(col.median()-t0)/1e3:7.2f}  max {int(col.max()-t0)/1e3:7.2f}  (n={len(col)})")
