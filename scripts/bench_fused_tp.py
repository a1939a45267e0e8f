"""Fused TP GEMM+collective kernels vs GEMM + NCCL at LLaMA-7B shapes (run with torchrun, 2+ GPUs)."""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realhf_b200.base.topology import ParallelContext, ProcessTopology
from realhf_b200.ops import functional as OF
from realhf_b200.ops import gemm as G
from realhf_b200.parallel.fused_tp import FusedTP

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
OF.set_gemm_impl(G.linear)
ctx = ParallelContext.build(ProcessTopology(1, 1, world), list(range(world)), rank, backend="nccl", sequence_parallel=True)
T, H, F = 16384, 4096, 11008
f = FusedTP(ctx, max_tokens=T, max_features=max(H, 2 * F // world), device=dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t), 4)


rows = T // world
# row-parallel down-proj: x [T, F/w] @ w [H, F/w]^T -> reduce-scatter over tokens
x = torch.randn(T, F // world, device=dev, dtype=torch.bfloat16) * 0.1
w = torch.randn(H, F // world, device=dev, dtype=torch.bfloat16) * 0.02
out_rs = torch.empty(rows, H, device=dev, dtype=torch.bfloat16)


def base_rs():
    y = G.gemm(x, w)
    dist.reduce_scatter_tensor(out_rs, y, group=ctx.tp_group)


res = dict(world=world, T=T, H=H, F=F)
res["gemm_rs_fused_ms"] = timeit(lambda: f._gemm_rs_raw(x, w, False))
res["gemm_then_nccl_rs_ms"] = timeit(base_rs)
res["gemm_only_ms"] = timeit(lambda: G.gemm(x, w))
# column-parallel gate_up: all-gather x_local [T/w, H] -> @ w [2F/w, H]^T
xl = torch.randn(rows, H, device=dev, dtype=torch.bfloat16) * 0.1
wc = torch.randn(2 * F // world, H, device=dev, dtype=torch.bfloat16) * 0.02
xg = torch.empty(T, H, device=dev, dtype=torch.bfloat16)


def base_ag():
    dist.all_gather_into_tensor(xg, xl, group=ctx.tp_group)
    return G.gemm(xg, wc)


res["ag_gemm_fused_ms"] = timeit(lambda: f._ag_gemm_raw(xl, wc, False))
res["nccl_ag_then_gemm_ms"] = timeit(base_ag)
res["gemm_only_ag_shape_ms"] = timeit(lambda: G.gemm(xg, wc))
y1 = f._ag_gemm_raw(xl, wc, False); y2 = base_ag()
res["ag_max_abs_diff"] = float((y1.float() - y2.float()).abs().max())
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
