"""Peer-memory collectives on >= 2 GPUs: symmetric buffers, barrier, one-/two-shot all-reduce (also inside a CUDA graph)."""
import os
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.distributed]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _ar_worker(rank, world):
    import torch.distributed as dist

    from realhf_b200.parallel.symm_mem import SymmetricBuffer
    dev = torch.device("cuda", rank)
    sb = SymmetricBuffer(64 << 20, device=dev)
    res = {}
    for dtype in (torch.bfloat16, torch.float32):
        for n in (8, 4096, 1 << 20, (1 << 22) + 8):
            for algo in (1, 2):
                g = torch.Generator(device=dev).manual_seed(100 * rank + n % 97)
                x = torch.randn(n, device=dev, dtype=torch.float32, generator=g).to(dtype)
                ref = x.clone().float()
                dist.all_reduce(ref)
                out = sb.all_reduce(x, algo=algo)
                err = (out.float() - ref).abs().max().item()
                tol = 1e-4 if dtype == torch.float32 else 0.1
                assert err <= tol * max(1.0, ref.abs().max().item()), (dtype, n, algo, err)
    # repeated calls + CUDA graph capture
    x = torch.ones(1 << 16, device=dev, dtype=torch.bfloat16) * (rank + 1)
    out = torch.empty_like(x)
    for _ in range(3):
        sb.all_reduce(x, out=out, algo=1)
    torch.cuda.synchronize()
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            sb.all_reduce(x, out=out, algo=1)
            sb.all_reduce(out, out=out, algo=2)
    for it in range(4):
        x.fill_(float(rank + 1 + it))
        graph.replay()
        torch.cuda.synchronize()
        expect = sum(r + 1 + it for r in range(world)) * world
        assert torch.all(out.float() == expect), (it, out[:4], expect)
    # timing vs NCCL for the decode-sized message (256 KiB)
    y = torch.randn(128 * 1024, device=dev, dtype=torch.bfloat16)

    def timeit(f, n=200):
        for _ in range(10):
            f()
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            f()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n * 1e3

    t_ours = timeit(lambda: sb.all_reduce(y, out=out[: y.numel()], algo=1))
    t_nccl = timeit(lambda: dist.all_reduce(y))
    return dict(us_ours=t_ours, us_nccl=t_nccl)


def test_symmetric_allreduce_matches_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from realhf_b200.base.testing import run_distributed
    world = min(torch.cuda.device_count(), 8)
    res = run_distributed(_ar_worker, world, backend="nccl", timeout=300)
    print("all-reduce 256KiB bf16 (us):", res[0])
