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
    # in-place variant: the producer writes into symmetric memory, regions alternate per call, one barrier per call
    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.parallel.fused_tp import FusedTP
    from realhf_b200.ops import gemm as G
    ctx = ParallelContext.build(ProcessTopology(1, 1, world), list(range(world)), rank, backend="nccl")
    f = FusedTP(ctx, max_tokens=256, max_features=4096, device=dev)
    for it in range(5):
        xs = (torch.randn(128, 2048, device=dev) * 0.3).to(torch.bfloat16)
        ws = (torch.randn(4096, 2048, device=dev) * 0.05).to(torch.bfloat16)
        ref = xs.float() @ ws.float().t()
        dist.all_reduce(ref)
        y = f.gemm_ar(xs, ws)
        torch.testing.assert_close(y.float(), ref, atol=0.15, rtol=3e-2)
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

    out_y = torch.empty_like(y)
    t_ours = timeit(lambda: sb.all_reduce(y, out=out_y, algo=1))
    t_nccl = timeit(lambda: dist.all_reduce(y))
    ysym = f.symm_out(128, 1024, torch.bfloat16)

    def symm_call():
        f._ar_calls += 1  # alternate regions as a producer would
        f.all_reduce_symm(ysym)
    t_symm = timeit(symm_call)
    xs = torch.randn(128, 2048, device=dev, dtype=torch.bfloat16)
    ws = torch.randn(4096, 2048, device=dev, dtype=torch.bfloat16)
    t_gemm_ar = timeit(lambda: f.gemm_ar(xs, ws))
    t_gemm = timeit(lambda: G.gemm(xs, ws))
    return dict(us_ours=t_ours, us_nccl=t_nccl, us_symm_inplace=t_symm, us_gemm_ar_1MB=t_gemm_ar, us_gemm_alone=t_gemm)


def test_symmetric_allreduce_matches_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from realhf_b200.base.testing import run_distributed
    world = min(torch.cuda.device_count(), 8)
    res = run_distributed(_ar_worker, world, backend="nccl", timeout=300)
    print("all-reduce 256KiB bf16 (us):", res[0])


def _fused_worker(rank, world):
    import torch.distributed as dist

    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.ops import functional as OF
    from realhf_b200.ops import gemm as G
    from realhf_b200.parallel.fused_tp import FusedTP
    from realhf_b200.parallel import tp as TP
    dev = torch.device("cuda", rank)
    OF.set_gemm_impl(G.linear)
    ctx = ParallelContext.build(ProcessTopology(1, 1, world), list(range(world)), rank, backend="nccl", sequence_parallel=True)
    T, H, F = 4096, 1024, 2816
    f = FusedTP(ctx, max_tokens=T, max_features=max(H, 2 * F // world, F // world), device=dev)
    torch.manual_seed(7)
    out = {}
    # ---- GEMM -> reduce-scatter (row-parallel, K sharded)
    x_full = (torch.randn(T, F, device=dev) * 0.5).to(torch.bfloat16)
    w_full = (torch.randn(H, F, device=dev) * 0.05).to(torch.bfloat16)
    kl = F // world
    x = x_full[:, rank * kl:(rank + 1) * kl].contiguous().requires_grad_(True)
    w = w_full[:, rank * kl:(rank + 1) * kl].contiguous().requires_grad_(True)
    ref_full = x_full.float() @ w_full.float().t()
    rows = T // world
    for it in range(3):  # repeated calls exercise the parity double-buffering and the device-side call counters
        y = f.gemm_rs(x, w)
    torch.cuda.synchronize()
    torch.testing.assert_close(y.float(), ref_full[rank * rows:(rank + 1) * rows], atol=0.25, rtol=3e-2)
    # backward of gemm_rs = all-gather -> GEMM of the gradient
    dy_local = (torch.randn(rows, H, device=dev) * 0.1).to(torch.bfloat16)
    y.backward(dy_local)
    dy_full = torch.empty(T, H, device=dev, dtype=torch.bfloat16)
    dist.all_gather_into_tensor(dy_full, dy_local)
    torch.testing.assert_close(x.grad.float(), dy_full.float() @ w.float(), atol=0.1, rtol=3e-2)
    torch.testing.assert_close(w.grad.float(), dy_full.float().t() @ x.detach().float(), atol=0.5, rtol=3e-2)
    # ---- all-gather -> GEMM (column-parallel, N sharded)
    xl = (torch.randn(rows, H, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True)
    wc = (torch.randn(2 * F // world, H, device=dev) * 0.05).to(torch.bfloat16).requires_grad_(True)
    for it in range(3):
        yc = f.ag_gemm(xl, wc)
    xg = torch.empty(T, H, device=dev, dtype=torch.bfloat16)
    dist.all_gather_into_tensor(xg, xl.detach())
    torch.testing.assert_close(yc.float(), xg.float() @ wc.float().t(), atol=0.25, rtol=3e-2)
    dyc = (torch.randn_like(yc) * 0.1)
    yc.backward(dyc)
    dx_full = dyc.float() @ wc.float()
    dist.all_reduce(dx_full)
    torch.testing.assert_close(xl.grad.float(), dx_full[rank * rows:(rank + 1) * rows], atol=0.25, rtol=3e-2)
    # ---- timing vs GEMM + NCCL
    def timeit(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    Tb, Hb, Fb = 16384, 4096, 11008
    fb = FusedTP(ctx, max_tokens=Tb, max_features=max(Hb, 2 * Fb // world), device=dev)
    xb = torch.randn(Tb, Fb // world, device=dev, dtype=torch.bfloat16)
    wb = torch.randn(Hb, Fb // world, device=dev, dtype=torch.bfloat16)
    with torch.no_grad():
        t_fused = timeit(lambda: fb._gemm_rs_raw(xb, wb, False))
        t_base = timeit(lambda: TP._reduce_scatter_first_dim(G.gemm(xb, wb), ctx))
        xl2 = torch.randn(Tb // world, Hb, device=dev, dtype=torch.bfloat16)
        wc2 = torch.randn(2 * Fb // world, Hb, device=dev, dtype=torch.bfloat16)
        t_fused_ag = timeit(lambda: fb._ag_gemm_raw(xl2, wc2, False))
        t_base_ag = timeit(lambda: G.gemm(TP._gather_first_dim(xl2, ctx), wc2))
    return dict(gemm_rs_ms=t_fused, gemm_nccl_rs_ms=t_base, ag_gemm_ms=t_fused_ag, nccl_ag_gemm_ms=t_base_ag)


def test_fused_tp_gemm_collectives():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from realhf_b200.base.testing import run_distributed
    world = 2
    res = run_distributed(_fused_worker, world, backend="nccl", timeout=400)
    print("fused TP (ms), rank 0:", res[0])


def _realloc_direct_worker(rank, world):
    """tp2 -> dp2 (and back) parameter reallocation where every transfer is ONE segment-copy kernel storing straight into
    the destination GPU's flat buffer (CUDA IPC), checked against the exact destination shards."""
    import torch.distributed as dist

    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.models import hf_io
    from realhf_b200.models.real_model import ReaLModel
    from realhf_b200.parallel import realloc
    from realhf_b200.parallel.symm_mem import SymmetricBuffer
    dev = torch.device("cuda", rank)
    cfg = hf_io.family("llama").make_test_config()
    cfg.n_layers = 4
    for (src_dims, dst_dims) in (((1, 1, 2), (1, 2, 1)), ((1, 2, 1), (1, 1, 2))):
        s_topo, d_topo = ProcessTopology(*src_dims), ProcessTopology(*dst_dims)
        src = ReaLModel(cfg, ParallelContext.fake(s_topo, rank), dtype=torch.bfloat16, device=dev).instantiate(seed=9)
        ref = ReaLModel(cfg, ParallelContext.fake(d_topo, rank), dtype=torch.bfloat16, device=dev).instantiate(seed=9)
        plan = realloc.derive_plan(cfg, s_topo, [0, 1], d_topo, [0, 1])
        exe = realloc.ReallocExecutor(plan, rank, 2, dev)
        sb = SymmetricBuffer(plan.dst_numel[rank] * 2, device=dev)
        dst_flat = sb.data()[: plan.dst_numel[rank] * 2].view(torch.bfloat16)
        dst_flat.fill_(float("nan"))
        torch.cuda.synchronize(); dist.barrier()
        exe.run(src.flat_param.data, dst_flat, peer_dst_ptrs={w: sb.data_ptrs[w] for w in range(world)})
        torch.cuda.synchronize(); dist.barrier()   # every peer's stores have landed
        assert torch.equal(dst_flat, ref.flat_param.data), (src_dims, dst_dims, (dst_flat != ref.flat_param.data).sum())
        # same plan through the pack / isend-irecv / unpack fallback
        dst2 = torch.full_like(ref.flat_param.data, float("nan"))
        exe.run(src.flat_param.data, dst2)
        torch.cuda.synchronize()
        assert torch.equal(dst2, ref.flat_param.data)
    return True


def test_realloc_direct_peer_stores():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from realhf_b200.base.testing import run_distributed
    assert all(run_distributed(_realloc_direct_worker, 2, backend="nccl", timeout=300))


def test_full_runtime_ppo_with_direct_realloc(tmp_path):
    """Launcher -> master + 2 model workers on 2 GPUs (NCCL): PPO where generation (dp2) and actor training (tp2) use
    different layouts, so the actor's weights are reallocated around every generation by direct peer stores."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import os
    import uuid

    import importlib
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import fixtures
    os.environ["PYTHONPATH"] = os.path.dirname(here) + os.pathsep + os.environ.get("PYTHONPATH", "")
    os.environ["REAL_FILEROOT"] = str(tmp_path / "fileroot")
    os.environ["REAL_NAME_RESOLVE_ROOT"] = str(tmp_path / "nr")
    from realhf_b200.base import constants, name_resolve
    importlib.reload(constants)
    name_resolve.reconfigure("nfs", record_root=str(tmp_path / "nr"))
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    actor, critic = str(tmp_path / "actor"), str(tmp_path / "critic")
    cfg, tok, words = fixtures.make_checkpoint(actor, "llama")
    fixtures.make_checkpoint(critic, "llama", is_critic=True, seed=5)
    data = str(tmp_path / "prompts.jsonl")
    fixtures.write_prompt_dataset(data, words, n=32)
    args = ["ppo", f"experiment_name=ppo-{uuid.uuid4().hex[:6]}", "trial_name=t0", "device=cuda", "dtype=bf16", "n_gpus_per_node=2",
            "allocation_mode=manual", f"dataset.path={data}", "dataset.train_bs_n_seqs=8", "dataset.max_prompt_len=16",
            "ppo.gen.max_new_tokens=6", "ppo.gen.min_new_tokens=2", "ppo.gen.top_k=20", "ppo.ppo_n_minibatches=2",
            "exp_ctrl.total_train_epochs=1", "exp_ctrl.benchmark_steps=2"]
    for role, path in (("actor", actor), ("ref", actor), ("critic", critic), ("rew", critic)):
        args += [f"{role}.type._class=llama", f"{role}.path={path}", f"{role}.gradient_checkpointing=false"]
    args += ["actor_gen.parallel.data_parallel_size=2", "actor_train.parallel.model_parallel_size=2",
             "critic_train.parallel.data_parallel_size=2", "critic_inf.parallel.data_parallel_size=2",
             "ref_inf.parallel.data_parallel_size=2", "rew_inf.parallel.data_parallel_size=2"]
    exp = build_experiment(args)
    main_start(exp, timeout=600)
    root = os.path.join(os.environ["REAL_FILEROOT"], "logs", exp.experiment_name, "t0")
    log = open(os.path.join(root, "master_worker-0")).read()
    assert log.count("[actor_train]") == 2 and "benchmark finished" in log, log[-3000:]
