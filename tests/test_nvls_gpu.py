"""NVSwitch multicast (NVLS) paths on >= 2 GPUs: VMM symmetric memory, `multimem` all-reduce (also in a CUDA graph), and the
ZeRO-1 optimizer whose gradient reduce-scatter / AdamW / parameter all-gather run as in-switch kernels (csrc/nvls.cu).
Numerics are compared with plain fp32 PyTorch / NCCL results."""
import os
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.distributed]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _need(n):
    if not torch.cuda.is_available() or torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


def _nvls_ar_worker(rank, world):
    import torch.distributed as dist

    from realhf_b200.parallel.symm_mem import VmmSymmetricBuffer, multicast_supported
    dev = torch.device("cuda", rank)
    if not multicast_supported(dev):
        return dict(skipped="no multicast support")
    sb = VmmSymmetricBuffer(64 << 20, device=dev)
    assert sb.mc_ptr != 0
    # unicast peer mappings work: write my rank into my buffer, read every peer's
    sb.data(dtype=torch.int32)[:4] = rank + 1
    torch.cuda.synchronize(); dist.barrier()
    for r in range(world):
        assert int(sb.data(r, dtype=torch.int32)[0]) == r + 1
    dist.barrier()
    for dtype in (torch.bfloat16, torch.float32, torch.float16):
        for n in (8, 4096, 1 << 17, 1 << 20, (1 << 22) + 8):
            for mode in (1, 2):
                g = torch.Generator(device=dev).manual_seed(100 * rank + n % 97)
                x = torch.randn(n, device=dev, dtype=torch.float32, generator=g).to(dtype)
                ref = x.clone().float()
                dist.all_reduce(ref)
                out = sb.nvls_all_reduce(x, mode=mode)
                err = (out.float() - ref).abs().max().item()
                tol = 1e-4 if dtype == torch.float32 else 0.1
                assert err <= tol * max(1.0, ref.abs().max().item()), (dtype, n, mode, err)
    # CUDA graph replay (epoch barriers advance on the device)
    x = torch.ones(1 << 16, device=dev, dtype=torch.bfloat16) * (rank + 1)
    out = torch.empty_like(x)
    for _ in range(2):
        sb.nvls_all_reduce(x, out=out, mode=1)
    torch.cuda.synchronize()
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            sb.nvls_all_reduce(x, out=out, mode=1)
            sb.nvls_all_reduce(out, out=out, mode=2)
    for it in range(4):
        x.fill_(float(rank + 1 + it))
        graph.replay()
        torch.cuda.synchronize()
        expect = sum(r + 1 + it for r in range(world)) * world
        assert torch.all(out.float() == expect), (it, out[:4], expect)

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

    times = {}
    for kib in (64, 256, 1024, 8192):
        y = torch.randn(kib * 512, device=dev, dtype=torch.bfloat16)
        o = torch.empty_like(y)
        times[kib] = dict(nvls_1shot_us=round(timeit(lambda: sb.nvls_all_reduce(y, out=o, mode=1)), 2),
                          nvls_2shot_us=round(timeit(lambda: sb.nvls_all_reduce(y, out=o, mode=2)), 2),
                          peer_1shot_us=round(timeit(lambda: sb.all_reduce(y, out=o, algo=1)), 2),
                          nccl_us=round(timeit(lambda: dist.all_reduce(y)), 2))
    # ZeRO building blocks at bandwidth-relevant sizes: reduce-scatter + all-gather of 512 MiB vs NCCL
    big = VmmSymmetricBuffer(512 << 20, device=dev)
    gbuf = big.data(dtype=torch.bfloat16)
    n_el = gbuf.numel()
    per = n_el // world
    src = (torch.randn(n_el, device=dev) * 0.1).to(torch.bfloat16)
    stats = torch.zeros(2, device=dev)
    gbuf.copy_(src)
    torch.cuda.synchronize(); dist.barrier()
    big.reduce_scatter_(rank * per * 2, per * 2, torch.bfloat16, 1.0 / world, stats)
    ref = src.clone().float()
    dist.all_reduce(ref)
    mine = gbuf[rank * per:(rank + 1) * per].float()
    torch.testing.assert_close(mine, ref[rank * per:(rank + 1) * per] / world, atol=2e-2, rtol=2e-2)
    ss_ref = (mine ** 2).sum()
    assert abs(stats[0].item() - ss_ref.item()) <= 1e-3 * ss_ref.item() and stats[1].item() == 0
    dist.barrier()
    big.all_gather_(rank * per * 2, per * 2)
    torch.cuda.synchronize(); dist.barrier()
    full = gbuf.float()
    for r in range(world):  # every rank now holds every rank's reduced slice
        torch.testing.assert_close(full[r * per:(r + 1) * per], ref[r * per:(r + 1) * per] / world, atol=2e-2, rtol=2e-2)
    nccl_in = src.clone()
    nccl_out = torch.empty(per, device=dev, dtype=torch.bfloat16)
    t_rs = timeit(lambda: big.reduce_scatter_(rank * per * 2, per * 2, torch.bfloat16, 1.0, stats), n=20)
    t_ag = timeit(lambda: big.all_gather_(rank * per * 2, per * 2), n=20)
    t_rs_nccl = timeit(lambda: dist.reduce_scatter_tensor(nccl_out, nccl_in), n=20)
    t_ag_nccl = timeit(lambda: dist.all_gather_into_tensor(nccl_in, nccl_out), n=20)
    times["zero_512MiB"] = dict(nvls_rs_us=round(t_rs, 1), nvls_ag_us=round(t_ag, 1), nccl_rs_us=round(t_rs_nccl, 1), nccl_ag_us=round(t_ag_nccl, 1))
    return times


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nvls_symmetric_memory_and_allreduce(world):
    _need(world)
    import json

    from realhf_b200.base.testing import run_distributed
    res = run_distributed(_nvls_ar_worker, world, backend="nccl", timeout=300)
    if "skipped" in res[0]:
        pytest.skip(res[0]["skipped"])
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/nvls_allreduce_tp{world}.json", "w") as f:
        json.dump({"world": world, "per_rank": res}, f, indent=1)
    print(json.dumps(res[0]))


def _zero_worker(rank, world, comm, n_steps=3, bucket_numel=1 << 16, state="fp32"):
    """dp=world ZeRO-1 training of a small LLaMA on the GPU with the requested transport; returns the losses and a checksum
    of the final parameters (identical on every rank when the all-gather is right)."""
    import types

    import torch.distributed as dist

    from realhf_b200.api.config import ModelName
    from realhf_b200.api.data import SequenceSample
    from realhf_b200.api.model import FinetuneSpec, Model
    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.engine.engine import TrainBackend
    from realhf_b200.interfaces import basic
    from realhf_b200.models import hf_io
    from realhf_b200.models.real_model import ReaLModel
    from realhf_b200.ops import functional as OF
    from realhf_b200.ops import gemm as G
    os.environ["REAL_ZERO_COMM"] = comm
    OF.set_gemm_impl(G.linear)
    dev = torch.device("cuda", rank)
    cfg = hf_io.family("llama").make_test_config()
    cfg.n_layers, cfg.hidden_dim, cfg.n_q_heads, cfg.n_kv_heads, cfg.head_dim, cfg.intermediate_dim, cfg.vocab_size = 4, 512, 4, 4, 128, 1024, 1024
    ctx = ParallelContext.build(ProcessTopology(1, world, 1), list(range(world)), rank, backend="nccl") if world > 1 else ParallelContext.single()
    m = ReaLModel(cfg, ctx, dtype=torch.bfloat16, device=dev).instantiate(seed=7)
    tok = types.SimpleNamespace(eos_token_id=1, pad_token_id=0)
    opt = dict(lr=1e-3, weight_decay=0.01, warmup_steps_proportion=0.0, lr_scheduler_type="constant", grad_dtype="bf16",
               gradient_clipping=1.0, bucket_numel=bucket_numel, state_dtype=state, use_master_weights=(state == "fp32"))
    model = TrainBackend(optimizer=opt).initialize(Model(ModelName("m", 0), m, tok, dev), FinetuneSpec(1, 10, 10))
    g = torch.Generator().manual_seed(3)
    lens = [64] * 16
    ids = torch.randint(2, cfg.vocab_size, (sum(lens),), generator=g)
    full = SequenceSample.from_default(seqlens=lens, ids=list(range(16)),
                                       data=dict(packed_input_ids=ids, prompt_mask=torch.zeros(sum(lens), dtype=torch.bool)))
    mine = full.split(world)[rank] if world > 1 else full
    mine.to_device(dev)
    itf = basic.SFTInterface()
    losses = [float(itf.train_step(model, mine, n_mbs=1)["loss"]) for _ in range(n_steps)]
    o = model.module.optim
    flat = m.flat_param.data.float()
    chk = torch.stack([flat.sum(), flat.abs().sum(), (flat * torch.arange(flat.numel(), device=dev) % 7).sum()])
    if world > 1:
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), f"replicas diverged after the all-gather: {lo} vs {hi}"
    return dict(losses=losses, nvls=bool(o.nvls), n_buckets=len(o.buckets), n_overlapped=o.n_overlapped, chk=chk.tolist(),
                grad_norm=float(o.last_grad_norm))


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("state", ["fp32", "bf16"])
def test_zero1_nvls_matches_nccl_and_single_gpu(world, state):
    _need(world)
    from realhf_b200.base.testing import run_distributed
    from realhf_b200.parallel.symm_mem import multicast_supported
    if not multicast_supported(torch.device("cuda", 0)):
        pytest.skip("no multicast support")
    ref = run_distributed(_zero_worker, 1, backend="nccl", comm="nccl", state=state)[0]
    nccl = run_distributed(_zero_worker, world, backend="nccl", comm="nccl", state=state)
    nvls = run_distributed(_zero_worker, world, backend="nccl", comm="nvls", state=state)
    assert all(r["nvls"] for r in nvls) and not any(r["nvls"] for r in nccl)
    assert nvls[0]["n_buckets"] > 4 and nvls[0]["n_overlapped"] >= nvls[0]["n_buckets"] // 2, nvls[0]
    tol = 3e-2 if state == "fp32" else 6e-2   # bf16 parameters; the bf16-state variant rounds stochastically
    for r in nvls + nccl:
        # a rank's loss is over its own data shard: compare the mean over ranks with the single-GPU loss
        pass
    for step in range(len(ref["losses"])):
        mean_nvls = sum(r["losses"][step] for r in nvls) / world
        mean_nccl = sum(r["losses"][step] for r in nccl) / world
        assert abs(mean_nvls - ref["losses"][step]) < tol * max(1.0, abs(ref["losses"][step])), (step, mean_nvls, ref["losses"])
        assert abs(mean_nvls - mean_nccl) < tol * max(1.0, abs(mean_nccl)), (step, mean_nvls, mean_nccl)
    assert abs(nvls[0]["grad_norm"] - nccl[0]["grad_norm"]) < 0.05 * max(1.0, nccl[0]["grad_norm"])


def _tp_gen_worker(rank, world, nvls, n_new=16, force_tokens=None):
    """CUDA-graph generation of a tp=world LLaMA whose row-parallel GEMMs write partial sums into symmetric memory and whose
    all-reduce lives inside the residual-add + RMSNorm kernel (multimem.ld_reduce).  Returns tokens + log-probs; with
    `force_tokens` (single GPU) the log-probs of THOSE continuations from one packed forward pass (teacher forcing)."""
    from realhf_b200.api.model import GenerationHyperparameters, ReaLModelConfig
    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.models import generation as gen
    from realhf_b200.models.real_model import ReaLModel
    from realhf_b200.ops import functional as OF
    from realhf_b200.ops import gemm as G
    from realhf_b200.ops import launches
    os.environ["REAL_TP_NVLS"] = "1" if nvls else "0"
    OF.set_gemm_impl(G.linear)
    dev = torch.device("cuda", rank)
    cfg = ReaLModelConfig(n_layers=3, n_kv_heads=8, n_q_heads=8, hidden_dim=1024, intermediate_dim=2816, vocab_size=32000, n_positions=2048,
                          embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, activation_function="silu", scale_attn_by_inverse_layer_idx=False,
                          use_attention_bias=False, use_attn_proj_bias=False, use_mlp_bias=False, layer_norm_type="rms", mlp_type="llama",
                          apply_rotary=True)
    if world > 1:
        ctx = ParallelContext.build(ProcessTopology(1, 1, world), list(range(world)), rank, backend="nccl")
        from realhf_b200.parallel.fused_tp import FusedTP
        ctx.symm = FusedTP(ctx, max_tokens=256, max_features=cfg.hidden_dim, device=dev)
    else:
        ctx = ParallelContext.single()
    m = ReaLModel(cfg, ctx, dtype=torch.bfloat16, device=dev).instantiate(seed=5, std=0.05)
    for p in m.parameters():
        p.requires_grad_(False)
    m.eval()
    lens = [9, 33, 17, 64, 5, 12, 40, 21]
    g0 = torch.Generator().manual_seed(11)
    ids = torch.randint(3, 32000, (sum(lens),), generator=g0).to(dev)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    if force_tokens is not None:
        # teacher forcing: log p(token_t | prompt, tokens_<t) for the given continuations, from the packed training-style forward
        ft = force_tokens.to(dev)
        seqs, new_lens = [], []
        for i, L in enumerate(lens):
            seqs.append(torch.cat([ids[int(cu[i]): int(cu[i + 1])], ft[i]]))
            new_lens.append(L + ft.shape[1])
        packed = torch.cat(seqs)
        cu2 = torch.tensor([0] + list(torch.tensor(new_lens).cumsum(0)), dtype=torch.int32, device=dev)
        with torch.no_grad():
            out = m(input_ids=packed, cu_seqlens=cu2, max_seqlen=max(new_lens))
            lg = torch.log_softmax(out.logits.float(), -1)
        lps = []
        for i, L in enumerate(lens):
            rows = torch.arange(L - 1, L - 1 + ft.shape[1], device=dev) + int(cu2[i])
            lps.append(lg[rows, ft[i]])
        return dict(logprobs=torch.stack(lps).cpu())
    g = GenerationHyperparameters(max_new_tokens=n_new, min_new_tokens=n_new, greedy=True, use_cuda_graph=True, force_cudagraph_recapture=True)
    launches.reset()
    out, _ = gen.generate(m, ids, cu, g, eos_id=2, pad_id=0)
    torch.cuda.synchronize()
    fused = getattr(ctx, "symm", None)
    return dict(tokens=out.tokens.cpu(), logprobs=out.logprobs.cpu(), nvls=bool(fused is not None and fused.nvls),
                ops={k: v for k, v in launches.by_op.items() if "nvls" in k or "symm" in k})


@pytest.mark.parametrize("world", [2, 4, 8])
def test_tp_decode_with_in_switch_allreduce_matches_single_gpu(world):
    _need(world)
    from realhf_b200.base.testing import run_distributed
    from realhf_b200.parallel.symm_mem import multicast_supported
    if not multicast_supported(torch.device("cuda", 0)):
        pytest.skip("no multicast support")
    res = run_distributed(_tp_gen_worker, world, backend="nccl", nvls=True)
    assert all(r["nvls"] for r in res), "FusedTP did not get a multicast mapping"
    assert res[0]["ops"].get("nvls_ar_add_rmsnorm", 0) > 0, res[0]["ops"]
    for r in res:
        assert torch.equal(r["tokens"], res[0]["tokens"]), "TP ranks disagree on the generated tokens"
    # a random-init model has nearly flat logits, so greedy tokens fork on rounding noise: check the tensor-parallel decode by
    # teacher forcing instead -- the single-GPU model must assign (almost) the same log-probs to the TP run's own tokens
    ref = run_distributed(_tp_gen_worker, 1, backend="nccl", nvls=False, force_tokens=res[0]["tokens"])[0]
    torch.testing.assert_close(res[0]["logprobs"], ref["logprobs"], atol=0.08, rtol=0.05)


def _ep_worker(rank, world, fused, sp=True):
    """Mixtral-style MoE under expert parallelism + sequence parallelism on `world` GPUs: SFT steps with the device-driven
    peer-store dispatch / combine (csrc/ep.cu) or the NCCL all-to-all path; returns losses and a weight checksum."""
    import test_parallel_cpu as T
    os.environ["REAL_EP_FUSED"] = "1" if fused else "0"
    os.environ["REAL_EP_MAX_ROWS"] = "4096"
    from realhf_b200.ops import functional as OF
    from realhf_b200.ops import gemm as G
    OF.set_gemm_impl(G.linear)
    import types

    from realhf_b200.api.config import ModelName
    from realhf_b200.api.model import FinetuneSpec, Model
    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.engine.engine import TrainBackend
    from realhf_b200.interfaces import basic
    from realhf_b200.models import hf_io
    from realhf_b200.models.real_model import ReaLModel
    from realhf_b200.ops import launches
    cfg = hf_io.family("mixtral").make_test_config()
    cfg.hidden_dim, cfg.intermediate_dim, cfg.n_q_heads, cfg.n_kv_heads, cfg.head_dim, cfg.vocab_size, cfg.n_layers = 512, 1024, 4, 4, 128, 1024, 2
    cfg.moe.num_experts, cfg.moe.top_k = 8, 2
    cfg.moe.expert_parallel = world > 1
    cfg.moe.aux_loss_coeff = 0.0
    dev = torch.device("cuda", rank)
    ctx = ParallelContext.build(ProcessTopology(1, 1, world), list(range(world)), rank, backend="nccl", sequence_parallel=world > 1 and sp) \
        if world > 1 else ParallelContext.single()
    m = ReaLModel(cfg, ctx, dtype=torch.bfloat16, device=dev).instantiate(seed=7)
    tok = types.SimpleNamespace(eos_token_id=1, pad_token_id=0)
    model = TrainBackend(optimizer=dict(lr=1e-3, weight_decay=0.0, warmup_steps_proportion=0.0, lr_scheduler_type="constant",
                                        grad_dtype="fp32")).initialize(Model(ModelName("m", 0), m, tok, dev), FinetuneSpec(1, 10, 10))
    batch = T._batch(16, vocab=1024)
    batch.to_device(dev)
    launches.reset()
    losses = [float(basic.SFTInterface().train_step(model, batch, n_mbs=1)["loss"]) for _ in range(3)]
    return dict(losses=losses, ep_ops={k: v for k, v in launches.by_op.items() if k.startswith("ep_")})


@pytest.mark.parametrize("world", [2, 4])
def test_fused_expert_parallel_matches_all_to_all_and_single_gpu(world):
    _need(world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from realhf_b200.base.testing import run_distributed
    ref = run_distributed(_ep_worker, 1, backend="nccl", fused=False)[0]
    a2a = run_distributed(_ep_worker, world, backend="nccl", fused=False)
    fus = run_distributed(_ep_worker, world, backend="nccl", fused=True)
    assert fus[0]["ep_ops"].get("ep_plan", 0) > 0 and fus[0]["ep_ops"].get("ep_move_rows", 0) > 0, fus[0]["ep_ops"]
    assert not a2a[0]["ep_ops"]
    for r in fus + a2a:
        for x, y in zip(r["losses"], ref["losses"]):
            assert abs(x - y) < 5e-2 * max(1.0, abs(y)), (r["losses"], ref["losses"])
    for x, y in zip(fus[0]["losses"], a2a[0]["losses"]):
        assert abs(x - y) < 2e-2 * max(1.0, abs(y)), (fus[0]["losses"], a2a[0]["losses"])


def _ep_replicated_worker(rank, world, graph):
    """Expert parallelism WITHOUT sequence parallelism (the layout of generation and of inference MFCs): tokens are replicated,
    every rank runs its experts through device-side offsets (no host sync -> CUDA-graph capturable) and the partial outputs are
    all-reduced.  Returns the logits of a packed forward and a short greedy generation."""
    from realhf_b200.ops import functional as OF
    from realhf_b200.ops import gemm as G
    OF.set_gemm_impl(G.linear)
    from realhf_b200.api.model import GenerationHyperparameters
    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.models import generation as gen
    from realhf_b200.models import hf_io
    from realhf_b200.models.real_model import ReaLModel
    cfg = hf_io.family("mixtral").make_test_config()
    cfg.hidden_dim, cfg.intermediate_dim, cfg.n_q_heads, cfg.n_kv_heads, cfg.head_dim, cfg.vocab_size, cfg.n_layers = 512, 1024, 4, 4, 128, 1024, 2
    cfg.moe.num_experts, cfg.moe.top_k = 8, 2
    cfg.moe.expert_parallel = world > 1
    dev = torch.device("cuda", rank)
    ctx = ParallelContext.build(ProcessTopology(1, 1, world), list(range(world)), rank, backend="nccl") if world > 1 else ParallelContext.single()
    m = ReaLModel(cfg, ctx, dtype=torch.bfloat16, device=dev).instantiate(seed=11).eval()
    for p in m.parameters():
        p.requires_grad_(False)
    g0 = torch.Generator().manual_seed(3)
    lens = [9, 30, 17, 5]
    ids = torch.randint(3, 1024, (sum(lens),), generator=g0).to(dev)
    cu = torch.tensor([0, 9, 39, 56, 61], dtype=torch.int32, device=dev)
    with torch.no_grad():
        logits = m(input_ids=ids, cu_seqlens=cu, max_seqlen=30).logits.float().cpu()
    g = GenerationHyperparameters(max_new_tokens=8, min_new_tokens=8, greedy=True, use_cuda_graph=graph)
    out, _ = gen.generate(m, ids, cu, g, eos_id=None, pad_id=0)
    return dict(logits=logits, tokens=out.tokens.cpu(), logprobs=out.logprobs.float().cpu())


@pytest.mark.parametrize("graph", [False, True])
def test_expert_parallel_with_replicated_tokens_matches_single_gpu(graph):
    _need(2)
    from realhf_b200.base.testing import run_distributed
    ref = run_distributed(_ep_replicated_worker, 1, backend="nccl", graph=graph)[0]
    eps = run_distributed(_ep_replicated_worker, 2, backend="nccl", graph=graph)
    for r in eps:
        assert torch.isfinite(r["logits"]).all()
        torch.testing.assert_close(r["logits"], ref["logits"], atol=0.06, rtol=0.05)
        torch.testing.assert_close(r["logprobs"], eps[0]["logprobs"], atol=1e-3, rtol=1e-3)   # ranks agree with each other
        assert torch.equal(r["tokens"], eps[0]["tokens"])


def test_expert_parallel_training_without_sequence_parallel_matches_single_gpu():
    """EP on a TP group that does not shard tokens (a model whose generation and training share one tp layout): the grouped
    GEMMs and the grouped wgrad get the offsets of the rank's experts INSIDE the globally sorted rows (offsets[0] > 0 on
    every rank but the first)."""
    _need(2)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from realhf_b200.base.testing import run_distributed
    ref = run_distributed(_ep_worker, 1, backend="nccl", fused=False)[0]
    rep = run_distributed(_ep_worker, 2, backend="nccl", fused=False, sp=False)
    for r in rep:
        for x, y in zip(r["losses"], ref["losses"]):
            assert abs(x - y) < 5e-2 * max(1.0, abs(y)), (r["losses"], ref["losses"])
