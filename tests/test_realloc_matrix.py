"""Parameter reallocation executed by 8 processes over a matrix of layout pairs.

Parity: `tests/comm/test_param_realloc.py:381-557` of the reference (6 x 6 layout pairs x {gpt2, llama} x {actor, critic} on 8
GPUs, source on the first K ranks, destination on the last K' ranks, overlapping in the middle).  Here the same worker function
runs on CPU over gloo (pack / isend-irecv / unpack transport) in the CPU suite and on 8 GPUs over NCCL plus the direct
peer-store transport in the GPU suite.  Equality is bit-exact against the shards obtained by sharding the full weights directly
in the destination layout; a round trip back to the source layout must reproduce the source shards; an EMA merge (eta < 1) must
equal the convex combination."""
import itertools

import pytest
import torch

LAYOUTS = [(1, 8, 1), (1, 1, 8), (2, 2, 2), (4, 1, 2), (1, 2, 4), (2, 4, 1), (1, 4, 1), (1, 1, 2), (2, 1, 1)]  # (pp, dp, tp)
PAIRS = [(a, b) for a, b in itertools.product(LAYOUTS, LAYOUTS) if a != b]


def _diff_slots(model, flat, want=None, **tol):
    """Names of the parameters whose values in `flat` differ from `want` (default: the model's own weights).  Compared slot by
    slot: the flat layout has alignment padding between slots that no transfer touches."""
    want = model.flat_param.data if want is None else want
    bad = []
    for n, sl in model.slots.items():
        a, b = flat[sl.offset: sl.offset + sl.numel], want[sl.offset: sl.offset + sl.numel]
        if not (torch.allclose(a, b, **tol) if tol else torch.equal(a, b)):
            bad.append(n)
    return bad


def _matrix_worker(rank, world, pairs, fam, critic, direct):
    import torch.distributed as dist

    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.models import hf_io
    from realhf_b200.models.real_model import ReaLModel
    from realhf_b200.parallel import realloc
    cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", rank) if cuda else torch.device("cpu")
    dtype = torch.bfloat16 if cuda else torch.float32
    cfg = hf_io.family(fam).make_test_config()
    cfg.n_layers = 8
    cfg.is_critic = critic
    checked = 0
    for src_dims, dst_dims in pairs:
        s_topo, d_topo = ProcessTopology(*src_dims), ProcessTopology(*dst_dims)
        ns, nd = s_topo.world_size(), d_topo.world_size()
        src_workers, dst_workers = list(range(ns)), list(range(world - nd, world))

        def shard(topo, workers, seed):
            if rank not in workers:
                return None
            return ReaLModel(cfg, ParallelContext.fake(topo, workers.index(rank)), dtype=dtype, device=dev).instantiate(seed=seed)

        src, ref = shard(s_topo, src_workers, 9), shard(d_topo, dst_workers, 9)
        fwd = realloc.derive_plan(cfg, s_topo, src_workers, d_topo, dst_workers)
        bwd = realloc.derive_plan(cfg, d_topo, dst_workers, s_topo, src_workers)
        esz = torch.empty(0, dtype=dtype).element_size()
        ex_f, ex_b = realloc.ReallocExecutor(fwd, rank, esz, dev), realloc.ReallocExecutor(bwd, rank, esz, dev)
        src_flat = src.flat_param.data if src is not None else None
        dst_flat = torch.full_like(ref.flat_param.data, float("nan")) if ref is not None else None
        # ---- source layout -> destination layout
        ex_f.run(src_flat, dst_flat)
        if cuda:
            torch.cuda.synchronize()
        dist.barrier()
        if ref is not None:
            bad = _diff_slots(ref, dst_flat)
            assert not bad, (src_dims, dst_dims, bad[:8])
            checked += 1
        # ---- and back, into a fresh buffer: must reproduce the source shards
        back = torch.full_like(src_flat, float("nan")) if src is not None else None
        ex_b.run(dst_flat, back)
        if cuda:
            torch.cuda.synchronize()
        dist.barrier()
        if src is not None:
            bad = _diff_slots(src, back)
            assert not bad, (dst_dims, src_dims, bad[:8])
        # ---- EMA merge into a destination that already holds other weights
        other = shard(d_topo, dst_workers, 21)
        if other is not None:
            before = other.flat_param.data.clone()
        ex_f.run(src_flat, other.flat_param.data if other is not None else None, eta=0.25)
        if cuda:
            torch.cuda.synchronize()
        dist.barrier()
        if other is not None:
            want = (0.25 * ref.flat_param.data.float() + 0.75 * before.float()).to(dtype)
            bad = _diff_slots(other, other.flat_param.data, want, atol=1e-2 if cuda else 1e-6, rtol=1e-2 if cuda else 1e-6)
            assert not bad, ("ema", src_dims, dst_dims, bad[:8])
    return checked


def _direct_worker(rank, world, pairs, fam):
    """GPU only: every transfer is ONE segment-copy kernel storing straight into the destination GPU's flat buffer."""
    import torch.distributed as dist

    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.models import hf_io
    from realhf_b200.models.real_model import ReaLModel
    from realhf_b200.parallel import realloc
    from realhf_b200.parallel.symm_mem import SymmetricBuffer
    dev = torch.device("cuda", rank)
    cfg = hf_io.family(fam).make_test_config()
    cfg.n_layers = 8
    cap = 0
    shards = []
    for src_dims, dst_dims in pairs:
        s_topo, d_topo = ProcessTopology(*src_dims), ProcessTopology(*dst_dims)
        ns, nd = s_topo.world_size(), d_topo.world_size()
        src_workers, dst_workers = list(range(ns)), list(range(world - nd, world))
        plan = realloc.derive_plan(cfg, s_topo, src_workers, d_topo, dst_workers)
        shards.append((s_topo, d_topo, src_workers, dst_workers, plan))
        cap = max(cap, max(plan.dst_numel.values()))
    sb = SymmetricBuffer(cap * 2, device=dev)   # one symmetric destination buffer reused by every pair (collective allocation)
    for s_topo, d_topo, src_workers, dst_workers, plan in shards:
        src = ReaLModel(cfg, ParallelContext.fake(s_topo, src_workers.index(rank)), dtype=torch.bfloat16, device=dev).instantiate(seed=9) \
            if rank in src_workers else None
        ref = ReaLModel(cfg, ParallelContext.fake(d_topo, dst_workers.index(rank)), dtype=torch.bfloat16, device=dev).instantiate(seed=9) \
            if rank in dst_workers else None
        exe = realloc.ReallocExecutor(plan, rank, 2, dev)
        dst_flat = None
        if ref is not None:
            dst_flat = sb.data()[: plan.dst_numel[rank] * 2].view(torch.bfloat16)
            dst_flat.fill_(float("nan"))
        torch.cuda.synchronize(); dist.barrier()
        exe.run(src.flat_param.data if src is not None else None, dst_flat, peer_dst_ptrs={w: sb.data_ptrs[w] for w in range(world)})
        torch.cuda.synchronize(); dist.barrier()
        if ref is not None:
            bad = _diff_slots(ref, dst_flat)
            assert not bad, (s_topo.dims, d_topo.dims, bad[:8])
    return True


@pytest.mark.parametrize("fam,critic", [("llama", False), ("gpt2", False), ("llama", True)])
def test_realloc_matrix_8_processes_gloo(fam, critic):
    from realhf_b200.base.testing import run_distributed
    pairs = PAIRS[{"llama": 0, "gpt2": 1}[fam] + (2 if critic else 0)::3]   # 24 of the 72 ordered pairs per case, different ones per case
    out = run_distributed(_matrix_worker, 8, backend="gloo", timeout=600, pairs=pairs, fam=fam, critic=critic, direct=False)
    assert sum(out) >= len(pairs)


@pytest.mark.gpu
@pytest.mark.parametrize("fam,critic", [("llama", False), ("gpt2", False), ("llama", True)])
def test_realloc_matrix_8_gpus(fam, critic):
    if torch.cuda.device_count() < 8:
        pytest.skip("needs 8 GPUs")
    from realhf_b200.base.testing import run_distributed
    out = run_distributed(_matrix_worker, 8, backend="nccl", timeout=900, pairs=PAIRS[::2], fam=fam, critic=critic, direct=False)
    assert sum(out) >= len(PAIRS[::2])


@pytest.mark.gpu
def test_realloc_matrix_8_gpus_direct_peer_stores():
    if torch.cuda.device_count() < 8:
        pytest.skip("needs 8 GPUs")
    from realhf_b200.base.testing import run_distributed
    assert all(run_distributed(_direct_worker, 8, backend="nccl", timeout=900, pairs=PAIRS[::3], fam="llama"))
