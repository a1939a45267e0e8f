"""Parameter reallocation plans: for many (pp,dp,tp) layout pairs the planned segments reproduce exactly the
destination shards obtained by sharding the full weights directly (single process, no process groups)."""
import itertools

import pytest
import torch

from realhf_b200.base.topology import ParallelContext, ProcessTopology
from realhf_b200.models import hf_io, sharding
from realhf_b200.models.real_model import ReaLModel
from realhf_b200.parallel import realloc

LAYOUTS = [(1, 1, 1), (1, 2, 1), (1, 1, 2), (2, 1, 2), (1, 2, 4), (4, 1, 2), (2, 2, 2), (1, 1, 8)]  # (pp, dp, tp)


def build_shards(cfg, layout, seed=3):
    pp, dp, tp = layout
    topo = ProcessTopology(pp, dp, tp)
    out = {}
    for r in range(topo.world_size()):
        ctx = ParallelContext.fake(topo, r)
        out[r] = ReaLModel(cfg, ctx, dtype=torch.float32).instantiate(seed=seed)
    return topo, out


@pytest.mark.parametrize("fam", ["llama", "gpt2"])
@pytest.mark.parametrize("pair", [(a, b) for a, b in itertools.product(LAYOUTS, LAYOUTS) if a != b][::3])
def test_plan_reproduces_destination_shards(fam, pair):
    cfg = hf_io.family(fam).make_test_config()
    cfg.n_layers = 8
    src_layout, dst_layout = pair
    s_topo, src = build_shards(cfg, src_layout, seed=3)
    d_topo, dst_ref = build_shards(cfg, dst_layout, seed=3)
    ns, nd = s_topo.world_size(), d_topo.world_size()
    # source on workers [0, ns), destination on the LAST nd workers of an 8-GPU box: they overlap in the middle
    src_workers = list(range(ns))
    dst_workers = list(range(8 - nd, 8))
    plan = realloc.derive_plan(cfg, s_topo, src_workers, d_topo, dst_workers)
    dst_flat = {w: torch.full((plan.dst_numel[w],), float("nan")) for w in dst_workers}
    for t in plan.transfers:
        s = src[src_workers.index(t.src_worker)].flat_param.data
        d = dst_flat[t.dst_worker]
        for so, do, ln in zip(t.src_off, t.dst_off, t.lens):
            d[do:do + ln] = s[so:so + ln]
    for r in range(nd):
        ref = dst_ref[r]
        got = dst_flat[dst_workers[r]]
        for name, slot in ref.slots.items():
            torch.testing.assert_close(got[slot.offset:slot.offset + slot.numel].view(slot.shape), ref.p[name].data, rtol=0, atol=0)


def test_executor_local_and_ema():
    cfg = hf_io.family("llama").make_test_config()
    topo = ProcessTopology(1, 1, 1)
    plan = realloc.derive_plan(cfg, topo, [0], topo, [0])
    src = ReaLModel(cfg, dtype=torch.float32).instantiate(seed=1)
    dst = ReaLModel(cfg, dtype=torch.float32).instantiate(seed=2)
    ex = realloc.ReallocExecutor(plan, 0, 4, "cpu")
    before = dst.flat_param.data.clone()
    ex.run(src.flat_param.data, dst.flat_param.data, eta=0.25)
    for name, slot in dst.slots.items():
        a, b = slot.offset, slot.offset + slot.numel
        torch.testing.assert_close(dst.flat_param.data[a:b], 0.25 * src.flat_param.data[a:b] + 0.75 * before[a:b])
    ex.run(src.flat_param.data, dst.flat_param.data)
    for name in dst.p:
        assert torch.equal(dst.p[name].data, src.p[name].data)


@pytest.mark.parametrize("pair", [((1, 2, 4), (2, 1, 4)), ((1, 1, 8), (1, 1, 8)), ((1, 1, 8), (1, 2, 4)), ((1, 2, 4), (1, 1, 8)),
                                  ((1, 4, 2), (1, 1, 8)), ((2, 1, 4), (1, 8, 1))])
def test_replicated_kv_heads_are_written_exactly_once(pair):
    """n_kv_heads < tp: several source ranks hold the same KV head.  Every destination element must be written exactly once
    (an EMA merge applied twice would be wrong), also when source and destination share the TP degree."""
    cfg = hf_io.family("llama").make_test_config()
    cfg.n_kv_heads, cfg.n_q_heads = 2, 8
    (spp, sdp, stp), (dpp, ddp, dtp) = pair
    s_topo, d_topo = ProcessTopology(spp, sdp, stp), ProcessTopology(dpp, ddp, dtp)
    plan = realloc.derive_plan(cfg, s_topo, list(range(8)), d_topo, list(range(8)))
    hits = {w: torch.zeros(n, dtype=torch.int32) for w, n in plan.dst_numel.items()}
    for t in plan.transfers:
        for do, ln in zip(t.dst_off, t.lens):
            hits[t.dst_worker][do:do + ln] += 1
    for w, h in hits.items():
        assert int(h.min()) == 1 and int(h.max()) == 1, (w, int(h.min()), int(h.max()))


def test_plan_derivation_of_a_7b_model_is_fast():
    """LLaMA-7B, dp8 -> dp4*tp2: 2.1 M row segments (column-split weights give one segment per row).  Each worker derives the
    plan at start-up; a quadratic interval routine here once cost ~10 minutes."""
    import time

    from realhf_b200.api.model import ReaLModelConfig
    cfg = ReaLModelConfig(n_layers=32, n_kv_heads=32, n_q_heads=32, hidden_dim=4096, intermediate_dim=11008, vocab_size=32000,
                          n_positions=4096, layer_norm_type="rms", mlp_type="llama", apply_rotary=True, use_attention_bias=False,
                          use_attn_proj_bias=False, use_mlp_bias=False)
    t0 = time.perf_counter()
    plan = realloc.derive_plan(cfg, ProcessTopology(1, 8, 1), list(range(8)), ProcessTopology(1, 4, 2), list(range(8)))
    dt = time.perf_counter() - t0
    assert dt < 20.0, dt
    assert len(plan.transfers) == 8 and all(t.src_worker == t.dst_worker for t in plan.transfers)   # dp-replicated source: local copies
    assert sum(t.numel for t in plan.transfers) == 8 * plan.dst_numel[0]


def test_per_worker_and_volume_only_plans_agree_with_the_full_plan():
    cfg = hf_io.family("llama").make_test_config()
    cfg.n_layers = 8
    s_topo, d_topo = ProcessTopology(2, 1, 2), ProcessTopology(1, 2, 4)
    sw, dw = [0, 1, 2, 3], list(range(8))
    full = realloc.derive_plan(cfg, s_topo, sw, d_topo, dw)
    vols = realloc.derive_plan(cfg, s_topo, sw, d_topo, dw, volumes_only=True)
    assert vols == {(t.src_worker, t.dst_worker): t.numel for t in full.transfers}
    for w in range(8):
        mine = realloc.derive_plan(cfg, s_topo, sw, d_topo, dw, for_worker=w)
        assert mine.dst_numel == full.dst_numel
        want = [t for t in full.transfers if w in (t.src_worker, t.dst_worker)]
        assert [(t.src_worker, t.dst_worker) for t in mine.transfers] == [(t.src_worker, t.dst_worker) for t in want]
        for a, b in zip(mine.transfers, want):
            assert list(a.src_off) == list(b.src_off) and list(a.dst_off) == list(b.dst_off) and list(a.lens) == list(b.lens)


@pytest.mark.parametrize("fam", ["mixtral", "qwen2", "gemma"])
@pytest.mark.parametrize("pair", [((1, 1, 1), (1, 1, 2)), ((1, 2, 2), (2, 1, 1)), ((2, 1, 2), (1, 2, 4)), ((1, 1, 4), (1, 2, 2)), ((1, 1, 2), (4, 1, 1))])
def test_plans_of_moe_gqa_and_tied_embedding_families(fam, pair):
    """Expert tensors ([E, 2F, H] / [E, H, F], F-sharded over TP), grouped-query attention with biases (qwen2) and tied embeddings
    (gemma): the planned segments reproduce the directly sharded destination and write every element exactly once."""
    cfg = hf_io.family(fam).make_test_config()
    src_layout, dst_layout = pair
    s_topo, src = build_shards(cfg, src_layout, seed=5)
    d_topo, dst_ref = build_shards(cfg, dst_layout, seed=5)
    ns, nd = s_topo.world_size(), d_topo.world_size()
    src_workers, dst_workers = list(range(ns)), list(range(8 - nd, 8))
    plan = realloc.derive_plan(cfg, s_topo, src_workers, d_topo, dst_workers)
    dst_flat = {w: torch.full((plan.dst_numel[w],), float("nan")) for w in dst_workers}
    hits = {w: torch.zeros(plan.dst_numel[w], dtype=torch.int32) for w in dst_workers}
    for t in plan.transfers:
        s = src[src_workers.index(t.src_worker)].flat_param.data
        for so, do, ln in zip(t.src_off, t.dst_off, t.lens):
            dst_flat[t.dst_worker][do:do + ln] = s[so:so + ln]
            hits[t.dst_worker][do:do + ln] += 1
    for r in range(nd):
        ref, got = dst_ref[r], dst_flat[dst_workers[r]]
        for name, slot in ref.slots.items():
            torch.testing.assert_close(got[slot.offset:slot.offset + slot.numel].view(slot.shape), ref.p[name].data, rtol=0, atol=0)
            assert int(hits[dst_workers[r]][slot.offset:slot.offset + slot.numel].max()) == 1, name


def test_interval_sweep_and_coalescing_against_brute_force():
    """`_uncovered` (linear sweep of sorted segments against a sorted cover) and `_coalesce_arrays` on random inputs, compared
    element by element with a set-based reference."""
    import random

    import numpy as np
    rng = random.Random(0)
    for _ in range(1500):
        cov, cov_set = [], set()
        for _round in range(3):
            segs, pos = [], 0
            for _k in range(rng.randint(0, 6)):
                pos += rng.randint(0, 5)
                ln = rng.randint(1, 6)
                segs.append((rng.randint(0, 100), pos, ln))
                pos += ln
            got = realloc._uncovered(list(segs), cov)
            want = []
            for so, do, ln in sorted(segs, key=lambda t: t[1]):
                for x in range(do, do + ln):
                    if x not in cov_set:
                        cov_set.add(x)
                        want.append((x, so + x - do))
            assert sorted((d + i, s + i) for s, d, l in got for i in range(l)) == sorted(want)
            assert all(cov[i][1] < cov[i + 1][0] for i in range(len(cov) - 1))           # sorted, disjoint, maximal
            assert {x for a, b in cov for x in range(a, b)} == cov_set
    for _ in range(300):
        n = rng.randint(1, 12)
        so, do, ln, s, d = [], [], [], 0, 0
        for _k in range(n):
            if rng.random() < 0.5:        # a gap on one side breaks adjacency
                s += rng.randint(1, 3)
            if rng.random() < 0.3:
                d += rng.randint(1, 3)
            l = rng.randint(1, 4)
            so.append(s); do.append(d); ln.append(l)
            s += l; d += l
        a, b, c = realloc._coalesce_arrays(np.array(so), np.array(do), np.array(ln))
        assert sorted((x + i, y + i) for x, y, l in zip(a, b, c) for i in range(l)) == sorted((x + i, y + i) for x, y, l in zip(so, do, ln) for i in range(l))
        assert all(not (a[i] + c[i] == a[i + 1] and b[i] + c[i] == b[i + 1]) for i in range(len(c) - 1))   # nothing left to merge


def test_whole_local_copy_is_detected_exactly_when_the_shards_coincide():
    cfg = hf_io.family("llama").make_test_config()

    def ex(src, sw, dst, dw, me):
        plan = realloc.derive_plan(cfg, ProcessTopology(*src), sw, ProcessTopology(*dst), dw, for_worker=me)
        return realloc.ReallocExecutor(plan, me, 4, "cpu")

    assert ex((1, 2, 1), [0, 1], (1, 1, 1), [1], 1).whole_local_copy()            # dp2 -> dp1 on one of its GPUs: same shard
    assert not ex((1, 2, 1), [0, 1], (1, 1, 1), [1], 0).whole_local_copy()        # worker 0 is not a destination at all
    assert not ex((1, 1, 1), [0], (1, 1, 1), [1], 1).whole_local_copy()           # the weights come from another GPU
    assert not ex((1, 2, 1), [0, 1], (1, 1, 2), [0, 1], 0).whole_local_copy()     # tp2 destination: half of every split tensor
    assert ex((1, 2, 2), [0, 1, 2, 3], (1, 1, 2), [2, 3], 2).whole_local_copy()   # tp2 on both sides, same tp position
    assert not ex((2, 1, 1), [0, 1], (1, 1, 1), [0], 0).whole_local_copy()        # pipeline stages merge: stage 1 arrives from a peer
