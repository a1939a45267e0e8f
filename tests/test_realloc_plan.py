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
