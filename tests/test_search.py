"""Native MCMC allocation search + simulator (host extension) through the Python driver."""
import pytest

from realhf_b200.ops import host

pytestmark = pytest.mark.skipif(host() is None, reason="host extension not built")


def _ppo(n_seqs=128):
    from realhf_b200.experiments.algos import PPOConfig
    cfg = PPOConfig(experiment_name="s", trial_name="t")
    for m in (cfg.actor, cfg.critic, cfg.ref, cfg.rew):
        m.type.size = 7
    cfg.dataset.train_bs_n_seqs = n_seqs
    return cfg


def test_simulator_respects_mesh_exclusivity_and_dependencies():
    h = host()
    prob = dict(n_gpus=4, mem_cap=100.0, link_bw=1e9, n_iters=1, role_bytes=[10.0, 10.0], meshes=[[0, 1, 2, 3], [0, 1], [2, 3]],
                edges=[(0, 1), (0, 2)],
                rpcs=[dict(name="a", role=0, kind=1, cands=[(0, 4, 1, 1, 100.0, 1.0, 1.0)]),
                      dict(name="b", role=0, kind=1, cands=[(1, 2, 1, 1, 50.0, 1.0, 1.0), (0, 4, 1, 1, 30.0, 1.0, 1.0)]),
                      dict(name="c", role=1, kind=1, cands=[(2, 2, 1, 1, 50.0, 1.0, 1.0), (0, 4, 1, 1, 30.0, 1.0, 1.0)])])
    disjoint = h.simulate_allocation(prob, [0, 0, 0])   # b and c on disjoint halves run concurrently after a
    assert disjoint["time_us"] == pytest.approx(150.0)
    shared = h.simulate_allocation(prob, [0, 1, 1])     # both on the full mesh: serialised, 100 + 30 + 30
    assert shared["time_us"] == pytest.approx(160.0)
    res = h.mcmc_search(prob, 4.0, 0.5, 1, 20000, 5)
    assert res[0]["choice"] == [0, 0, 0] and res[0]["time_us"] == pytest.approx(150.0)
    over = dict(prob, mem_cap=1.5)
    assert h.simulate_allocation(over, [0, 0, 0])["cost"] > 1e6  # memory penalty


def test_search_ppo_7b_prefers_data_parallel_on_b200():
    from realhf_b200.search.engine import search_rpc_allocations
    cfg = _ppo()
    allocs, det = search_rpc_allocations(cfg.global_device_mesh, list(cfg.rpcs.values()), cfg.models, seq_len=128, num_gen_tokens=512,
                                         time_limit_s=2.0, return_details=True)
    assert len(allocs) == 6 and det["best"]["max_mem"] <= 180e9
    by = {a.rpc.name: a for a in allocs}
    # four 7B models fit on every 180 GB GPU: nothing needs pipeline stages, training avoids tensor parallelism
    assert all(a.parallel.pipeline_parallel_size == 1 for a in allocs)
    assert by["actor_train"].parallel.model_parallel_size <= 2
    assert all(a.parallel.world_size == a.device_mesh.n_gpus for a in allocs)


def test_search_mode_resolves_into_experiment_config():
    cfg = _ppo()
    cfg.allocation_mode = "search"
    sysc = cfg.initial_setup()
    assert len(sysc.model_worker) == 8 and len(sysc.model_rpcs) == 6


# ------------------------------------------------------------------------------------------ table-driven cost model (search/cost_model.py)

_SHAPE_7B = dict(h=4096, L=32, f=11008, v=32000, n=32 * (4 * 4096 * 4096 + 3 * 4096 * 11008) + 2 * 32000 * 4096)


def test_profile_table_interpolates_and_extrapolates():
    from realhf_b200.search.cost_model import ProfileTable
    t = ProfileTable([dict(layer="block", op="fwd", bs=1, seqlen=100, time_us=10.0), dict(layer="block", op="fwd", bs=4, seqlen=100, time_us=40.0),
                      dict(layer="block", op="fwd", bs=8, seqlen=100, time_us=60.0), dict(layer="block", op="fwd", bs=1, seqlen=4000, time_us=999.0),
                      dict(layer="block", op="decode", bs=16, seqlen=384, time_us=120.0)])
    assert t.time_us("block", "fwd", 250, 128) == pytest.approx(25.0)       # between 100 and 400 tokens, nearest profiled length = 100
    assert t.time_us("block", "fwd", 600, 128) == pytest.approx(50.0)
    assert t.time_us("block", "fwd", 1600, 128) == pytest.approx(60.0 + (20.0 / 400) * 800)   # slope of the last two points
    assert t.time_us("block", "fwd", 50, 128) == pytest.approx(5.0)
    assert t.time_us("block", "fwd", 1, 128) == pytest.approx(2.5)          # floor: a quarter of the smallest measurement
    assert t.time_us("block", "fwd", 4000, 3500) == pytest.approx(999.0)    # nearest sequence length in log space
    assert t.time_us("block", "decode", 16, 512) == pytest.approx(120.0)
    assert t.time_us("head", "fwd", 10, 10) is None and not t.has("head", "fwd")


def test_shipped_7b_table_reproduces_the_measured_mfc_times():
    """Calibration check: per-MFC device times of the headline PPO run (profiles/bench_n{1,2,8}_r2_*.json) at dp = N."""
    from realhf_b200.api.config import ModelInterfaceType as T
    from realhf_b200.search.cost_model import CommModel, ProfileTable, estimate_mfc
    from realhf_b200.search.engine import HardwareModel
    hw, cm, tb = HardwareModel(), CommModel(), ProfileTable.find("llama-7")
    assert tb is not None and "NOT a layer-profiler run" in tb.meta["source"]
    measured = {1: dict(inf=875.0, train=3227.5, gen=5138.9), 2: dict(inf=455.0, train=1652.3, gen=3402.2), 8: dict(inf=118.0, train=665.9, gen=2622.9)}
    for n, ms in measured.items():
        inf = estimate_mfc(T.INFERENCE, 128, _SHAPE_7B, n, 1, 1, hw, tb, cm, 128, 512)
        tr = estimate_mfc(T.TRAIN_STEP, 128, _SHAPE_7B, n, 1, 1, hw, tb, cm, 128, 512, n_minibatches=4, optimizer_bytes_per_param=4.0)
        gen = estimate_mfc(T.GENERATE, 128, _SHAPE_7B, n, 1, 1, hw, tb, cm, 128, 512, trainable_role=True)
        assert inf.time_us / 1e3 == pytest.approx(ms["inf"], rel=0.10)
        assert tr.time_us / 1e3 == pytest.approx(ms["train"], rel=0.12)
        assert gen.time_us / 1e3 == pytest.approx(ms["gen"], rel=0.22)       # the N=8 run has 0.5 s no per-kernel model explains (ROADMAP)
    # tensor-parallel decode pays two in-graph all-reduces per layer: at 128 prompts it does not beat data-parallel decode on 8 GPUs
    dp8 = estimate_mfc(T.GENERATE, 128, _SHAPE_7B, 8, 1, 1, hw, tb, cm, 128, 512, trainable_role=True).time_us
    tp8 = estimate_mfc(T.GENERATE, 128, _SHAPE_7B, 1, 8, 1, hw, tb, cm, 128, 512, trainable_role=True).time_us
    assert tp8 > dp8
    # memory: ZeRO-1 shards the optimizer state over dp, a pipeline splits the weights
    a = estimate_mfc(T.TRAIN_STEP, 128, _SHAPE_7B, 1, 1, 1, hw, tb, cm, 128, 512, n_minibatches=4)
    b = estimate_mfc(T.TRAIN_STEP, 128, _SHAPE_7B, 8, 1, 1, hw, tb, cm, 128, 512, n_minibatches=4)
    c = estimate_mfc(T.TRAIN_STEP, 128, _SHAPE_7B, 2, 1, 4, hw, tb, cm, 128, 512, n_minibatches=4)
    assert a.mem_static == pytest.approx(16 * _SHAPE_7B["n"], rel=1e-6) and b.mem_static < 0.4 * a.mem_static and c.mem_static < b.mem_static


def test_realloc_cost_comes_from_the_planner():
    from realhf_b200.search.cost_model import CommModel, config_from_shape, realloc_time_us
    from realhf_b200.search.engine import HardwareModel
    hw, cm = HardwareModel(), CommModel()
    cfg = config_from_shape(dict(h=1024, L=4, f=2816, v=32000))
    n_bytes = 2.0 * cfg.n_params()
    g8, lo, hi = list(range(8)), [0, 1, 2, 3], [4, 5, 6, 7]
    assert realloc_time_us(cfg, (8, 1, 1), g8, (8, 1, 1), g8, hw, cm) == 0.0
    # dp8 -> dp4 x tp2 on the same GPUs: every GPU copies its own half locally (read + write at HBM speed), no link traffic
    local = realloc_time_us(cfg, (8, 1, 1), g8, (4, 2, 1), g8, hw, cm)
    assert local == pytest.approx(cm.coll_latency_us + 2 * (n_bytes / 2) / hw.hbm_bw * 1e6, rel=0.02)
    # dp8 -> dp4 on half of the same GPUs: the destination shard IS the source shard on every GPU, the runtime aliases it
    assert realloc_time_us(cfg, (8, 1, 1), g8, (4, 1, 1), hi, hw, cm) == cm.coll_latency_us
    # dp4 on GPUs 0-3 -> dp4 on GPUs 4-7: a full copy over NVLink per destination GPU
    remote = realloc_time_us(cfg, (4, 1, 1), lo, (4, 1, 1), hi, hw, cm)
    assert remote == pytest.approx(cm.coll_latency_us + n_bytes / cm.p2p_bw * 1e6, rel=0.02)
    # tp4 on GPUs 0-3 -> dp4 on GPUs 4-7: every source GPU feeds its quarter to all four destinations
    fan = realloc_time_us(cfg, (1, 4, 1), lo, (4, 1, 1), hi, hw, cm)
    assert fan == pytest.approx(remote, rel=0.05)
    # two nodes: NIC bandwidth
    far = realloc_time_us(cfg, (4, 1, 1), lo, (4, 1, 1), [8, 9, 10, 11], hw, cm)
    assert far > 10 * remote


def test_simulator_honours_planned_realloc_times_and_the_search_reranks_with_them():
    h = host()
    prob = dict(n_gpus=2, mem_cap=100.0, link_bw=1e6, n_iters=1, role_bytes=[100.0], meshes=[[0, 1], [0]], edges=[(0, 1)], realloc_latency_us=0.0,
                rpcs=[dict(name="gen", role=0, kind=0, cands=[(1, 1, 1, 1, 10.0, 1.0, 1.0)]),
                      dict(name="train", role=0, kind=2, cands=[(0, 2, 1, 1, 10.0, 1.0, 1.0)])])
    closed = h.simulate_allocation(prob, [0, 0])["time_us"]
    assert closed == pytest.approx(10.0 + 10.0 + 100.0 / 1e6 * 1e6)                       # gen pays the closed-form realloc: bytes / link_bw
    planned = h.simulate_allocation(dict(prob, realloc_table=[([0, 0, 2, 1, 1, 1, 1, 1, 1], 3.0)]), [0, 0])["time_us"]
    assert planned == pytest.approx(23.0)
    from realhf_b200.search.engine import search_rpc_allocations
    cfg = _ppo()
    allocs, det = search_rpc_allocations(cfg.global_device_mesh, list(cfg.rpcs.values()), cfg.models, seq_len=128, num_gen_tokens=512,
                                         time_limit_s=1.0, return_details=True)
    assert all("time_us_closed_form" in r for r in det["results"])
    assert det["results"] == sorted(det["results"], key=lambda r: r["cost"])
    for key, us in det["problem"].get("realloc_table", []):
        assert len(key) == 9 and us > 0


def test_layer_profiler_rows_feed_the_cost_model():
    """CPU run of the layer profiler (wall-clock timing) on a toy model: block / embedding / head / decode / optimizer rows, and a
    cost estimate computed from nothing but those rows."""
    import torch

    from realhf_b200.api.config import ModelInterfaceType as T
    from realhf_b200.models import hf_io
    from realhf_b200.search import layers
    from realhf_b200.search.cost_model import ProfileTable, estimate_mfc
    from realhf_b200.search.engine import HardwareModel
    cfg = hf_io.family("llama").make_test_config()
    kw = dict(device="cpu", dtype=torch.float32)
    rows = layers.profile_layers(cfg, [2, 4], [16], **kw) + layers.profile_head(cfg, [2, 4], [16], **kw) \
        + layers.profile_decode(cfg, [2], [16], n_tokens=4, **kw) + layers.profile_optimizer(cfg, **kw)
    kinds = {(r["layer"], r["op"]) for r in rows}
    assert {("block", "fwd"), ("block", "fwd_bwd"), ("embedding", "fwd"), ("head", "fwd"), ("head", "fwd_bwd"), ("head", "decode"),
            ("block", "decode"), ("optimizer", "step")} <= kinds
    assert all(r["time_us"] >= 0 for r in rows)
    # wall-clock rows of a loaded CI machine are noisy: the composition checks below run on the same rows with deterministic times
    for r in rows:
        n = r["bs"] if r["op"] in ("decode", "step") else r["bs"] * r["seqlen"]
        r["time_us"] = 50.0 + 3.0 * n
    tb = ProfileTable(rows)
    shape = dict(h=cfg.hidden_dim, L=cfg.n_layers, f=cfg.intermediate_dim, v=cfg.vocab_size, n=cfg.n_params())
    one = estimate_mfc(T.INFERENCE, 4, shape, 1, 1, 1, HardwareModel(), tb, None, 8, 8)
    two = estimate_mfc(T.INFERENCE, 4, shape, 2, 1, 1, HardwareModel(), tb, None, 8, 8)
    blk = tb.time_us("block", "fwd", 64, 16)
    assert blk == pytest.approx(50.0 + 3.0 * 64)
    assert one.time_us >= cfg.n_layers * blk and two.time_us < one.time_us
    gen = estimate_mfc(T.GENERATE, 4, shape, 1, 1, 1, HardwareModel(), tb, None, 8, 8)
    assert gen.breakdown["decode_step"] >= cfg.n_layers * tb.time_us("block", "decode", 4, 12) * 0.99


def test_allocation_use_cache_reuses_the_stored_search_result(tmp_path, monkeypatch):
    from realhf_b200.base import constants
    from realhf_b200.search import engine
    monkeypatch.setattr(constants, "PROFILER_CACHE_PATH", str(tmp_path / "profiler"))
    cfg = _ppo()
    cfg.allocation_mode, cfg.allocation_use_cache = "search", True
    calls = []
    real = engine.search_rpc_allocations
    monkeypatch.setattr(engine, "search_rpc_allocations", lambda *a, **k: calls.append(1) or real(*a, **dict(k, time_limit_s=0.5)))
    first = cfg._get_rpc_allocations()
    second = cfg._get_rpc_allocations()
    assert len(calls) == 1                                   # the second resolution came from the cache file
    assert [(a.rpc.name, a.device_mesh.global_ranks(), a.parallel) for a in first] == [(a.rpc.name, a.device_mesh.global_ranks(), a.parallel) for a in second]
    other = _ppo(n_seqs=64)                                   # a different problem does not hit the same entry
    other.allocation_mode, other.allocation_use_cache = "search", True
    other._get_rpc_allocations()
    assert len(calls) == 2


def test_measured_mfc_times_from_the_profile_experiment_override_the_estimate():
    """Rows of `quickstart profile` (handle, layout, global batch, sequence length -> seconds) replace the modelled time of exactly
    those candidates; the search then follows the measurement."""
    from realhf_b200.search.engine import HardwareModel, MFCProfile, build_problem, search_rpc_allocations
    cfg = _ppo()
    rpcs = list(cfg.rpcs.values())
    hw = HardwareModel()
    base, table, _ = build_problem(cfg.global_device_mesh, rpcs, cfg.models, 128, 512, 4, hw)
    # claim that tensor-parallel generation over all 8 GPUs takes 1 ms, measured
    rows = [dict(handle="generate", interface="ppo_actor", layout="d1m8p1", bs=128, seqlen=128, n_mbs=1, secs=0.001),
            dict(handle="generate", interface="ppo_actor", layout="d1m8p1", bs=128, seqlen=128, n_mbs=2, secs=0.5),     # slower n_mbs: ignored
            dict(handle="train_step", interface="sft", layout="d8m1p1", bs=128, seqlen=640, n_mbs=1, secs=0.001)]     # other interface: ignored
    prof = MFCProfile(rows)
    prob, table2, _ = build_problem(cfg.global_device_mesh, rpcs, cfg.models, 128, 512, 4, hw, mfc_profile=prof)
    gi = next(i for i, r in enumerate(rpcs) if r.name == "actor_gen")
    hit = [c for c in table2[gi] if (c[1], c[2], c[3]) == (1, 8, 1) and c[0] == 0]
    assert hit and hit[0][4] == pytest.approx(1000.0)
    ti = next(i for i, r in enumerate(rpcs) if r.name == "actor_train")
    assert sorted(c[4] for c in table2[ti]) == sorted(c[4] for c in table[ti])          # the sft row did not touch a ppo_actor MFC
    allocs = search_rpc_allocations(cfg.global_device_mesh, rpcs, cfg.models, seq_len=128, num_gen_tokens=512, time_limit_s=1.0, mfc_profile=prof)
    gen = next(a for a in allocs if a.rpc.name == "actor_gen")
    assert (gen.parallel.data_parallel_size, gen.parallel.model_parallel_size, gen.parallel.pipeline_parallel_size) == (1, 8, 1)


def test_simulator_does_not_charge_memory_for_an_aliased_replica():
    """A tp = pp = 1 replica on GPUs of the training mesh aliases the training weights (system/model_worker.py::_param_realloc): the
    memory model must not count a second copy, or the search would reject allocations that fit."""
    h = host()
    base = dict(n_gpus=4, mem_cap=100.0, link_bw=1e9, n_iters=1, role_bytes=[60.0], meshes=[[0, 1, 2, 3], [2, 3], [0, 1]], edges=[(0, 1)])
    train = dict(name="train", role=0, kind=2, cands=[(0, 4, 1, 1, 10.0, 60.0, 10.0)])
    sub = dict(name="gen", role=0, kind=0, cands=[(1, 2, 1, 1, 10.0, 0.0, 10.0), (1, 1, 2, 1, 10.0, 0.0, 10.0)])
    prob = dict(base, rpcs=[sub, train])
    aliased = h.simulate_allocation(prob, [0, 0])      # dp2 on GPUs 2-3 of a dp4 training mesh: same shard, no second copy
    resharded = h.simulate_allocation(prob, [1, 0])    # tp2: a real replica of half the weights
    assert aliased["max_mem"] == pytest.approx(60.0 + 10.0)
    assert resharded["max_mem"] == pytest.approx(60.0 + 10.0 + 30.0)
