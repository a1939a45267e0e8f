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
