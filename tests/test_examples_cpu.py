"""Example algorithms built on the public extension points run end to end on CPU (single process)."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples", "new_algorithms"))
sys.path.insert(0, os.path.join(ROOT, "examples", "customized_exp"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from test_engine_cpu import make_model  # noqa: E402

from realhf_b200.api.data import SequenceSample  # noqa: E402
from realhf_b200.interfaces import basic  # noqa: E402


def _prompts(bs=4):
    plens = [5, 7, 4, 6][:bs]
    return SequenceSample.from_default(seqlens=plens, ids=list(range(bs)), data=dict(packed_prompts=torch.randint(2, 128, (sum(plens),))))


def test_grpo_iteration():
    import grpo
    torch.manual_seed(0)
    actor, ref, rew = make_model("actor"), make_model("ref", train=False), make_model("reward", critic=True, train=False, seed=9)
    itf = grpo.GRPOInterface(group_size=3, n_minibatches=2, generation_config=dict(max_new_tokens=6, min_new_tokens=2, top_k=30))
    data = itf.generate(actor, _prompts())
    assert data.bs == 4 and all(len(l) == 3 for l in data.seqlens["packed_input_ids"])
    data.update_(itf.inference(ref, data))
    data.update_(basic.PairedRewardInterface().inference(rew, data))
    assert data.data["rewards"].shape[0] == 12
    before = actor.module.module.flat_param.data.clone()
    st = itf.train_step(actor, data)
    assert not torch.equal(before, actor.module.module.flat_param.data) and st["actor_loss"] == st["actor_loss"]


def test_remax_iteration():
    import reinforce
    torch.manual_seed(0)
    actor, rew = make_model("actor"), make_model("reward", critic=True, train=False, seed=9)
    g = dict(max_new_tokens=6, min_new_tokens=2, top_k=30)
    sample, greedy = reinforce.ReinforceInterface(generation_config=g), reinforce.ReinforceInterface(generation_config=g, greedy=True)
    p = _prompts()
    data = sample.generate(actor, p)
    gd = greedy.generate(actor, _prompts())
    rw = basic.PairedRewardInterface()
    data.update_(rw.inference(rew, data))
    gr = rw.inference(rew, gd)
    gr.remap_keys_({"rewards": "greedy_rewards"})
    data.update_(gr)
    st = sample.train_step(actor, data)
    assert "baseline" in st and st["loss"] == st["loss"]
    # greedy decoding is deterministic
    gd2 = greedy.generate(actor, _prompts())
    assert gd2.flat_seqlens("packed_input_ids") != [] and gd2.data["packed_input_ids"].shape == gd2.data["packed_input_ids"].shape


def test_example_experiments_resolve():
    import grpo, ppo_external_reward, ppo_ref_ema, reinforce  # noqa: F401
    from realhf_b200.apps.quickstart import build_experiment
    c = build_experiment(["ppo-ref-ema", "experiment_name=e", "trial_name=t", "allocation_mode=d8m1p1", "ref_ema_eta=0.01"])
    s = c.initial_setup()
    pairs = s.master_worker[0].sync_param_pairs
    assert len(pairs) == 1 and pairs[0][0].role == "actor" and pairs[0][1].role == "ref"
    hooks = [h for r in s.model_rpcs if r.name == "actor_train" for h in r._post_hooks]
    assert hooks and hooks[0].eta == 0.01
