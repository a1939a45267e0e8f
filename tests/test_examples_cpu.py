"""Example algorithms built on the public extension points run end to end on CPU (single process)."""
import os
import sys
import types

import pytest
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


def _script_argv(path):
    """The `python3 -m realhf_b200.apps.quickstart <algo> k=v ...` command of a launch script, shell variables expanded."""
    import re
    import shlex
    text = open(path).read()
    env = {"MODEL_PATH": "/ckpt/base", "SFT_MODEL_PATH": "/ckpt/sft", "RW_MODEL_PATH": "/ckpt/rw", "CLUSTER_SPEC_PATH": "/x.json"}
    for m in re.finditer(r"^([A-Z_]+)=([^\s$]+)\s*(?:#.*)?$", text, flags=re.M):
        env.setdefault(m.group(1), m.group(2))
    cmd = text[text.index("python3 -m realhf_b200.apps.quickstart"):].replace("\\\n", " ")
    cmd = re.sub(r"\$\{([A-Z_]+)[^}]*\}|\$([A-Z_]+)", lambda m: env.get(m.group(1) or m.group(2), ""), cmd)
    argv = shlex.split(cmd)
    return argv[3:]


@pytest.mark.parametrize("script", sorted(
    os.path.join(d, f) for d, _, fs in os.walk(os.path.join(ROOT, "examples")) for f in fs if f.endswith(".sh")))
def test_every_example_launch_script_is_a_valid_command_line(script, tmp_path, monkeypatch):
    """Launch scripts rot silently when an option is renamed: parse each one with the real override parser and, for the
    runtime experiments, resolve the allocation and build the system config the launcher would pickle."""
    monkeypatch.setenv("REAL_FILEROOT", str(tmp_path))
    from realhf_b200.apps.quickstart import build_experiment
    argv = _script_argv(script)
    exp = build_experiment(argv)
    assert exp.experiment_name and "_" not in exp.trial_name
    if hasattr(exp, "run_local"):   # profile sweep: in-process, nothing to resolve
        assert exp.handles == ["generate", "inference", "train_step"] and exp.batch_sizes == [32, 128]
        return
    # model shapes normally come from the checkpoints' config.json; the scripts point at placeholder paths
    from realhf_b200.models import hf_io
    monkeypatch.setattr(hf_io, "config_from_hf_path", lambda fam, path, is_critic=False: _llama7b(is_critic))
    sys_cfg = exp.initial_setup()
    assert len(sys_cfg.model_worker) == exp.n_nodes * exp.n_gpus_per_node
    names = {r.name for r in sys_cfg.model_rpcs}
    assert names == set(exp.rpcs)


def _llama7b(is_critic):
    from realhf_b200.api.model import ReaLModelConfig
    return ReaLModelConfig(n_layers=32, n_kv_heads=32, n_q_heads=32, hidden_dim=4096, head_dim=128, intermediate_dim=11008,
                           vocab_size=32000, n_positions=4096, is_critic=is_critic)


def _load_example(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location(f"example_{name}", os.path.join(ROOT, "examples", f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_visualize_dfg_example_covers_every_algorithm(tmp_path, monkeypatch):
    viz = _load_example("visualize_dfg")
    expect = {"sft": 1, "rw": 1, "dpo": 2, "ppo": 6, "grpo": 4, "reinforce": 5}
    for algo, n in expect.items():
        dot = str(tmp_path / f"{algo}.dot")
        monkeypatch.setattr(sys, "argv", ["visualize_dfg.py", "--algo", algo, "--dot", dot])
        G = viz.main()
        assert G.number_of_nodes() == n
        text = open(dot).read()
        assert text.startswith("digraph dfg {") and text.count("->") >= G.number_of_edges()
    assert "actor_gen" in text or "sample_gen" in text


def test_load_and_eval_rw_example_scores_sequences(tmp_path):
    import fixtures
    rw = _load_example("load_and_eval_rw")
    ckpt = str(tmp_path / "rw")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "llama", is_critic=True)
    model = rw.load_reward_model(ckpt, "llama", "cpu")
    seqs = [torch.randint(0, cfg.vocab_size, (n,)) for n in (5, 9, 1)]
    per_token, scores = rw.score(model, seqs)
    assert [v.numel() for v in per_token] == [5, 9, 1] and scores.shape == (3,)
    # packing must not leak between sequences: scoring one sequence alone gives the same values
    alone, _ = rw.score(model, [seqs[1]])
    torch.testing.assert_close(alone[0], per_token[1], atol=1e-5, rtol=1e-5)


def test_batch_not_divisible_by_the_dp_degree_of_a_balanced_mfc_is_rejected_at_launch():
    from realhf_b200.apps.quickstart import build_experiment
    exp = build_experiment(["gen", "experiment_name=g", "trial_name=t", "device=cpu", "n_gpus_per_node=4", "allocation_mode=d4m1p1",
                            "dataset.train_bs_n_seqs=6", "model.type._class=llama"])
    with pytest.raises(ValueError, match="not divisible by dp = 4"):
        exp.initial_setup()
