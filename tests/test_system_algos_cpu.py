"""The remaining quickstart experiments through the WHOLE runtime on CPU (launcher -> master + model workers over ZMQ / gloo):
reward modelling with evaluation, DPO (ref inference -> actor training), generation-only."""
import json
import os
import sys
import uuid

import pytest

pytestmark = pytest.mark.distributed
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fixtures  # noqa: E402
from test_system_cpu import _env  # noqa: E402


def _master_log(exp):
    return open(os.path.join(os.environ["REAL_FILEROOT"], "logs", exp.experiment_name, "t0", "master_worker-0")).read()


def test_rw_experiment_with_eval(tmp_path):
    _env(tmp_path)
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    ckpt = str(tmp_path / "llama")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "llama")
    train, valid = str(tmp_path / "pairs.jsonl"), str(tmp_path / "valid.jsonl")
    fixtures.write_pair_dataset(train, words, n=32)
    fixtures.write_pair_dataset(valid, words, n=8, seed=3)
    exp = build_experiment(["rw", f"experiment_name=rw-{uuid.uuid4().hex[:6]}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_gpus_per_node=2",
                            "allocation_mode=manual", "allocation.parallel.data_parallel_size=2", "model.type._class=llama", f"model.path={ckpt}",
                            f"dataset.train_path={train}", f"dataset.valid_path={valid}", "dataset.train_bs_n_seqs=8", "dataset.max_seqlen=64",
                            "exp_ctrl.total_train_epochs=1", "exp_ctrl.eval_freq_steps=2", "model.optimizer.grad_dtype=fp32",
                            "model.gradient_checkpointing=false"])
    main_start(exp, timeout=600)
    log = _master_log(exp)
    assert log.count("[trainDefault]") >= 2 and "eval " in log, log[-2500:]


def test_dpo_experiment(tmp_path):
    _env(tmp_path)
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    ckpt = str(tmp_path / "llama")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "llama")
    train = str(tmp_path / "pairs.jsonl")
    fixtures.write_pair_dataset(train, words, n=32)
    args = ["dpo", f"experiment_name=dpo-{uuid.uuid4().hex[:6]}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_gpus_per_node=2",
            "allocation_mode=manual", f"dataset.train_path={train}", "dataset.train_bs_n_seqs=8", "dataset.max_seqlen=64",
            "exp_ctrl.total_train_epochs=1", "exp_ctrl.benchmark_steps=3", "actor_train.parallel.data_parallel_size=2",
            "ref_inf.parallel.model_parallel_size=2"]
    for role in ("actor", "ref"):
        args += [f"{role}.type._class=llama", f"{role}.path={ckpt}", f"{role}.optimizer.grad_dtype=fp32", f"{role}.gradient_checkpointing=false"]
    exp = build_experiment(args)
    main_start(exp, timeout=600)
    log = _master_log(exp)
    assert log.count("[actor_train]") == 3 and "benchmark finished" in log, log[-2500:]


def test_generation_experiment_writes_jsonl(tmp_path):
    _env(tmp_path)
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    ckpt = str(tmp_path / "llama")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "llama")
    data = str(tmp_path / "prompts.jsonl")
    fixtures.write_prompt_dataset(data, words, n=16)
    out = str(tmp_path / "gen.jsonl")
    exp = build_experiment(["gen", f"experiment_name=gen-{uuid.uuid4().hex[:6]}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_gpus_per_node=2",
                            "allocation_mode=manual", "allocation.parallel.data_parallel_size=2", "model.type._class=llama", f"model.path={ckpt}",
                            "model.backend=inference", f"dataset.path={data}", "dataset.train_bs_n_seqs=8", "dataset.max_prompt_len=16",
                            "gen.max_new_tokens=5", "gen.min_new_tokens=2", f"output_file={out}", "exp_ctrl.total_train_epochs=1"])
    main_start(exp, timeout=600)
    rows = [json.loads(l) for l in open(out)]
    assert len(rows) == 16 and all("answer" in r or "generated" in r or len(r) >= 2 for r in rows), rows[:2]


@pytest.mark.parametrize("case", [("search", "ppo", 2), ("d2m2p1", "sft", 4)])
def test_allocation_modes_through_runtime(tmp_path, case):
    """`allocation_mode=search` (C++ MCMC search) for PPO and a regex layout with TP x DP on 4 workers for SFT."""
    mode, algo, n = case
    _env(tmp_path)
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    ckpt, crit = str(tmp_path / "llama"), str(tmp_path / "critic")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "llama")
    fixtures.make_checkpoint(crit, "llama", is_critic=True, seed=5)
    name = f"m-{uuid.uuid4().hex[:6]}"
    if algo == "sft":
        data = str(tmp_path / "sft.jsonl")
        fixtures.write_sft_dataset(data, words, n=32)
        args = ["sft", f"experiment_name={name}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_nodes=1", f"n_gpus_per_node={n}",
                f"allocation_mode={mode}", "model.type._class=llama", f"model.path={ckpt}", f"dataset.train_path={data}",
                "dataset.train_bs_n_seqs=8", "dataset.max_seqlen=64", "exp_ctrl.total_train_epochs=1", "exp_ctrl.benchmark_steps=2",
                "model.optimizer.grad_dtype=fp32", "model.gradient_checkpointing=false"]
    else:
        data = str(tmp_path / "prompts.jsonl")
        fixtures.write_prompt_dataset(data, words, n=32)
        args = ["ppo", f"experiment_name={name}", "trial_name=t0", "device=cpu", "dtype=fp32", f"n_gpus_per_node={n}", f"allocation_mode={mode}",
                f"dataset.path={data}", "dataset.train_bs_n_seqs=8", "dataset.max_prompt_len=16", "ppo.gen.max_new_tokens=6",
                "ppo.gen.min_new_tokens=2", "ppo.gen.top_k=20", "ppo.ppo_n_minibatches=2", "exp_ctrl.total_train_epochs=1",
                "exp_ctrl.benchmark_steps=2"]
        for role, path in (("actor", ckpt), ("ref", ckpt), ("critic", crit), ("rew", crit)):
            args += [f"{role}.type._class=llama", f"{role}.path={path}", f"{role}.optimizer.grad_dtype=fp32", f"{role}.gradient_checkpointing=false"]
    exp = build_experiment(args)
    main_start(exp, timeout=600)
    assert "benchmark finished" in _master_log(exp)


@pytest.mark.parametrize("opt", ["model.zero_stage=3", "model.offload=true"])
def test_sharded_optimizer_variants_save_hf_checkpoint(tmp_path, opt):
    """ZeRO-3 (parameters sharded between calls) and host-offloaded optimizer state through the runtime, incl. a periodic
    HuggingFace checkpoint that `transformers` can load."""
    _env(tmp_path)
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    ckpt = str(tmp_path / "llama")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "llama")
    data = str(tmp_path / "sft.jsonl")
    fixtures.write_sft_dataset(data, words, n=32)
    exp = build_experiment(["sft", f"experiment_name=o-{uuid.uuid4().hex[:6]}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_nodes=1",
                            "n_gpus_per_node=2", "allocation_mode=d2m1p1", "model.type._class=llama", f"model.path={ckpt}", f"dataset.train_path={data}",
                            "dataset.train_bs_n_seqs=8", "dataset.max_seqlen=64", "exp_ctrl.total_train_epochs=1", "exp_ctrl.benchmark_steps=3",
                            "exp_ctrl.save_freq_steps=2", "model.optimizer.grad_dtype=fp32", "model.gradient_checkpointing=false", opt])
    main_start(exp, timeout=600)
    found = [os.path.join(d, f) for d, _, fs in os.walk(os.path.join(os.environ["REAL_FILEROOT"], "checkpoints")) for f in fs if f == "config.json"]
    assert found, "no checkpoint written"
    import transformers
    transformers.AutoModelForCausalLM.from_pretrained(os.path.dirname(found[0]))


@pytest.mark.parametrize("script,exp_name,roles", [
    ("new_algorithms/grpo.py", "grpo", ("actor", "ref", "rew")),
    ("new_algorithms/reinforce.py", "reinforce", ("actor", "rew")),
    ("customized_exp/ppo_ref_ema.py", "ppo-ref-ema", ("actor", "ref", "critic", "rew")),
    ("customized_exp/ppo_external_reward.py", "ppo-external-reward", ("actor", "ref", "critic", "rew")),
])
def test_custom_experiment_script_through_runtime(tmp_path, script, exp_name, roles):
    """`python examples/<script> <experiment> ...`: the user's file registers an interface and / or an experiment, the launcher
    starts workers that re-import it (REAL_USER_CODE) and the run completes -- group sampling (GRPO), key remaps (ReMax),
    an EMA parameter-reallocation hook into the reference model, a Python reward function instead of a reward model."""
    import subprocess
    _env(tmp_path)
    ckpt, crit = str(tmp_path / "llama"), str(tmp_path / "critic")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "llama")
    fixtures.make_checkpoint(crit, "llama", is_critic=True, seed=5)
    data = str(tmp_path / "prompts.jsonl")
    fixtures.write_prompt_dataset(data, words, n=32)
    name = f"ex-{uuid.uuid4().hex[:6]}"
    args = [sys.executable, os.path.join(ROOT, "examples", script), exp_name, f"experiment_name={name}", "trial_name=t0",
            "device=cpu", "dtype=fp32", "n_gpus_per_node=2", "allocation_mode=heuristic", f"dataset.path={data}", "dataset.train_bs_n_seqs=8",
            "dataset.max_prompt_len=16", "exp_ctrl.total_train_epochs=1", "exp_ctrl.benchmark_steps=2"]
    for role in roles:
        args += [f"{role}.type._class=llama", f"{role}.path={crit if role in ('critic', 'rew') else ckpt}"]
        if role in ("actor", "critic"):
            args += [f"{role}.optimizer.grad_dtype=fp32", f"{role}.gradient_checkpointing=false"]
    if exp_name.startswith("ppo"):
        args += ["ppo.gen.max_new_tokens=6", "ppo.gen.min_new_tokens=2", "ppo.gen.top_k=20", "ppo.ppo_n_minibatches=2"]
    if exp_name == "ppo-ref-ema":
        args += ["ref_ema_eta=0.5"]
    r = subprocess.run(args, cwd=ROOT, env=dict(os.environ), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    log = open(os.path.join(os.environ["REAL_FILEROOT"], "logs", name, "t0", "master_worker-0")).read()
    assert "benchmark finished" in log, log[-2000:]


def test_ppo_pipeline_generation_and_mixed_layouts(tmp_path):
    """PPO where generation and actor training are pipelined (pp2: more micro-batches than requested), critic training is
    tp2, critic inference pp2, reference dp2, reward tp2: realloc + data transfer between all of them, and the periodic
    actor checkpoint (saved from a pipeline layout) loads in `transformers`."""
    _env(tmp_path)
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    ckpt, crit = str(tmp_path / "llama"), str(tmp_path / "critic")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "llama")
    fixtures.make_checkpoint(crit, "llama", is_critic=True, seed=5)
    data = str(tmp_path / "prompts.jsonl")
    fixtures.write_prompt_dataset(data, words, n=32)
    args = ["ppo", f"experiment_name=pp-{uuid.uuid4().hex[:6]}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_gpus_per_node=2",
            "allocation_mode=manual", f"dataset.path={data}", "dataset.train_bs_n_seqs=8", "dataset.max_prompt_len=16",
            "ppo.gen.max_new_tokens=6", "ppo.gen.min_new_tokens=2", "ppo.gen.top_k=20", "ppo.ppo_n_minibatches=2",
            "exp_ctrl.total_train_epochs=1", "exp_ctrl.benchmark_steps=2", "exp_ctrl.save_freq_steps=1",
            "actor_gen.parallel.pipeline_parallel_size=2", "actor_train.parallel.pipeline_parallel_size=2",
            "critic_train.parallel.model_parallel_size=2", "critic_inf.parallel.pipeline_parallel_size=2",
            "ref_inf.parallel.data_parallel_size=2", "rew_inf.parallel.model_parallel_size=2"]
    for role, path in (("actor", ckpt), ("ref", ckpt), ("critic", crit), ("rew", crit)):
        args += [f"{role}.type._class=llama", f"{role}.path={path}", f"{role}.optimizer.grad_dtype=fp32", f"{role}.gradient_checkpointing=false"]
    exp = build_experiment(args)
    main_start(exp, timeout=600)
    assert "benchmark finished" in _master_log(exp)
    found = [os.path.join(d, f) for d, _, fs in os.walk(os.path.join(os.environ["REAL_FILEROOT"], "checkpoints")) for f in fs
             if f == "config.json" and "/actor/" in d + "/"]
    assert found
    import transformers
    transformers.AutoModelForCausalLM.from_pretrained(os.path.dirname(found[0]))


@pytest.mark.parametrize("mode", ["heuristic", "pipe_model"])
def test_two_node_cluster_shape_through_runtime(tmp_path, mode):
    """n_nodes=2 x n_gpus_per_node=2 (four workers on this host standing in for two nodes): per-MFC device meshes that are node
    halves / whole nodes, pipeline stages across the node boundary, and data / parameter movement between them."""
    _env(tmp_path)
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    ckpt, crit = str(tmp_path / "llama"), str(tmp_path / "critic")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "llama")
    fixtures.make_checkpoint(crit, "llama", is_critic=True, seed=5)
    data = str(tmp_path / "prompts.jsonl")
    fixtures.write_prompt_dataset(data, words, n=32)
    name = f"mn-{uuid.uuid4().hex[:6]}"
    args = ["ppo", f"experiment_name={name}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_nodes=2", "n_gpus_per_node=2",
            f"allocation_mode={mode}", f"dataset.path={data}", "dataset.train_bs_n_seqs=8", "dataset.max_prompt_len=16",
            "ppo.gen.max_new_tokens=6", "ppo.gen.min_new_tokens=2", "ppo.gen.top_k=20", "ppo.ppo_n_minibatches=2",
            "exp_ctrl.total_train_epochs=1", "exp_ctrl.benchmark_steps=2"]
    if mode == "pipe_model":
        args.append("ppo.gen.greedy=True")   # greedy decoding emits no keep-mask: the graph must not wait for one (it used to hang)
    for role, path in (("actor", ckpt), ("ref", ckpt), ("critic", crit), ("rew", crit)):
        args += [f"{role}.type._class=llama", f"{role}.path={path}", f"{role}.optimizer.grad_dtype=fp32", f"{role}.gradient_checkpointing=false"]
    exp = build_experiment(args)
    sys_cfg = exp.initial_setup()
    assert len(sys_cfg.model_worker) == 4
    assert any("packed_logits_mask" in r.output_keys for r in sys_cfg.model_rpcs) == (mode != "pipe_model")
    main_start(exp, timeout=900)
    assert "benchmark finished" in _master_log(exp)


def test_ppo_zero3_training_layout_with_reallocated_generation_layout(tmp_path):
    """ZeRO-3 keeps 1/dp of the trainable weights between calls; generation / inference run on OTHER layouts of the same
    models, so every step gathers the shards for the parameter reallocation and releases them again (this used to crash)."""
    _env(tmp_path)
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    ckpt, crit = str(tmp_path / "llama"), str(tmp_path / "critic")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "llama")
    fixtures.make_checkpoint(crit, "llama", is_critic=True, seed=5)
    data = str(tmp_path / "prompts.jsonl")
    fixtures.write_prompt_dataset(data, words, n=16)
    name = f"z3-{uuid.uuid4().hex[:6]}"
    args = ["ppo", f"experiment_name={name}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_gpus_per_node=2", "allocation_mode=manual",
            "actor_gen.parallel.model_parallel_size=2", "actor_train.parallel.data_parallel_size=2",
            "critic_train.parallel.data_parallel_size=2", "critic_inf.parallel.model_parallel_size=2",
            "rew_inf.parallel.data_parallel_size=2", "ref_inf.parallel.data_parallel_size=2", "actor.zero_stage=3", "critic.zero_stage=3",
            f"dataset.path={data}", "dataset.train_bs_n_seqs=8", "dataset.max_prompt_len=16", "ppo.gen.max_new_tokens=6",
            "ppo.gen.min_new_tokens=2", "ppo.gen.top_k=20", "ppo.ppo_n_minibatches=2", "exp_ctrl.total_train_epochs=1"]
    for role, path in (("actor", ckpt), ("ref", ckpt), ("critic", crit), ("rew", crit)):
        args += [f"{role}.type._class=llama", f"{role}.path={path}", f"{role}.optimizer.grad_dtype=fp32", f"{role}.gradient_checkpointing=false"]
    exp = build_experiment(args)
    main_start(exp, timeout=600)
    log = _master_log(exp)
    assert log.count("[actor_train] step") == 2


def test_sft_evaluation_under_pipeline_parallelism_reports_the_last_stage(tmp_path):
    """pp=2 + periodic evaluation: only the last stage has the loss; the first stage used to crash on the missing statistics
    and the master used to log the first worker's (empty) reply."""
    import json
    _env(tmp_path)
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    ckpt = str(tmp_path / "llama")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "llama")
    data = str(tmp_path / "sft.jsonl")
    fixtures.write_sft_dataset(data, words, n=16)
    name = f"ppev-{uuid.uuid4().hex[:6]}"
    exp = build_experiment(["sft", f"experiment_name={name}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_gpus_per_node=2",
                            "allocation_mode=manual", "allocation.parallel.pipeline_parallel_size=2", "model.type._class=llama",
                            f"model.path={ckpt}", f"dataset.train_path={data}", f"dataset.valid_path={data}", "dataset.train_bs_n_seqs=8",
                            "dataset.valid_bs_n_seqs=8", "dataset.max_seqlen=64", "exp_ctrl.total_train_epochs=1", "exp_ctrl.eval_freq_steps=1",
                            "model.optimizer.grad_dtype=fp32", "model.gradient_checkpointing=false"])
    main_start(exp, timeout=600)
    stats = [json.loads(l) for l in open(os.path.join(os.environ["REAL_FILEROOT"], "logs", name, "t0", "stats.jsonl"))]
    evals = [r for r in stats if r["rpc"] == "eval/default"]
    train = [r for r in stats if r["rpc"] == "trainDefault"]
    assert len(evals) == 2 and all(3.0 < r["loss"] < 7.0 and r["ppl"] > 20 for r in evals), evals   # ~ln(vocab), not 0.0
    assert abs(evals[0]["loss"] - train[0]["loss"]) < 0.5


def test_start_subcommand_launches_a_registered_experiment_from_user_code(tmp_path):
    """`python -m realhf_b200.apps.main start -e NAME -f TRIAL --user_code FILE ...` (parity: apps/main.py `start`): the experiment is
    built from the registry, the launcher flags configure the run, the workers import the same user code."""
    _env(tmp_path)
    from realhf_b200.apps import main as M
    ckpt = str(tmp_path / "llama")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "llama")
    data = str(tmp_path / "sft.jsonl")
    fixtures.write_sft_dataset(data, words, n=32)
    code = tmp_path / "my_exp.py"
    code.write_text(f'''
from realhf_b200.api.system import register_experiment
from realhf_b200.experiments.algos import SFTConfig


def make():
    c = SFTConfig(device="cpu", dtype="fp32", n_nodes=1, n_gpus_per_node=2, allocation_mode="manual")
    c.allocation.parallel.data_parallel_size = 2
    c.model.type._class, c.model.path = "llama", {ckpt!r}
    c.model.optimizer.grad_dtype, c.model.gradient_checkpointing = "fp32", False
    c.dataset.train_path, c.dataset.train_bs_n_seqs, c.dataset.max_seqlen = {data!r}, 8, 64
    c.exp_ctrl.total_train_epochs, c.exp_ctrl.benchmark_steps = 1, 2
    return c


register_experiment("mysft", make)
''')
    name = "mysft"
    M.main(["start", "-e", name, "-f", "t0", "--mode", "local", "--user_code", str(code), "--recover_mode", "disabled", "--timeout", "600",
            "--allocation_mode", "d2m1p1"])
    log = open(os.path.join(os.environ["REAL_FILEROOT"], "logs", name, "t0", "master_worker-0")).read()
    assert "benchmark finished" in log and log.count("[trainDefault]") == 2
    assert name in M.main(["find_config", "-r", "mys.*"])
    with pytest.raises(SystemExit):
        M.main(["start", "-e", "doesnotexist", "-f", "t0"])


def test_generation_replica_on_a_sub_mesh_aliases_the_training_weights(tmp_path):
    """actor_train dp2 on both workers, actor_gen dp1 on worker 1 only: the generation replica's shard on worker 1 IS the training
    shard (tp = pp = 1), so the reallocation aliases the flat buffer instead of copying it; training still changes what is generated
    from (the run completes, losses are finite, the alias is logged once)."""
    _env(tmp_path)
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    actor, critic = str(tmp_path / "actor"), str(tmp_path / "critic")
    cfg, tok, words = fixtures.make_checkpoint(actor, "llama")
    fixtures.make_checkpoint(critic, "llama", is_critic=True, seed=5)
    data = str(tmp_path / "prompts.jsonl")
    fixtures.write_prompt_dataset(data, words, n=32)
    args = ["ppo", f"experiment_name=alias-{uuid.uuid4().hex[:6]}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_gpus_per_node=2",
            "allocation_mode=manual", f"dataset.path={data}", "dataset.train_bs_n_seqs=8", "dataset.max_prompt_len=16",
            "ppo.gen.max_new_tokens=6", "ppo.gen.min_new_tokens=2", "ppo.gen.top_k=20", "ppo.ppo_n_minibatches=2",
            "exp_ctrl.total_train_epochs=1", "exp_ctrl.benchmark_steps=3"]
    for role, path in (("actor", actor), ("ref", actor), ("critic", critic), ("rew", critic)):
        args += [f"{role}.type._class=llama", f"{role}.path={path}", f"{role}.optimizer.grad_dtype=fp32", f"{role}.gradient_checkpointing=false"]
    args += ["actor_gen.parallel.data_parallel_size=1", "actor_gen.device_mesh=NODE01:1", "actor_train.parallel.data_parallel_size=2",
             "critic_train.parallel.data_parallel_size=2", "critic_inf.parallel.data_parallel_size=2",
             "ref_inf.parallel.data_parallel_size=2", "rew_inf.parallel.data_parallel_size=2"]
    exp = build_experiment(args)
    main_start(exp, timeout=900)
    log = _master_log(exp)
    assert log.count("[actor_train]") == 3 and "benchmark finished" in log, log[-3000:]
    w1 = open(os.path.join(os.environ["REAL_FILEROOT"], "logs", exp.experiment_name, "t0", "model_worker-1")).read()
    assert w1.count("aliased, no copy") == 1, w1[-3000:]


def test_every_dp_rank_of_a_train_mfc_gets_enough_sequences_for_its_minibatches(tmp_path):
    """16 prompts, critic_train on dp4 with 4 PPO minibatches: a purely token-balanced split hands some rank 3 sequences and another 5,
    the 3-sequence rank would take 3 optimizer steps while its peers take 4 and wait in the gradient collective of the 4th.  The master
    asks the partitioner for at least `n_minibatches` sequences per rank; when the batch is too small for that the interface fails
    with the reason (12 prompts) instead of desynchronising the group."""
    from realhf_b200.interfaces.ppo import _n_minibatches
    assert _n_minibatches(4, 7, None) == 4 and _n_minibatches(4, 3, None) == 3      # no DP group: fewer steps are harmless
    _env(tmp_path)
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    actor, critic = str(tmp_path / "actor"), str(tmp_path / "critic")
    cfg, tok, words = fixtures.make_checkpoint(actor, "llama")
    fixtures.make_checkpoint(critic, "llama", is_critic=True, seed=5)
    data = str(tmp_path / "prompts.jsonl")
    fixtures.write_prompt_dataset(data, words, n=48)
    args = ["ppo", f"experiment_name=mb-{uuid.uuid4().hex[:6]}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_gpus_per_node=4",
            "allocation_mode=manual", f"dataset.path={data}", "dataset.train_bs_n_seqs=16", "dataset.max_prompt_len=16",
            "ppo.gen.max_new_tokens=4", "ppo.gen.min_new_tokens=2", "ppo.gen.top_k=20", "ppo.ppo_n_minibatches=4",
            "exp_ctrl.total_train_epochs=1", "exp_ctrl.benchmark_steps=3"]
    for role, path in (("actor", actor), ("ref", actor), ("critic", critic), ("rew", critic)):
        args += [f"{role}.type._class=llama", f"{role}.path={path}", f"{role}.optimizer.grad_dtype=fp32", f"{role}.gradient_checkpointing=false"]
    args += ["actor_gen.parallel.data_parallel_size=4", "actor_train.parallel.data_parallel_size=2", "actor_train.parallel.model_parallel_size=2",
             "critic_train.parallel.data_parallel_size=4", "critic_inf.parallel.data_parallel_size=4",
             "ref_inf.parallel.data_parallel_size=4", "rew_inf.parallel.data_parallel_size=4"]
    exp = build_experiment(args)
    main_start(exp, timeout=900)
    log = _master_log(exp)
    assert log.count("[critic_train]") == 3 and "benchmark finished" in log, log[-3000:]
