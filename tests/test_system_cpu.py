"""Whole runtime on CPU/gloo: launcher -> master worker + 2 model workers (separate OS processes) -> GPT-2 SFT with
DP=2 on a synthetic dataset (BASELINE.json config #1), then PPO with parameter reallocation between two layouts."""
import os
import sys
import uuid

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fixtures  # noqa: E402

pytestmark = pytest.mark.distributed


def _env(tmp_path):
    os.environ["PYTHONPATH"] = ROOT + os.pathsep + os.environ.get("PYTHONPATH", "")
    os.environ["REAL_FILEROOT"] = str(tmp_path / "fileroot")
    os.environ["REAL_NAME_RESOLVE_ROOT"] = str(tmp_path / "nr")
    import importlib

    from realhf_b200.base import constants, name_resolve
    importlib.reload(constants)
    name_resolve.reconfigure("nfs", record_root=str(tmp_path / "nr"))


def test_gpt2_sft_dp2_gloo(tmp_path):
    _env(tmp_path)
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    ckpt = str(tmp_path / "gpt2")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "gpt2")
    data = str(tmp_path / "sft.jsonl")
    fixtures.write_sft_dataset(data, words, n=64)
    exp = build_experiment([
        "sft", f"experiment_name=sft-{uuid.uuid4().hex[:6]}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_nodes=1",
        "n_gpus_per_node=2", "allocation_mode=manual", "allocation.parallel.data_parallel_size=2", "model.type._class=gpt2",
        f"model.path={ckpt}", f"dataset.train_path={data}", "dataset.train_bs_n_seqs=16", "dataset.max_seqlen=64",
        "exp_ctrl.total_train_epochs=2", "exp_ctrl.save_freq_steps=4", "model.optimizer.lr=1e-3",
        "model.optimizer.warmup_steps_proportion=0.0", "model.optimizer.grad_dtype=fp32", "model.gradient_checkpointing=false",
        "tensorboard=True"])
    main_start(exp, timeout=600)
    log = open(os.path.join(os.environ["REAL_FILEROOT"], "logs", exp.experiment_name, "t0", "master_worker-0")).read()
    import glob
    assert glob.glob(os.path.join(os.environ["REAL_FILEROOT"], "logs", exp.experiment_name, "t0", "tensorboard", "events.out.tfevents.*"))
    losses = [float(l.split("loss=")[1].split(",")[0]) for l in log.splitlines() if "[trainDefault]" in l and "loss=" in l]
    assert len(losses) == 8, log[-3000:]
    assert losses[-1] < losses[0], losses
    import json
    stats = [json.loads(l) for l in open(os.path.join(os.environ["REAL_FILEROOT"], "logs", exp.experiment_name, "t0", "stats.jsonl"))]
    assert [r["step"] for r in stats] == list(range(8)) and abs(stats[0]["loss"] - losses[0]) < 1e-3
    save_root = os.path.join(os.environ["REAL_FILEROOT"], "checkpoints")
    found = [os.path.join(d, f) for d, _, fs in os.walk(save_root) for f in fs if f == "config.json"]
    assert found, "no checkpoint was written"
    # 4 steps per epoch, saved every 4 steps: named after the step that just finished (last step of its epoch)
    assert sorted(os.path.basename(os.path.dirname(f)) for f in found) == ["epoch0epochstep4globalstep4", "epoch1epochstep4globalstep8"]
    import transformers
    transformers.AutoModelForCausalLM.from_pretrained(os.path.dirname(found[0]))


def test_ppo_with_realloc_gloo(tmp_path):
    _env(tmp_path)
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    actor = str(tmp_path / "actor")
    critic = str(tmp_path / "critic")
    cfg, tok, words = fixtures.make_checkpoint(actor, "llama")
    fixtures.make_checkpoint(critic, "llama", is_critic=True, seed=5)
    data = str(tmp_path / "prompts.jsonl")
    fixtures.write_prompt_dataset(data, words, n=32)
    common = ["type._class=llama"]
    args = ["ppo", f"experiment_name=ppo-{uuid.uuid4().hex[:6]}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_gpus_per_node=2",
            "allocation_mode=manual", f"dataset.path={data}", "dataset.train_bs_n_seqs=8", "dataset.max_prompt_len=16",
            "ppo.gen.max_new_tokens=6", "ppo.gen.min_new_tokens=2", "ppo.gen.top_k=20", "ppo.ppo_n_minibatches=2",
            "exp_ctrl.total_train_epochs=1", "exp_ctrl.benchmark_steps=2"]
    for role, path in (("actor", actor), ("ref", actor), ("critic", critic), ("rew", critic)):
        args += [f"{role}.type._class=llama", f"{role}.path={path}", f"{role}.optimizer.grad_dtype=fp32", f"{role}.gradient_checkpointing=false"]
    # generation: dp2; actor training: tp2 (a different layout => parameter reallocation around actor_gen)
    args += ["actor_gen.parallel.data_parallel_size=2", "actor_train.parallel.model_parallel_size=2",
             "critic_train.parallel.data_parallel_size=2", "critic_inf.parallel.data_parallel_size=2",
             "ref_inf.parallel.data_parallel_size=2", "rew_inf.parallel.data_parallel_size=2"]
    exp = build_experiment(args)
    main_start(exp, timeout=900)
    log = open(os.path.join(os.environ["REAL_FILEROOT"], "logs", exp.experiment_name, "t0", "master_worker-0")).read()
    assert log.count("[actor_train]") == 2 and log.count("[critic_train]") == 2, log[-3000:]
    assert "benchmark finished" in log
    assert log.count("throughput:") == 2 and "TFLOP/s total" in log, log[-2000:]


def test_failure_detection_raises_instead_of_hanging(tmp_path):
    """A worker that dies during setup (unreadable dataset: the launcher's preflight only checks that the file exists) must
    surface as a JobException from the launcher."""
    _env(tmp_path)
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    from realhf_b200.scheduler.client import JobException
    ckpt = str(tmp_path / "gpt2")
    fixtures.make_checkpoint(ckpt, "gpt2")
    bad = tmp_path / "corrupt.jsonl"
    bad.write_text('{"id": 0, "prompt": "a", "answer": "b"}\n{this is not json\n')
    exp = build_experiment([
        "sft", f"experiment_name=bad-{uuid.uuid4().hex[:6]}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_nodes=1",
        "n_gpus_per_node=1", "allocation_mode=manual", "model.type._class=gpt2", f"model.path={ckpt}",
        f"dataset.train_path={bad}", "dataset.train_bs_n_seqs=8", "exp_ctrl.total_train_epochs=1"])
    with pytest.raises((JobException, TimeoutError)) as ei:
        main_start(exp, timeout=180)
    assert isinstance(ei.value, JobException), "the launcher should notice the failed worker long before the timeout"


def test_recover_resume_continues_from_saved_step(tmp_path):
    """recover_mode=save dumps RecoverInfo + model states at exit; a `resume` run restarts counting from there."""
    _env(tmp_path)
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    from realhf_b200.base import recover
    ckpt = str(tmp_path / "gpt2")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "gpt2")
    data = str(tmp_path / "sft.jsonl")
    fixtures.write_sft_dataset(data, words, n=32)
    name = f"rec-{uuid.uuid4().hex[:6]}"
    common = ["sft", f"experiment_name={name}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_nodes=1", "n_gpus_per_node=1",
              "allocation_mode=manual", "model.type._class=gpt2", f"model.path={ckpt}", f"dataset.train_path={data}",
              "dataset.train_bs_n_seqs=8", "dataset.max_seqlen=64", "exp_ctrl.total_train_epochs=2", "model.optimizer.grad_dtype=fp32",
              "model.gradient_checkpointing=false"]
    main_start(build_experiment(common + ["recover_mode=save", "exp_ctrl.benchmark_steps=3"]), timeout=600)
    info = recover.load_recover_info(name, "t0")
    assert info is not None and info.last_step_info.global_step >= 2, info
    main_start(build_experiment(common + ["recover_mode=resume", "exp_ctrl.benchmark_steps=5"]), timeout=300)
    root = os.path.join(os.environ["REAL_FILEROOT"], "logs", name, "t0")
    log = open(os.path.join(root, "master_worker-0")).read()
    steps = [int(l.split("] step ")[1].split(":")[0]) for l in log.splitlines() if "[trainDefault] step" in l]
    assert steps == [0, 1, 2, 3, 4], steps                       # the resumed run continues the global step count
    lrs = [float(l.split("lr=")[1].split(",")[0].split()[0]) for l in log.splitlines() if "[trainDefault] step" in l]
    assert lrs[3] < lrs[2], lrs                                   # LR schedule position restored with the optimizer state
    wlog = open(os.path.join(root, "model_worker-0")).read()
    assert "recover run: loading" in wlog, wlog[-1500:]


def test_auto_recover_after_injected_fault(tmp_path):
    """recover_mode=auto: a model worker dies in its 3rd train step (fault injection), the launcher saves the recover states,
    restarts everything as a recover run and the experiment finishes with a continuous step count."""
    _env(tmp_path)
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    ckpt = str(tmp_path / "gpt2")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "gpt2")
    data = str(tmp_path / "sft.jsonl")
    fixtures.write_sft_dataset(data, words, n=32)
    name = f"auto-{uuid.uuid4().hex[:6]}"
    exp = build_experiment(["sft", f"experiment_name={name}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_nodes=1", "n_gpus_per_node=1",
                            "allocation_mode=manual", "model.type._class=gpt2", f"model.path={ckpt}", f"dataset.train_path={data}",
                            "dataset.train_bs_n_seqs=8", "dataset.max_seqlen=64", "exp_ctrl.total_train_epochs=2",
                            "model.optimizer.grad_dtype=fp32", "model.gradient_checkpointing=false", "recover_mode=auto", "recover_retries=1"])
    os.environ["REAL_FAULT_INJECT"] = "0:train_step:3"
    try:
        main_start(exp, timeout=600)
    finally:
        os.environ.pop("REAL_FAULT_INJECT", None)
    log = open(os.path.join(os.environ["REAL_FILEROOT"], "logs", name, "t0", "master_worker-0")).read()
    steps = [int(l.split("] step ")[1].split(":")[0]) for l in log.splitlines() if "[trainDefault] step" in l]
    assert steps[:2] == [0, 1] and steps[-1] == 7 and len(steps) >= 8, steps
    wlog = open(os.path.join(os.environ["REAL_FILEROOT"], "logs", name, "t0", "model_worker-0")).read()
    assert "injected fault" in wlog and "recover run: loading" in wlog


def test_pause_resume_and_stop_through_the_controller(tmp_path):
    """pause: the master publishes PAUSED and no further step runs; resume continues; stop ends the run early and cleanly."""
    import threading
    import time
    _env(tmp_path)
    from realhf_b200.apps import main as M
    from realhf_b200.apps.quickstart import build_experiment
    from realhf_b200.apps.remote import status_key
    from realhf_b200.base import name_resolve
    ckpt = str(tmp_path / "gpt2")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "gpt2")
    data = str(tmp_path / "sft.jsonl")
    fixtures.write_sft_dataset(data, words, n=64)
    name = f"ctl-{uuid.uuid4().hex[:6]}"
    exp = build_experiment(["sft", f"experiment_name={name}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_nodes=1", "n_gpus_per_node=1",
                            "allocation_mode=manual", "model.type._class=gpt2", f"model.path={ckpt}", f"dataset.train_path={data}",
                            "dataset.train_bs_n_seqs=4", "dataset.max_seqlen=64", "exp_ctrl.total_train_epochs=40",
                            "model.optimizer.grad_dtype=fp32", "model.gradient_checkpointing=false"])
    err = []

    def runner():
        try:
            M.main_start(exp, timeout=600)
        except Exception as e:  # noqa: BLE001
            err.append(e)
    th = threading.Thread(target=runner, daemon=True)
    th.start()
    log_path = os.path.join(os.environ["REAL_FILEROOT"], "logs", name, "t0", "master_worker-0")
    skey = status_key(name, "t0", "master_worker", 0)

    def wait_for(pred, what, timeout=180):
        t0 = time.time()
        while time.time() - t0 < timeout:
            if pred():
                return
            time.sleep(0.2)
        raise AssertionError(f"timed out waiting for {what}")

    def status():
        try:
            return name_resolve.get(skey)
        except name_resolve.NameEntryNotFoundError:
            return None

    n_steps = lambda: open(log_path).read().count("[trainDefault] step") if os.path.exists(log_path) else 0
    wait_for(lambda: status() == "RUNNING", "the master to come up")  # main_start wipes the trial's keys before launching
    M.pause_experiment(name, "t0")
    wait_for(lambda: status() == "PAUSED", "PAUSED status")
    frozen = n_steps()
    time.sleep(1.5)
    assert n_steps() == frozen
    M.resume_experiment(name, "t0")
    wait_for(lambda: n_steps() >= frozen + 3, "steps after resume")
    M.pause_experiment(name, "t0")
    wait_for(lambda: status() == "PAUSED", "second pause")
    frozen = n_steps()
    time.sleep(1.0)
    assert n_steps() == frozen
    M.stop_experiment(name, "t0")
    th.join(timeout=120)
    assert not th.is_alive() and not err, err
    log = open(log_path).read()
    assert "stop requested by the controller" in log and n_steps() < 16 * 40


def test_workers_exit_when_the_controller_is_killed_and_lost_status(tmp_path):
    """Orphan protection: SIGKILL the launcher -> its liveness lease expires -> master and model workers exit on their own.
    Also the controller-side view: a status whose lease ran out reads LOST."""
    import re
    import subprocess
    import sys
    import time

    import psutil
    _env(tmp_path)
    from realhf_b200.apps import main as M
    from realhf_b200.apps.remote import status_key
    from realhf_b200.base import name_resolve
    ckpt = str(tmp_path / "gpt2")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "gpt2")
    data = str(tmp_path / "sft.jsonl")
    fixtures.write_sft_dataset(data, words, n=64)
    name = f"orphan-{uuid.uuid4().hex[:6]}"
    env = dict(os.environ, REAL_STATUS_TTL="5")
    launcher_log = str(tmp_path / "launcher.log")
    with open(launcher_log, "w") as lf:
        p = subprocess.Popen([sys.executable, "-m", "realhf_b200.apps.quickstart", "sft", f"experiment_name={name}", "trial_name=t0",
                              "device=cpu", "dtype=fp32", "n_nodes=1", "n_gpus_per_node=1", "allocation_mode=manual",
                              "model.type._class=gpt2", f"model.path={ckpt}", f"dataset.train_path={data}", "dataset.train_bs_n_seqs=4",
                              "dataset.max_seqlen=64", "exp_ctrl.total_train_epochs=200", "model.optimizer.grad_dtype=fp32",
                              "model.gradient_checkpointing=false"], env=env, stdout=lf, stderr=subprocess.STDOUT)
    pids = []
    try:
        master_log = os.path.join(os.environ["REAL_FILEROOT"], "logs", name, "t0", "master_worker-0")
        t0 = time.time()
        while time.time() - t0 < 240:
            if os.path.exists(master_log) and "[trainDefault] step" in open(master_log).read():
                break
            assert p.poll() is None, open(launcher_log).read()[-3000:]
            time.sleep(0.5)
        else:
            raise AssertionError("the run never started stepping")
        pids = [int(x) for x in re.findall(r"started (?:master|model)_worker/\d+ \(pid (\d+)\)", open(launcher_log).read())]
        assert len(pids) == 2, open(launcher_log).read()[-2000:]
        p.kill()   # SIGKILL: no cleanup handler runs, the workers are in their own sessions and keep running
        p.wait()

        def dead(pid):
            try:
                return psutil.Process(pid).status() == psutil.STATUS_ZOMBIE
            except psutil.NoSuchProcess:
                return True
        t0 = time.time()
        while time.time() - t0 < 60 and not all(dead(x) for x in pids):
            time.sleep(0.5)
        assert all(dead(x) for x in pids), "workers survived the controller"
        assert "liveness key expired" in open(master_log).read()
    finally:
        if p.poll() is None:
            p.kill()
        for x in pids:
            try:
                os.kill(x, 9)
            except OSError:
                pass

    # controller-side: a published status whose lease is stale reads LOST, an unpublished one UNKNOWN
    class _Sched:
        run_name = "x"
    ctl = M.Controller(name, "t1", _Sched(), 2)
    repo = name_resolve.DEFAULT_REPOSITORY
    repo.add(status_key(name, "t1", "model_worker", 0), "RUNNING", replace=True, keepalive_ttl=600)
    assert ctl.statuses() == {"master_worker/0": "UNKNOWN", "model_worker/0": "RUNNING", "model_worker/1": "UNKNOWN"}
    repo._keepalive.pop(status_key(name, "t1", "model_worker", 0))   # stop refreshing, then age the entry past its TTL
    old = time.time() - 3600
    os.utime(repo._file(status_key(name, "t1", "model_worker", 0)), (old, old))
    assert ctl.statuses()["model_worker/0"] == "LOST"


def test_dataset_size_not_a_multiple_of_the_batch_size_across_epochs(tmp_path):
    """20 samples, batches of 8, 2 epochs: the 4 left over from the first epoch wait in the buffer while the second epoch's
    data (the same ids again) is fetched -- the run must neither trip over duplicate ids nor starve."""
    _env(tmp_path)
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    ckpt = str(tmp_path / "gpt2")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "gpt2")
    data = str(tmp_path / "sft.jsonl")
    fixtures.write_sft_dataset(data, words, n=20)
    name = f"odd-{uuid.uuid4().hex[:6]}"
    exp = build_experiment(["sft", f"experiment_name={name}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_nodes=1", "n_gpus_per_node=1",
                            "allocation_mode=manual", "model.type._class=gpt2", f"model.path={ckpt}", f"dataset.train_path={data}",
                            "dataset.train_bs_n_seqs=8", "dataset.max_seqlen=64", "exp_ctrl.total_train_epochs=2",
                            "model.optimizer.grad_dtype=fp32", "model.gradient_checkpointing=false"])
    main_start(exp, timeout=600)
    log = open(os.path.join(os.environ["REAL_FILEROOT"], "logs", name, "t0", "master_worker-0")).read()
    steps = [int(l.split("] step ")[1].split(":")[0]) for l in log.splitlines() if "[trainDefault] step" in l]
    assert steps == list(range(4)), steps   # 20 // 8 = 2 steps per epoch


def test_graceful_stop_saves_recover_states_and_resume_continues(tmp_path):
    """`stop_experiment` on a run launched with recover_mode=save ends it cleanly after the current step and leaves recover
    states; a `resume` run picks up at the next step (no step repeated, none skipped)."""
    import threading
    import time
    _env(tmp_path)
    from realhf_b200.apps import main as M
    from realhf_b200.apps.quickstart import build_experiment
    ckpt = str(tmp_path / "gpt2")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "gpt2")
    data = str(tmp_path / "sft.jsonl")
    fixtures.write_sft_dataset(data, words, n=64)
    name = f"sr-{uuid.uuid4().hex[:6]}"

    def args(mode):
        return ["sft", f"experiment_name={name}", "trial_name=t0", "device=cpu", "dtype=fp32", "n_gpus_per_node=1", "allocation_mode=manual",
                "model.type._class=gpt2", f"model.path={ckpt}", f"dataset.train_path={data}", "dataset.train_bs_n_seqs=4",
                "dataset.max_seqlen=64", "exp_ctrl.total_train_epochs=30", "model.optimizer.grad_dtype=fp32",
                "model.gradient_checkpointing=false", f"recover_mode={mode}"]
    log_path = os.path.join(os.environ["REAL_FILEROOT"], "logs", name, "t0", "master_worker-0")

    def steps():
        if not os.path.exists(log_path):
            return []
        return [int(l.split("] step ")[1].split(":")[0]) for l in open(log_path).read().splitlines() if "[trainDefault] step" in l]
    err = []

    def first_run():
        try:
            M.main_start(build_experiment(args("save")), timeout=600)
        except Exception as e:  # noqa: BLE001
            err.append(e)
    th = threading.Thread(target=first_run, daemon=True)
    th.start()
    t0 = time.time()
    while len(steps()) < 3 and time.time() - t0 < 240:
        time.sleep(0.2)
    M.stop_experiment(name, "t0")
    th.join(timeout=240)
    assert not th.is_alive() and not err, err
    s1 = steps()
    assert 3 <= len(s1) < 30 * 16 and s1 == list(range(len(s1)))
    assert os.path.exists(os.path.join(os.environ["REAL_FILEROOT"], "recover", name, "t0", "recover_info.pkl"))
    M.main_start(build_experiment(args("resume") + [f"exp_ctrl.benchmark_steps={s1[-1] + 4}"]), timeout=600)
    s2 = steps()
    assert s2[: len(s1) + 3] == list(range(len(s1) + 3)), (s1[-3:], s2[len(s1) - 2: len(s1) + 4])


def test_bench_master_runtime_arm_tiny(tmp_path):
    """`bench.py --runtime master` drives the benchmark config through quickstart -> launcher -> master + model workers and
    prints the bench JSON line (toy shapes on CPU here; the GPU box runs it with the headline shapes)."""
    import json
    import subprocess
    import sys as _sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, REAL_FILEROOT=str(tmp_path / "fileroot"), REAL_NAME_RESOLVE_ROOT=str(tmp_path / "nr"))
    p = subprocess.run([_sys.executable, os.path.join(root, "bench.py"), "--runtime", "master", "--tiny", "--gpus", "2", "--layers", "2",
                        "--prompts", "8", "--prompt-len", "8", "--new-tokens", "6", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900, env=env, cwd=str(tmp_path))
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["n_gpus"] == 2 and out["steps"] == 2 and out["value"] > 0
    assert set(out["config"]["mfc_ms"]) == {"actor_gen", "rew_inf", "ref_inf", "critic_inf", "actor_train", "critic_train"}
