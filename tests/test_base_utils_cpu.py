"""Unit tests of the plumbing utilities: name_resolve backends, frequency controls, device-mesh naming, the dotted-override parser,
the Slurm script builder and the process-topology rank math."""
import os
import threading
import time

import numpy as np
import pytest

from realhf_b200.api import quickstart as Q
from realhf_b200.base import name_resolve, timeutil
from realhf_b200.base.topology import ProcessTopology


@pytest.mark.parametrize("kind", ["memory", "nfs", "redis"])
def test_name_resolve_backends(tmp_path, kind):
    srv = None
    if kind == "redis":  # the repository speaks RESP2 itself; the test double is the bundled mini server (with AUTH)
        srv = name_resolve.MiniRedisServer(password="s3cret").start()
        repo = name_resolve.make_repository("redis", host=srv.host, port=srv.port, password="s3cret")
    else:
        repo = name_resolve.make_repository(kind, **({"record_root": str(tmp_path)} if kind == "nfs" else {}))
    repo.add("exp/t/a", "1")
    with pytest.raises(name_resolve.NameEntryExistsError):
        repo.add("exp/t/a", "2")
    repo.add("exp/t/a", "2", replace=True)
    assert repo.get("exp/t/a") == "2"
    repo.add_subentry("exp/t/peers", "x")
    repo.add_subentry("exp/t/peers", "y")
    assert sorted(repo.get_subtree("exp/t/peers")) == ["x", "y"]
    assert len(repo.find_subtree("exp/t")) >= 3
    with pytest.raises(name_resolve.NameEntryNotFoundError):
        repo.get("exp/t/missing")
    # wait() sees a key published later by another thread
    threading.Timer(0.2, lambda: repo.add("exp/t/late", "ok")).start()
    assert repo.wait("exp/t/late", timeout=5) == "ok"
    with pytest.raises(TimeoutError):
        repo.wait("exp/t/never", timeout=0.2)
    repo.delete("exp/t/a")
    with pytest.raises(name_resolve.NameEntryNotFoundError):
        repo.get("exp/t/a")
    repo.clear_subtree("exp/t")
    assert repo.find_subtree("exp/t") == []
    repo.reset()
    if srv is not None:
        srv.stop()


def test_redis_lease_expires_without_keepalive_and_survives_with_it():
    """TTL'd keys are leases: the server drops them when the owner stops refreshing (a silently dead worker reads LOST), the
    keep-alive thread of a live owner keeps them; a second client sees both; wrong passwords are refused."""
    srv = name_resolve.MiniRedisServer(password="pw").start()
    owner = name_resolve.make_repository("redis", host=srv.host, port=srv.port, password="pw")
    owner.KEEPALIVE_POLL = 0.1
    other = name_resolve.make_repository("redis", host=srv.host, port=srv.port, password="pw")
    owner.add("lease/kept", "1", keepalive_ttl=1)
    other._conn.call("SET", "lease/orphan", "1", "EX", 1)  # nobody refreshes this one
    time.sleep(2.2)
    assert other.get("lease/kept") == "1"
    with pytest.raises(name_resolve.NameEntryNotFoundError):
        other.get("lease/orphan")
    owner.reset()  # delete_on_exit keys go away with their owner
    with pytest.raises(name_resolve.NameEntryNotFoundError):
        other.get("lease/kept")
    with pytest.raises(RuntimeError):
        name_resolve.make_repository("redis", host=srv.host, port=srv.port, password="wrong").get("x")
    srv.stop()


def test_watch_names_fires_when_a_key_disappears(tmp_path):
    repo = name_resolve.make_repository("nfs", record_root=str(tmp_path))
    repo.add("w/alive", "1")
    fired = threading.Event()
    repo.watch_names(["w/alive"], fired.set, poll_frequency=0.1, wait_timeout=5)
    time.sleep(0.3)
    assert not fired.is_set()
    repo.delete("w/alive")
    assert fired.wait(5)


def test_frequency_controls():
    f = timeutil.FrequencyControl(frequency_steps=3)
    assert [f.check() for _ in range(7)] == [False, False, True, False, False, True, False]
    f2 = timeutil.FrequencyControl(frequency_steps=3)
    f2.load_state_dict(f.state_dict())
    assert f2.check() is False and f2.check() is True  # count carried over: 1 -> 2 -> 3
    t = timeutil.FrequencyControl(frequency_seconds=0.15)
    assert t.check() is False
    time.sleep(0.2)
    assert t.check() is True and t.check() is False
    assert timeutil.FrequencyControl(frequency_steps=100, initial_value=True).check() is True
    ctl = timeutil.EpochStepTimeFreqCtl(freq_epoch=1, freq_step=2)
    fires = [ctl.check(epochs=int(i == 3), steps=1) for i in range(5)]
    assert fires == [False, True, False, True, True] or fires[1] and fires[3]


def test_device_mesh_names_and_strategies():
    m = Q.make_device_mesh_from_name("NODE[01-02]", "NODE01:0,1,2,3", n_nodes=2, n_gpus_per_node=8)
    assert m.n_gpus == 4 and m.mapping[0, :4].all() and not m.mapping[1].any()
    whole = Q.make_device_mesh_from_name("NODE[01-02]", "NODE[01-02]", n_nodes=2, n_gpus_per_node=8)
    assert whole.n_gpus == 16 and whole.contain(m) and whole.overlap(m)
    other = Q.make_device_mesh_from_name("NODE[01-02]", "NODE02:4,5,6,7", n_nodes=2, n_gpus_per_node=8)
    assert not other.overlap(m)
    for bad in ("NODE01:1,2", "NODE01:0,2", "NODE01:0,1,2"):
        with pytest.raises(ValueError):
            Q.make_device_mesh_from_name(None, bad, n_nodes=1, n_gpus_per_node=8)
    subs = whole.sub_device_meshes()
    assert any(s.n_gpus == 1 for s in subs) and any(s.n_gpus == 8 for s in subs) and all(s.n_gpus in (1, 2, 4, 8, 16) for s in subs)
    strategies = Q.find_parallel_strategies(m)
    assert {(p.model_parallel_size, p.pipeline_parallel_size, p.data_parallel_size) for p in strategies} == \
        {(1, 1, 4), (1, 2, 2), (1, 4, 1), (2, 1, 2), (2, 2, 1), (4, 1, 1)}


def test_dotted_overrides_cover_the_reference_surface():
    from realhf_b200.experiments.algos import PPOConfig
    cfg = PPOConfig(experiment_name="a", trial_name="b")
    Q.parse_overrides(cfg, ["actor.optimizer.lr=1e-4", "actor_train.parallel.model_parallel_size=2", "exp_ctrl.save_freq_steps=null",
                            "ppo.gen.top_k=50", "ppo.gen.use_cuda_graph=True", "actor.type._class=qwen2", "actor.type.size=13",
                            "actor_gen.n_mbs=4", "ppo.kl_ctl=0.05", "dataset.path=/x/y.jsonl"])
    assert cfg.actor.optimizer.lr == 1e-4 and cfg.actor_train.parallel.model_parallel_size == 2 and cfg.exp_ctrl.save_freq_steps is None
    assert cfg.ppo.gen.top_k == 50 and cfg.ppo.gen.use_cuda_graph is True and cfg.actor.type._class == "qwen2" and cfg.actor.type.size == 13
    assert cfg.actor_gen.n_mbs == 4 and cfg.ppo.kl_ctl == 0.05 and cfg.dataset.path == "/x/y.jsonl"
    with pytest.raises((AttributeError, KeyError, ValueError)):
        Q.parse_overrides(cfg, ["actor.no_such_field=1"])


def test_slurm_script_builder(tmp_path, monkeypatch):
    monkeypatch.setenv("REAL_FILEROOT", str(tmp_path))
    import importlib

    from realhf_b200.base import constants
    importlib.reload(constants)
    from realhf_b200.scheduler import client as C
    importlib.reload(C)
    s = C.SlurmSchedulerClient("exp", "trial", partition="gpu", container_image="img:latest")
    cmd = C.remote_worker_cmd("exp", "trial", True, "model_worker")
    script = s.build_script("model_worker", cmd, count=16, cpu=8, gpu=1, mem=64000, nodelist="n[1-2]", time_limit="1:00:00",
                            env_vars={"REAL_MODE": "SLURM", "X": "a b"})
    assert "#SBATCH --ntasks=16" in script and "#SBATCH --nodes=2" in script and "#SBATCH --gpus-per-task=1" in script
    assert "#SBATCH --nodelist=n[1-2]" in script and "export X='a b'" in script and "--container-image=img:latest" in script
    mp = [l for l in script.splitlines() if l.startswith("srun")][0].split()[-1]
    lines = open(mp).read().splitlines()
    assert len(lines) == 16 and lines[3].startswith("3 ") and "-i 3 -g 16" in lines[3] and "-w model_worker" in lines[3]


def test_process_topology_rank_math():
    topo = ProcessTopology(2, 3, 4)  # (pp, dp, tp)
    assert topo.world_size() == 24
    seen = set()
    for r in range(24):
        c = topo.get_coord(r)
        assert topo.get_rank(pipe=c.pipe, data=c.data, model=c.model) == r
        seen.add((c.pipe, c.data, c.model))
    assert len(seen) == 24
    # tensor-parallel peers are adjacent ranks (they share a node / NVLink domain)
    c0 = topo.get_coord(0)
    assert sorted(topo.get_rank(pipe=c0.pipe, data=c0.data, model=t) for t in range(4)) == [0, 1, 2, 3]


def test_layer_profiler_cpu(tmp_path, monkeypatch):
    """The layer profiler that feeds the allocation search runs on CPU (wall clock) and writes its table."""
    import torch

    from realhf_b200.models import hf_io
    from realhf_b200.search import layers
    monkeypatch.setattr(layers.constants, "PROFILER_CACHE_PATH", str(tmp_path))
    cfg = hf_io.family("llama").make_test_config()
    rows = layers.profile_layers(cfg, [2], [16, 32], device="cpu", dtype=torch.float32)
    assert {(r["layer"], r["op"]) for r in rows} == {("block", "fwd"), ("block", "fwd_bwd"), ("embedding", "fwd")}
    assert all(r["time_us"] == r["time_us"] for r in rows)
    p = layers.dump_profile(rows, "llama-test")
    assert os.path.exists(p)


def test_sequence_buffer_readiness_and_reuse():
    """AsyncIOSequenceBuffer: a batch is ready for an MFC when all its input keys are present; slots are freed once every
    consumer has read them."""
    import asyncio

    import torch

    from realhf_b200.api.config import ModelInterfaceAbstraction, ModelInterfaceType
    from realhf_b200.api.data import SequenceSample
    from realhf_b200.api.dfg import MFCDef, build_graph
    from realhf_b200.system.buffer import AsyncIOSequenceBuffer
    A = lambda t: ModelInterfaceAbstraction(t, {})
    gen = MFCDef("gen", 2, ModelInterfaceType.GENERATE, A("x"), "actor", input_keys=("packed_prompts",), output_keys=("packed_input_ids",))
    train = MFCDef("train", 2, ModelInterfaceType.TRAIN_STEP, A("x"), "actor", input_keys=("packed_input_ids",))
    build_graph([gen, train])
    buf = AsyncIOSequenceBuffer([gen, train], max_size=16)

    def meta(i, key):
        return SequenceSample.from_default(seqlens=[3], ids=[i], data={key: torch.zeros(3)}).meta()

    async def run():
        await buf.put_batch([meta(i, "packed_prompts") for i in range(3)])
        assert buf.n_ready_for(gen) == 3 and buf.n_ready_for(train) == 0
        ids, _ = await buf.get_batch_for_rpc(gen)
        assert len(ids) == 2
        await buf.amend_batch(ids, [meta(i, "packed_input_ids") for i in ids])
        assert buf.n_ready_for(train) == 2
        ids2, _ = await buf.get_batch_for_rpc(train)
        assert sorted(ids2) == sorted(ids)
        done = buf.pop_fully_consumed()
        assert sorted(done) == sorted(ids) and buf.n_ready_for(gen) == 1

    asyncio.run(run())


def test_name_resolve_lease_expires_when_the_owner_dies(tmp_path):
    """A key written with a TTL is kept alive by its owner's background thread and reads as missing once the owner is gone."""
    import subprocess
    import sys
    import time
    from realhf_b200.base import name_resolve
    root = str(tmp_path / "nr")
    code = ("import sys, time\n"
            "from realhf_b200.base import name_resolve\n"
            f"r = name_resolve.NfsNameRecordRepository({root!r})\n"
            "r.add('a/b/status', 'RUNNING', keepalive_ttl=3.0)\n"
            "r.add('a/b/final', 'COMPLETED')\n"
            "print('up', flush=True)\n"
            "time.sleep(120)\n")
    p = subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, text=True)
    try:
        assert p.stdout.readline().strip() == "up"
        repo = name_resolve.NfsNameRecordRepository(root)
        time.sleep(4.5)   # 1.5 TTLs: the owner's keep-alive thread (one touch per second) must have refreshed the lease
        assert repo.get("a/b/status") == "RUNNING"
        p.kill()
        p.wait()
        gone = []
        repo.watch_names(["a/b/status"], lambda: gone.append(1), poll_frequency=0.2)
        t0 = time.time()
        while not gone and time.time() - t0 < 20:
            time.sleep(0.1)
        assert gone, "watch_names did not fire after the lease expired"
        with pytest.raises(name_resolve.NameEntryNotFoundError):
            repo.get("a/b/status")
        assert repo.get("a/b/final") == "COMPLETED"   # keys without a lease never expire
        repo.add("a/b/status", "COMPLETED", replace=True)  # a terminal status drops the lease
        time.sleep(2.0)
        assert repo.get("a/b/status") == "COMPLETED"
    finally:
        if p.poll() is None:
            p.kill()


def test_cluster_spec_drives_fileroot_slurm_defaults_and_search_memory(tmp_path, monkeypatch):
    import importlib
    import json

    from realhf_b200.base import cluster
    example = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "cluster_config.json")
    raw = json.load(open(example))
    raw.update(fileroot=str(tmp_path / "shared"), gpu_memory_gb=141, n_gpus_per_node=4)
    path = str(tmp_path / "cluster.json")
    json.dump(raw, open(path, "w"))
    monkeypatch.setenv("CLUSTER_SPEC_PATH", path)
    monkeypatch.delenv("REAL_FILEROOT", raising=False)
    cs = cluster.spec()
    assert cs.cluster_name == "blackwell-pod-a" and cs.node_type("NODE07") == "b200x8" and cs.gpu_type("NODE12") == "b200"
    assert cluster.node_name_is_node_type("NODE01", ["x", "b200x8"]) and not cluster.node_name_is_node_type("NODE01", "a100")
    assert cs.node_names([1, 12]) == ["NODE01", "NODE12"]
    with pytest.raises(KeyError):
        cs.node_type("login-1")
    from realhf_b200.base import constants
    importlib.reload(constants)
    try:
        assert constants.FILEROOT == str(tmp_path / "shared")
        from realhf_b200.scheduler import client as C
        importlib.reload(C)
        s = C.SlurmSchedulerClient("exp", "trial")
        script = s.build_script("model_worker", C.remote_worker_cmd("exp", "trial", True, "model_worker"), count=8, gpu=1)
        assert "#SBATCH --partition=b200" in script and "#SBATCH --nodes=2" in script          # 8 workers on 4-GPU nodes
        assert "--container-image=registry.example.com/rlhf/realhf-b200:gpu" in script and "/dev/infiniband:/dev/infiniband" in script
        script = s.build_script("master_worker", C.remote_worker_cmd("exp", "trial", True, "master_worker"), count=1, gpu=0)
        assert "--container-image=registry.example.com/rlhf/realhf-b200:cpu" in script
        from realhf_b200.search.engine import HardwareModel
        assert HardwareModel.from_measured().mem_cap == 141e9
        # malformed specs fail loudly
        bad = str(tmp_path / "bad.json")
        json.dump({"cluster_type": "slurm", "cluster_name": "x", "fileroot": "/x", "gpu_count": 8}, open(bad, "w"))
        with pytest.raises(ValueError, match="unknown keys"):
            cluster.ClusterSpec.from_file(bad)
    finally:
        monkeypatch.delenv("CLUSTER_SPEC_PATH")
        monkeypatch.setenv("REAL_FILEROOT", str(tmp_path / "fr"))
        importlib.reload(constants)
    assert cluster.spec().cluster_type == "local" and cluster.spec().gpu_memory_gb == 180


def test_config_reference_doc_is_up_to_date():
    """docs/expconfig.md is generated from the dataclasses; a renamed or added option must be regenerated with the code."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "gen_config_docs.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_nodelist_parsing_and_formatting_roundtrip():
    from realhf_b200.api.quickstart import make_device_mesh_from_name
    from realhf_b200.base.cluster import format_nodelist, parse_nodelist
    assert parse_nodelist("NODE[01-03,07],gpu12") == ["NODE01", "NODE02", "NODE03", "NODE07", "gpu12"]
    assert parse_nodelist("NODE01,NODE02") == ["NODE01", "NODE02"] and parse_nodelist("n[9-11]") == ["n9", "n10", "n11"]
    assert format_nodelist(["NODE01", "NODE02", "NODE03", "NODE07"]) == "NODE[01-03,07]" and format_nodelist(["NODE05"]) == "NODE05"
    for s in ("NODE[01-04]", "NODE[01-02,05]", "a1,b[3-4]"):
        assert parse_nodelist(format_nodelist(parse_nodelist(s))) == parse_nodelist(s)
    for bad in ("NODE[03-01]", "NODE01,NODE01", ""):
        with pytest.raises(ValueError):
            parse_nodelist(bad)
    m = make_device_mesh_from_name("NODE[01-04]", "NODE[01-02,04]", n_nodes=4, n_gpus_per_node=8)
    assert m.mapping.sum(1).tolist() == [8, 8, 0, 8]


def test_operations_cli_status_pause_stop_find_config(tmp_path, monkeypatch, capsys):
    """`python -m realhf_b200.apps.main ...` from another shell: talks to a trial only through the name-resolve store."""
    from realhf_b200.apps import main as M
    from realhf_b200.apps.remote import control_key, status_key
    from realhf_b200.base import name_resolve
    name_resolve.reconfigure("nfs", record_root=str(tmp_path / "nr"))
    name_resolve.add(status_key("e1", "t1", "master_worker", 0), "RUNNING", replace=True)
    name_resolve.add(status_key("e1", "t1", "model_worker", 0), "ERROR", replace=True)
    out = M.main(["status", "-e", "e1", "-f", "t1"])
    assert out == {"master_worker/0": "RUNNING", "model_worker/0": "ERROR"} and "model_worker/0: ERROR" in capsys.readouterr().out
    M.main(["pause", "-e", "e1", "-f", "t1"])
    assert name_resolve.get(control_key("e1", "t1", "master_worker", 0)) == "pause"
    M.main(["resume", "-e", "e1", "-f", "t1"])
    assert name_resolve.get(control_key("e1", "t1", "master_worker", 0)) == "resume"
    calls = []
    import subprocess
    monkeypatch.setattr(subprocess, "run", lambda cmd, **kw: calls.append(cmd))
    M.main(["stop", "-e", "e1", "-f", "t1", "--mode", "slurm"])
    assert name_resolve.get(control_key("e1", "t1", "master_worker", 0)) == "exit"
    assert calls == [["scancel", "--name", "e1_t1:master_worker"], ["scancel", "--name", "e1_t1:model_worker"]]
    assert M.main(["find_config", "-r", "p(po|rofile)$"]) == ["ppo", "profile"]   # example experiments may be registered too
    with pytest.raises(SystemExit):
        M.main(["status", "-e", "e1"])   # trial name is required


def test_lora_options_are_rejected_loudly(tmp_path, monkeypatch):
    monkeypatch.setenv("REAL_FILEROOT", str(tmp_path))
    from realhf_b200.apps.quickstart import build_experiment
    exp = build_experiment(["rw", "experiment_name=e", "trial_name=t", "device=cpu", "n_gpus_per_node=1", "is_sft_lora=True"])
    with pytest.raises(NotImplementedError, match="LoRA"):
        exp.initial_setup()


def test_quickstart_help_lists_every_option(capsys):
    from realhf_b200.apps.quickstart import build_experiment
    with pytest.raises(SystemExit) as e:
        build_experiment(["ppo", "--help"])
    out = capsys.readouterr().out
    assert e.value.code == 0 and "actor_train.parallel.model_parallel_size" in out and "ppo.gen.max_new_tokens" in out
    with pytest.raises(SystemExit):
        build_experiment(["no-such-experiment"])


def test_package_root_exports_the_public_names():
    import realhf_b200
    from realhf_b200 import MFCDef, PPOConfig, SequenceSample  # noqa: F401  (the spelling user code uses)
    for name in realhf_b200.__all__:
        assert getattr(realhf_b200, name) is not None
    assert realhf_b200.PPOConfig().ppo.kl_ctl == 0.1 and "PPOConfig" in dir(realhf_b200)
    with pytest.raises(AttributeError):
        realhf_b200.NoSuchName


def test_generation_options_are_validated_on_the_command_line():
    from realhf_b200.apps.quickstart import build_experiment
    with pytest.raises(SystemExit, match="ppo.gen.*min_new_tokens > max_new_tokens"):
        build_experiment(["ppo", "experiment_name=e", "trial_name=t", "ppo.gen.max_new_tokens=6"])
    exp = build_experiment(["ppo", "experiment_name=e", "trial_name=t", "ppo.gen.max_new_tokens=6", "ppo.gen.min_new_tokens=2",
                            "ppo.gen.temperature=0.0"])
    assert exp.ppo.gen.greedy and exp.ppo.gen.temperature == 1.0   # temperature 0 means greedy, as in the dataclass's own check


def test_launcher_preflight_names_every_bad_path(tmp_path, monkeypatch):
    monkeypatch.setenv("REAL_FILEROOT", str(tmp_path))
    from realhf_b200.apps.main import main_start, preflight
    from realhf_b200.apps.quickstart import build_experiment
    good = tmp_path / "ckpt"
    good.mkdir()
    (good / "config.json").write_text("{}")
    data = tmp_path / "d.jsonl"
    data.write_text("{}\\n")
    exp = build_experiment(["dpo", "experiment_name=e", "trial_name=t", "device=cpu", f"actor.path={good}", "ref.path=/no/such/dir",
                            "dataset.train_path=/no/such/file.jsonl"])
    with pytest.raises(FileNotFoundError) as e:
        main_start(exp)          # fails before any process is started
    msg = str(e.value)
    assert "ref.path: `/no/such/dir` is not a directory" in msg and "/no/such/file.jsonl" in msg and "actor.path" not in msg
    exp = build_experiment(["dpo", "experiment_name=e", "trial_name=t", "device=cpu", f"actor.path={good}", f"ref.path={good}",
                            f"dataset.train_path={data}"])
    preflight(exp)               # nothing to complain about


def test_network_and_importing_helpers(tmp_path, monkeypatch):
    import socket
    import sys

    from realhf_b200.base import importing, network
    p = network.find_free_port()
    with socket.socket() as s:
        s.bind(("127.0.0.1", p))           # still free
    monkeypatch.setenv("REAL_HOST_IP", "10.1.2.3")
    assert network.gethostip() == "10.1.2.3"
    monkeypatch.delenv("REAL_HOST_IP")
    monkeypatch.setenv("REAL_MODE", "LOCAL")
    assert network.gethostip() == "127.0.0.1"
    monkeypatch.setenv("REAL_MODE", "SLURM")
    socket.inet_aton(network.gethostip())  # some valid IPv4 address, whatever this container resolves to
    code = tmp_path / "my_exp.py"
    code.write_text("import dataclasses\n@dataclasses.dataclass\nclass C:\n    x: int = 3\nVALUE = C().x + 1\n")
    mod = importing.import_usercode(str(code), "my_user_code")
    assert mod.VALUE == 4 and sys.modules["my_user_code"] is mod
    import pickle
    assert pickle.loads(pickle.dumps(mod.C(5))).x == 5          # classes resolve through sys.modules
    bad = tmp_path / "broken.py"
    bad.write_text("raise RuntimeError('boom')\n")
    with pytest.raises(RuntimeError):
        importing.import_usercode(str(bad), "broken_user_code")
    assert "broken_user_code" not in sys.modules
    with pytest.raises(FileNotFoundError):
        importing.import_usercode(str(tmp_path / "missing.py"))
    names = importing.import_package_modules("realhf_b200.utils")
    assert "realhf_b200.utils.padding" in names and all(not n.rsplit(".", 1)[1].startswith("_") for n in names)


def test_security_read_key_and_saveload_helpers(tmp_path, monkeypatch):
    import torch

    from realhf_b200.base import saveload_utils, security
    (tmp_path / "keys" / "wandb").mkdir(parents=True)
    (tmp_path / "keys" / "wandb" / "default").write_text("  secret-token\n")
    monkeypatch.setenv("REAL_KEY_ROOT", str(tmp_path / "keys"))
    assert security.read_key("wandb") == "secret-token"
    sd = {f"w{i}": torch.zeros(100, dtype=torch.float32) for i in range(5)}            # 400 bytes each
    shards = saveload_utils.split_state_dict_into_shards(sd, max_bytes=900)
    assert [sorted(s) for s in shards] == [["w0", "w1"], ["w2", "w3"], ["w4"]]
    assert len(saveload_utils.split_state_dict_into_shards({"big": torch.zeros(1000)}, max_bytes=10)) == 1
    src, dst = tmp_path / "src", tmp_path / "dst"
    src.mkdir()
    for fn in ("config.json", "tokenizer.json", "tokenizer_config.json", "generation_config.json", "vocab.txt", "modeling_custom.py"):
        (src / fn).write_text("{}")
    for fn in ("model-00001-of-00002.safetensors", "pytorch_model.bin", "model.safetensors.index.json", "optimizer.pt"):
        (src / fn).write_text("x")
    copied = saveload_utils.copy_hf_configs(str(src), str(dst))
    assert sorted(copied) == sorted(["config.json", "tokenizer.json", "tokenizer_config.json", "generation_config.json", "vocab.txt",
                                     "modeling_custom.py"])
    from safetensors.torch import save_file
    save_file({"a": torch.arange(4.0)}, str(tmp_path / "f.safetensors"))
    torch.save({"b": torch.ones(2)}, str(tmp_path / "f.bin"))
    assert saveload_utils.load_weight_file(str(tmp_path / "f.safetensors"))["a"].tolist() == [0, 1, 2, 3]
    assert saveload_utils.load_safetensor(str(tmp_path / "f.bin"))["b"].tolist() == [1, 1]


def test_gpu_utils_local_index_by_host_rendezvous(monkeypatch):
    """Workers of one host find their local GPU id as their position among the workers that published the same host name."""
    import threading
    import uuid

    from realhf_b200.base import gpu_utils, name_resolve
    name_resolve.reconfigure("memory")
    exp, trial = "gpuid-" + uuid.uuid4().hex[:6], "t"
    hosts = ["nodeA", "nodeA", "nodeB", "nodeA", "nodeB", "nodeB"]
    out = [None] * len(hosts)

    def run(i):
        out[i] = gpu_utils.local_gpu_index(exp, trial, "model_worker", i, len(hosts), host=hosts[i], n_gpus=8, timeout=20)
    ts = [threading.Thread(target=run, args=(i,)) for i in reversed(range(len(hosts)))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert out == [0, 1, 0, 2, 1, 2]
    with pytest.raises(RuntimeError):
        gpu_utils.local_gpu_index(exp, trial, "model_worker", 3, len(hosts), host="nodeA", n_gpus=2, timeout=5)
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "4,5,6")
    assert gpu_utils.gpu_count() == 3
    monkeypatch.setenv("REAL_LOCAL_GPU", "2")
    monkeypatch.setenv("REAL_ISOLATE_GPUS", "1")
    assert gpu_utils.isolate_cuda_device(exp, trial, "model_worker", 0, 1) == 0
    import os
    assert os.environ["CUDA_VISIBLE_DEVICES"] == "6" and os.environ["REAL_LOCAL_GPU"] == "0"


def test_asyncio_gather_or_raise_cancels_the_survivors():
    import asyncio

    from realhf_b200.base import asyncio_utils
    state = {"cancelled": 0, "finished": 0}

    async def sleeper(t):
        try:
            await asyncio.sleep(t)
            state["finished"] += 1
            return t
        except asyncio.CancelledError:
            state["cancelled"] += 1
            raise

    async def failing():
        await asyncio.sleep(0.05)
        raise ValueError("mfc failed")

    assert asyncio.run(asyncio_utils.gather_or_raise([sleeper(0.01), sleeper(0.02)])) == [0.01, 0.02]
    with pytest.raises(ValueError):
        asyncio.run(asyncio_utils.gather_or_raise([sleeper(5), failing(), sleeper(5)]))
    assert state["cancelled"] == 2 and state["finished"] == 2


def test_padding_roundtrips():
    import torch

    from realhf_b200.utils import padding
    torch.manual_seed(0)
    B, S, H = 4, 7, 3
    lens = torch.tensor([7, 2, 5, 1])
    x = torch.randn(B, S, H, requires_grad=True)
    # right padding
    mask = torch.arange(S)[None] < lens[:, None]
    packed, idx, cu, mx = padding.unpad_input(x, mask)
    assert packed.shape == (15, H) and cu.tolist() == [0, 7, 9, 14, 15] and mx == 7 and cu.dtype == torch.int32
    back = padding.pad_input(packed, idx, B, S)
    assert torch.equal(back, x * mask[..., None])
    back.sum().backward()
    assert torch.equal(x.grad, mask[..., None].expand_as(x).float())      # gradients reach exactly the real tokens
    # left padding and a hole in the middle keep token order
    lmask = torch.arange(S)[None] >= (S - lens)[:, None]
    lmask[0, 3] = False
    p2, idx2, cu2, _ = padding.unpad_input(x.detach(), lmask)
    assert cu2.tolist() == [0, 6, 8, 13, 14] and torch.equal(p2[:3], x.detach()[0, :3]) and torch.equal(p2[3:6], x.detach()[0, 4:])
    # packed <-> padded by lengths
    pk, cu3 = padding.pack_padded(x.detach(), lens)
    assert torch.equal(pk, packed.detach()) and cu3.tolist() == cu.tolist()
    right = padding.pad_packed(pk, cu3)
    left = padding.pad_packed(pk, cu3, seqlen=9, left=True, value=-1.0)
    assert torch.equal(right, x.detach() * mask[..., None]) and left.shape == (B, 9, H)
    assert torch.equal(left[1, -2:], x.detach()[1, :2]) and (left[1, :-2] == -1).all()
    # sequence-parallel padding of the packed token axis
    ids = torch.arange(15)
    ids2, cu4, mx4, n_pad = padding.pad_sequence_parallel_input(ids, cu, 7, tp=4, pad_id=0)
    assert ids2.numel() == 16 and n_pad == 1 and cu4.tolist() == [0, 7, 9, 14, 15, 16] and mx4 == 7
    assert padding.pad_sequence_parallel_input(ids2, cu4, 7, tp=4)[3] == 0


def test_numpy_ray_and_slurm_helpers():
    import numpy as np

    from realhf_b200.base import numpy_utils, ray_utils, slurm_utils
    assert numpy_utils.shape_leq((2, 3), (2, 4)) and not numpy_utils.shape_leq((2, 5), (2, 4)) and not numpy_utils.shape_leq((2,), (2, 4))
    assert numpy_utils.shape_union((2, 3), (1, 7), (4, 1)) == (4, 7)
    a, b = np.arange(2 * 3).reshape(2, 3), np.arange(2 * 2 * 2).reshape(2, 2, 2) + 100
    packed = np.concatenate([a, b.reshape(2, 4)], axis=-1)
    back = numpy_utils.split_to_shapes(packed, {"a": (2, 3), "b": (2, 2, 2)}, axis=1)
    assert np.array_equal(back["a"], a) and np.array_equal(back["b"], b)
    with pytest.raises(ValueError):
        numpy_utils.split_to_shapes(packed, {"a": (2, 3)}, axis=1)
    assert isinstance(ray_utils.check_ray_availability(), bool) and isinstance(slurm_utils.check_slurm_availability(), bool)
    nodes = slurm_utils.parse_nodelist("NODE[01-03,07],NODE10", "NODE")
    assert nodes == ["NODE01", "NODE02", "NODE03", "NODE07", "NODE10"]
    assert slurm_utils.nodelist_from_nodes(nodes, "NODE") == "NODE[01-03,07,10]"
    assert slurm_utils.nodelist_from_nodes(["NODE05"], "NODE") == "NODE05"
    assert slurm_utils.parse_nodelist(slurm_utils.nodelist_from_nodes(nodes, "NODE"), "NODE") == nodes
    with pytest.raises(ValueError):
        slurm_utils.parse_nodelist("gpu[1-2]", "NODE")
    assert slurm_utils.are_ones_contiguous(np.array([0, 1, 1, 0])) and not slurm_utils.are_ones_contiguous(np.array([1, 0, 1]))
    assert sorted(["node10", "node2", "node1"], key=slurm_utils.slurm_hostname_key) == ["node1", "node2", "node10"]
    assert slurm_utils.parse_node_id("NODE07", "NODE") == 7
