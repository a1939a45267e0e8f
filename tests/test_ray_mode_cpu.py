"""Ray mode (`mode=ray`): the scheduler client that runs workers as Ray tasks, rank-ordered GPU packing, and the queue
transport of the worker control plane.  This image has no Ray: `tests/fake_ray` provides the slice of its API that the
client uses, with tasks as spawned subprocesses, so the whole launcher -> master + model workers path runs end to end.
Reference: realhf/system/controller.py:348-575, system/worker_control.py:48-150."""
import json
import os
import queue
import sys
import threading
import time
import uuid

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAKE = os.path.join(ROOT, "tests", "fake_ray")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fixtures  # noqa: E402


@pytest.fixture()
def fake_ray(monkeypatch):
    try:
        import ray
        if "fake_ray" not in (ray.__file__ or ""):
            pytest.skip("a real Ray is installed: these tests drive the stand-in only")
    except ImportError:
        pass
    monkeypatch.syspath_prepend(FAKE)
    monkeypatch.setenv("PYTHONPATH", FAKE + os.pathsep + ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    for m in [m for m in sys.modules if m == "ray" or m.startswith("ray.")]:
        del sys.modules[m]
    import ray
    yield ray
    ray.shutdown()
    for m in [m for m in sys.modules if m == "ray" or m.startswith("ray.")]:
        del sys.modules[m]


def test_gpu_workers_are_packed_in_rank_order():
    from realhf_b200.scheduler.ray import pack_gpu_workers, required_resources
    nodes = ["node:10.0.0.1", "node:10.0.0.2"]
    place = pack_gpu_workers(12, nodes, 16.0)
    assert place[:8] == [("node:10.0.0.1", i) for i in range(8)]
    assert place[8:] == [("node:10.0.0.2", i) for i in range(4)]
    with pytest.raises(RuntimeError, match="GPU workers requested"):
        pack_gpu_workers(17, nodes, 16.0)
    with pytest.raises(ValueError, match="different numbers"):
        pack_gpu_workers(4, nodes, 15.0)
    with pytest.raises(RuntimeError, match="no `node:"):
        pack_gpu_workers(1, [], 8.0)
    need = required_resources([dict(cpu=4, gpu=1, mem=2048, count=8), dict(cpu=2, gpu=0, mem=1024, count=1)])
    assert need == {"CPU": 34.0, "GPU": 8.0, "memory": 17.0}


def _ok_task(worker_type, index, world, exp, trial, env, slot):
    return (worker_type, index, os.environ.get("MARK"), os.environ.get("FAKE_RAY_GPU_IDS"), slot)


def _bad_task(worker_type, index, world, exp, trial, env, slot):
    raise ValueError("boom")


def _slow_task(worker_type, index, world, exp, trial, env, slot):
    time.sleep(60)


def test_client_submits_finds_fails_and_cancels(fake_ray, monkeypatch):
    import realhf_b200.scheduler.ray as R
    from realhf_b200.scheduler.client import JobException, JobState, make
    monkeypatch.setenv("FAKE_RAY_RESOURCES", json.dumps({"CPU": 32, "GPU": 4, "memory": 64 * 1024 ** 3, "node:10.0.0.1": 1.0,
                                                         "node:10.0.0.2": 1.0}))
    monkeypatch.setattr(R, "run_ray_worker", _ok_task)
    c = make("ray", "exp", "trial")
    assert isinstance(c, R.RaySchedulerClient) and fake_ray.is_initialized()
    c.submit_array("model_worker", "unused", count=4, cpu=2, gpu=1, mem=1024, env_vars={"MARK": "m1"})
    c.submit_array("master_worker", "unused", count=1, cpu=1, gpu=0, mem=512, env_vars={"MARK": "m2"})
    c.wait(timeout=60)
    assert {i.name: i.state for i in c.find_all()} == {**{f"model_worker/{i}": JobState.COMPLETED for i in range(4)},
                                                      "master_worker/0": JobState.COMPLETED}
    assert [i.name for i in c.find_all("master_worker.*")] == ["master_worker/0"]
    launched = {r["opts"]["name"]: r for r in fake_ray._STATE["launched"]}
    # two GPUs per node: ranks 0,1 on the first node, 2,3 on the second, each holding half of its node resource
    assert [next(iter(launched[f"model_worker/{i}"]["opts"]["resources"])) for i in range(4)] == ["node:10.0.0.1"] * 2 + ["node:10.0.0.2"] * 2
    assert all(launched[f"model_worker/{i}"]["opts"]["resources"] == {f"node:10.0.0.{1 + i // 2}": 0.5} for i in range(4))
    assert "resources" not in launched["master_worker/0"]["opts"]
    results = {n: fake_ray.get(r["ref"]) for n, r in launched.items()}
    assert results["model_worker/3"] == ("model_worker", 3, "m1", "1", 1) and results["master_worker/0"][2:] == ("m2", "", None)
    assert c.find("nobody/0").state == JobState.NOT_FOUND
    c.stop_all()

    # a failing task: FAILED, and wait() raises the scheduler's exception
    monkeypatch.setattr(R, "run_ray_worker", _bad_task)
    c = make("ray", "exp", "trial2")
    try:   # a task that dies quickly surfaces at launch (the first poll), like an ImportError on the remote side would ...
        c.submit_array("master_worker", "unused", count=1)
        with pytest.raises(JobException) as ei:   # ... and one that dies later as a FAILED job
            c.wait(timeout=60)
        assert ei.value.reason == JobState.FAILED and c.find("master_worker/0").state == JobState.FAILED
    except fake_ray.exceptions.RayTaskError as e:
        assert "boom" in str(e)
    c.stop_all()
    monkeypatch.setattr(R, "run_ray_worker", _slow_task)
    c = make("ray", "exp", "trial3")
    c.submit_array("master_worker", "unused", count=2)
    assert [i.state for i in c.find_all()] == [JobState.RUNNING] * 2
    with pytest.raises(TimeoutError):
        c.wait(timeout=0.5)
    refs = dict(c._refs)
    c.stop_all()
    assert all(not r._proc.is_alive() for r in refs.values())
    assert c.find_all() == []


def test_short_resources_are_reported(fake_ray, monkeypatch, caplog):
    import realhf_b200.scheduler.ray as R
    from realhf_b200.scheduler.client import make
    monkeypatch.setenv("FAKE_RAY_RESOURCES", json.dumps({"CPU": 1, "GPU": 0, "memory": 1024 ** 3, "node:10.0.0.1": 1.0}))
    monkeypatch.setattr(R, "run_ray_worker", _ok_task)
    c = make("ray", "exp", "trial")
    msgs = []
    monkeypatch.setattr(R.logger, "critical", lambda m, *a: msgs.append(m))
    c.submit_array("master_worker", "unused", count=2, cpu=4, mem=512)
    assert msgs and "CPU: need 8" in msgs[0]
    with pytest.raises(RuntimeError, match="GPU workers requested"):
        c.submit_array("model_worker", "unused", count=1, gpu=1)
    c.stop_all()


def test_queue_transport_of_the_worker_control_plane(tmp_path):
    from realhf_b200.base import name_resolve
    from realhf_b200.system.worker_control import WorkerControlPanel, WorkerException, WorkerServer, WorkerServerStatus
    name_resolve.reconfigure("nfs", record_root=str(tmp_path))
    try:
        comms = {f"model_worker/{i}": (queue.Queue(8), queue.Queue(8)) for i in range(2)}
        servers = {n: WorkerServer("e", "t", n, comm=c) for n, c in comms.items()}
        servers["model_worker/1"].register_handler("slow", lambda: time.sleep(0.4) or "late")
        servers["model_worker/1"].register_handler("mul", lambda a, b: a * b)
        panel = WorkerControlPanel("e", "t", timeout=5)
        for n, (rq, pq) in comms.items():
            panel.attach_queues(n, rq, pq)
        assert panel.worker_names == sorted(comms)
        assert panel.request("model_worker/0", "ping") == "pong"
        assert panel.group_request("status")["model_worker/1"]["status"] == "READY"
        assert panel.request("model_worker/1", "mul", a=6, b=7) == 42
        with pytest.raises(RuntimeError, match="no handler"):
            panel.request("model_worker/0", "mul", a=1, b=2)
        # a reply that misses its deadline is reported as LOST and is NOT handed to the next request
        with pytest.raises(WorkerException):
            panel.request("model_worker/1", "slow", timeout=0.05)
        time.sleep(0.6)
        assert panel.request("model_worker/1", "mul", a=2, b=3) == 6
        # pause / resume / exit flags reach a main loop exactly as with the ZMQ transport
        srv = servers["model_worker/0"]
        srv.set_status(WorkerServerStatus.RUNNING)
        steps = []
        th = threading.Thread(target=lambda: [steps.append(1) or time.sleep(0.01) for _ in iter(lambda: srv.wait_while_paused(poll=0.01), False)])
        th.start()
        panel.request("model_worker/0", "pause")
        time.sleep(0.1)
        n = len(steps)
        time.sleep(0.1)
        assert len(steps) == n and panel.pulse()["model_worker/0"] == WorkerServerStatus.PAUSED
        panel.request("model_worker/0", "exit")
        th.join(timeout=5)
        assert not th.is_alive()
        # queue-transport workers publish no routable address: connecting by name says how to reach them instead
        other = WorkerControlPanel("e", "t")
        with pytest.raises(ValueError, match="attach_queues"):
            other.connect(["model_worker/0"])
        for s in servers.values():
            s.close()
    finally:
        name_resolve.reconfigure("nfs")


@pytest.mark.distributed
def test_sft_experiment_runs_in_ray_mode(tmp_path, fake_ray):
    os.environ["PYTHONPATH"] = FAKE + os.pathsep + ROOT + os.pathsep + os.environ.get("PYTHONPATH", "")
    os.environ["REAL_FILEROOT"] = str(tmp_path / "fileroot")
    os.environ["REAL_NAME_RESOLVE_ROOT"] = str(tmp_path / "nr")
    import importlib

    from realhf_b200.base import constants, name_resolve
    importlib.reload(constants)
    name_resolve.reconfigure("nfs", record_root=str(tmp_path / "nr"))
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    ckpt = str(tmp_path / "gpt2")
    cfg, tok, words = fixtures.make_checkpoint(ckpt, "gpt2")
    data = str(tmp_path / "sft.jsonl")
    fixtures.write_sft_dataset(data, words, n=32)
    exp = build_experiment([
        "sft", f"experiment_name=sftray-{uuid.uuid4().hex[:6]}", "trial_name=t0", "mode=ray", "device=cpu", "dtype=fp32", "n_nodes=1",
        "n_gpus_per_node=2", "allocation_mode=manual", "allocation.parallel.data_parallel_size=2", "model.type._class=gpt2",
        f"model.path={ckpt}", f"dataset.train_path={data}", "dataset.train_bs_n_seqs=16", "dataset.max_seqlen=64",
        "exp_ctrl.total_train_epochs=2", "model.optimizer.lr=1e-3", "model.optimizer.warmup_steps_proportion=0.0",
        "model.optimizer.grad_dtype=fp32", "model.gradient_checkpointing=false"])
    try:
        main_start(exp, timeout=600)
    finally:
        name_resolve.reconfigure("nfs")
    log_dir = os.path.join(os.environ["REAL_FILEROOT"], "logs", exp.experiment_name, "t0")
    log = open(os.path.join(log_dir, "master_worker-0")).read()       # the task wrote the log file a process would have
    losses = [float(l.split("loss=")[1].split(",")[0]) for l in log.splitlines() if "[trainDefault]" in l and "loss=" in l]
    assert len(losses) == 4 and losses[-1] < losses[0], log[-3000:]
    assert os.path.getsize(os.path.join(log_dir, "model_worker-1")) > 0
    names = sorted(r["opts"]["name"] for r in fake_ray._STATE["launched"])
    assert names == ["master_worker/0", "model_worker/0", "model_worker/1"]
