"""Worker control plane: request / response endpoints, group requests, pause / resume / exit flags, lost workers."""
import os
import sys
import threading
import time

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


@pytest.fixture()
def store(tmp_path, monkeypatch):
    from realhf_b200.base import name_resolve
    name_resolve.reconfigure("nfs", record_root=str(tmp_path))
    yield name_resolve
    name_resolve.reconfigure("nfs")


def test_panel_talks_to_servers_and_flags_reach_the_main_loop(store):
    from realhf_b200.system.worker_control import WorkerControlPanel, WorkerServer, WorkerServerStatus
    servers = [WorkerServer("e", "t", f"model_worker/{i}", host="127.0.0.1") for i in range(3)]
    master = WorkerServer("e", "t", "master_worker/0", host="127.0.0.1")
    state = {"step": 0}
    master.register_handler("progress", lambda: dict(step=state["step"]))
    master.register_handler("add", lambda a, b=1: a + b)
    for s in servers + [master]:
        s.set_status(WorkerServerStatus.RUNNING)
    panel = WorkerControlPanel("e", "t", timeout=5)
    assert panel.discover() == ["master_worker/0", "model_worker/0", "model_worker/1", "model_worker/2"]
    panel.connect()
    assert panel.request("master_worker/0", "ping") == "pong"
    assert panel.request("master_worker/0", "add", a=2, b=5) == 7
    assert set(panel.pulse().values()) == {WorkerServerStatus.RUNNING}
    res = panel.group_request("add", worker_names=["master_worker/0"], worker_kwargs={"master_worker/0": dict(a=10)})
    assert res == {"master_worker/0": 11}
    with pytest.raises(RuntimeError, match="no handler"):
        panel.request("model_worker/1", "nonsense")

    # a main loop that honours pause / resume / exit between its steps
    def loop():
        while master.wait_while_paused(poll=0.01):
            state["step"] += 1
            time.sleep(0.01)
    th = threading.Thread(target=loop)
    th.start()
    time.sleep(0.1)
    assert panel.request("master_worker/0", "pause") == "pausing"
    time.sleep(0.1)
    frozen = panel.request("master_worker/0", "progress")["step"]
    time.sleep(0.15)
    assert panel.request("master_worker/0", "progress")["step"] == frozen
    assert panel.pulse()["master_worker/0"] == WorkerServerStatus.PAUSED
    panel.request("master_worker/0", "resume")
    time.sleep(0.1)
    assert panel.request("master_worker/0", "progress")["step"] > frozen
    panel.request("master_worker/0", "exit")
    th.join(timeout=5)
    assert not th.is_alive()
    # a dead worker is reported LOST and does not wedge the panel
    servers[2].close()
    st = panel.group_request("status", timeout=0.5)
    assert isinstance(st["model_worker/2"], Exception) and isinstance(st["model_worker/0"], dict)
    assert panel.pulse()["model_worker/0"] == WorkerServerStatus.RUNNING
    panel.close()
    for s in servers[:2] + [master]:
        s.close()
