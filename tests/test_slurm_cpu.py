"""Slurm plumbing offline: hostlist syntax, node inventory parsing, resource allocation of job steps, generated hostfile /
multi-prog / sbatch files, job-state folding, and the scheduler client driven by canned `sbatch` / `scontrol` / `squeue` output."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realhf_b200.scheduler import slurm as S  # noqa: E402

SCONTROL = """\
NodeName=b200-01 Arch=x86_64 CPUAlloc=0 CPUTot=224 RealMemory=2000000 AllocMem=0 Gres=gpu:b200:8(S:0-1) State=IDLE Partitions=dev AllocTRES=
NodeName=b200-02 Arch=x86_64 CPUAlloc=64 CPUTot=224 RealMemory=2000000 AllocMem=500000 Gres=gpu:b200:8(S:0-1) State=MIXED Partitions=dev AllocTRES=cpu=64,mem=500000M,gres/gpu=4,gres/gpu:b200=4
NodeName=b200-03 Arch=x86_64 CPUAlloc=0 CPUTot=224 RealMemory=2000000 AllocMem=0 Gres=gpu:b200:8 State=IDLE+DRAIN Partitions=dev AllocTRES=
NodeName=cpu-01 Arch=x86_64 CPUAlloc=8 CPUTot=128 RealMemory=512000 AllocMem=32000 Gres=(null) State=MIXED Partitions=dev AllocTRES=cpu=8,mem=32000M
"""


def test_hostlist_roundtrip():
    hosts = S.parse_nodelist("b200-[01-03,07],cpu-01,x[9-11]")
    assert hosts == ["b200-01", "b200-02", "b200-03", "b200-07", "cpu-01", "x9", "x10", "x11"]
    assert S.parse_nodelist(S.compress_nodelist(hosts)) == sorted(hosts, key=lambda h: (h.rstrip("0123456789"), len(h), h)) or \
        sorted(S.parse_nodelist(S.compress_nodelist(hosts))) == sorted(hosts)
    assert S.compress_nodelist(["n01", "n02", "n04"]) == "n[01-02,04]"
    assert S.parse_nodelist(None) == [] and S.parse_nodelist("single") == ["single"]


def test_node_inventory_reports_free_resources_and_skips_drained_nodes():
    nodes = S.parse_scontrol_nodes(SCONTROL)
    assert set(nodes) == {"b200-01", "b200-02", "cpu-01"}
    assert nodes["b200-01"] == S.SlurmResource(224, 2000000, 8, "b200")
    assert nodes["b200-02"] == S.SlurmResource(160, 1500000, 4, "b200")
    assert nodes["cpu-01"].gpu == 0 and nodes["cpu-01"].cpu == 120


def _infos(tmp_path, n_model=12):
    common = dict(run_name="exp_t0", partition="dev", log_dir=str(tmp_path), container_image="img:latest", container_mounts="/data:/data")
    mw = S.SlurmLaunchInfo(worker_type="model_worker", cmd="python -m w -i {jobstep_id} -g {n_jobsteps} -p {wprocs_per_jobstep} -j {wprocs_in_job} -o {wproc_offset}",
                           wprocs_in_job=n_model, resource=S.SlurmResource(cpu=16, mem=100000, gpu=1, gpu_type="b200"), **common)
    master = S.SlurmLaunchInfo(worker_type="master_worker", cmd="python -m m -i {jobstep_id}", wprocs_in_job=1,
                               resource=S.SlurmResource(cpu=8, mem=32000), **common)
    return mw, master


def test_allocation_packs_gpu_steps_in_rank_order_and_respects_free_gpus(tmp_path):
    mw, master = _infos(tmp_path)
    S.allocate_resources([master, mw], S.parse_scontrol_nodes(SCONTROL))
    # 8 free GPUs on b200-01, then the 4 free ones on b200-02: consecutive ranks share a node (one NVSwitch domain)
    assert mw.hosts == ["b200-01"] * 8 + ["b200-02"] * 4
    assert master.hosts and master.hosts[0] in ("b200-01", "b200-02", "cpu-01")
    with pytest.raises(S.SlurmResourceNotEnoughException):
        S.allocate_resources(list(_infos(tmp_path, n_model=13)), S.parse_scontrol_nodes(SCONTROL))
    mw2, _ = _infos(tmp_path, n_model=4)
    mw2.exclude = "b200-01"
    S.allocate_resources([mw2], S.parse_scontrol_nodes(SCONTROL))
    assert mw2.hosts == ["b200-02"] * 4
    mw3, _ = _infos(tmp_path, n_model=2)
    mw3.resource = S.SlurmResource(cpu=16, mem=1000, gpu=1, gpu_type="h100")
    with pytest.raises(S.SlurmResourceNotEnoughException):
        S.allocate_resources([mw3], S.parse_scontrol_nodes(SCONTROL))


def test_generated_files(tmp_path):
    mw, master = _infos(tmp_path, n_model=10)
    mw.wprocs_per_jobstep = 1
    mw.env_vars = {"REAL_MODE": "SLURM", "X": "a b"}
    mw.time_limit = "2:00:00"
    S.allocate_resources([mw, master], S.parse_scontrol_nodes(SCONTROL))
    path = mw.commit()
    script = open(path).read()
    assert "#SBATCH --ntasks=10" in script and "#SBATCH --nodes=2" in script and "#SBATCH --nodelist=b200-[01-02]" in script
    assert "#SBATCH --gpus-per-task=b200:1" in script and "#SBATCH --distribution=arbitrary" in script and "#SBATCH --time=2:00:00" in script
    assert "export X='a b'" in script and "--container-image=img:latest --container-mounts=/data:/data" in script
    assert "--multi-prog" in script and "SLURM_HOSTFILE" in script
    assert open(tmp_path / "model_worker.hostfile").read().split() == ["b200-01"] * 8 + ["b200-02"] * 2
    mp = open(tmp_path / "model_worker.multiprog").read().strip().splitlines()
    assert len(mp) == 10 and mp[3] == "3 python -m w -i 3 -g 10 -p 1 -j 10 -o 0"
    # several worker processes per job step (CPU-side workers): the last step takes the remainder
    cw = S.SlurmLaunchInfo(run_name="exp_t0", worker_type="data_worker", cmd="w -i {jobstep_id} -p {wprocs_per_jobstep} -j {wprocs_in_job}",
                           wprocs_in_job=10, wprocs_per_jobstep=4, resource=S.SlurmResource(cpu=4, mem=1000), log_dir=str(tmp_path))
    assert cw.n_jobsteps == 3
    assert cw.multiprog().strip().splitlines() == ["0 w -i 0 -p 4 -j 10", "1 w -i 1 -p 4 -j 10", "2 w -i 2 -p 2 -j 10"]


def test_job_state_folding():
    st = S.parse_job_states("101|RUNNING|exp_t0:model_worker|b200-[01-02]\n102|PENDING|exp_t0:master_worker|\n")
    assert st["101"][0] == "RUNNING" and st["102"][0] == "PENDING"
    acct = S.parse_job_states("201|COMPLETED|a|n1\n201.batch|COMPLETED|batch|n1\n201.0|FAILED|w|n1\n202|CANCELLED by 1000|b|n2\n203|OUT_OF_MEMORY|c|n3\n")
    assert acct["201"][0] == "FAILED" and acct["202"][0] == "CANCELLED" and acct["203"][0] == "FAILED"
    calls = []

    def fake(argv):
        calls.append(argv[0])
        if argv[0] == "squeue":
            return "301|RUNNING|x|n1\n"
        return "302|COMPLETED|y|n2\n302.0|COMPLETED|y|n2\n"
    res = S.job_states(["301", "302", "303"], run=fake)
    assert res["301"][0] == "RUNNING" and res["302"][0] == "COMPLETED" and res["303"][0] == "COMPLETED" and calls == ["squeue", "sacct"]


def test_scheduler_client_allocates_then_submits(tmp_path, monkeypatch):
    monkeypatch.setenv("REAL_FILEROOT", str(tmp_path))
    from realhf_b200.scheduler import client as C
    submitted = []

    def fake(argv):
        if argv[:3] == ["scontrol", "show", "nodes"]:
            return SCONTROL
        if argv[0] == "sinfo":
            return "b200-[01-03],cpu-01\n"
        if argv[0] == "sbatch":
            submitted.append(open(argv[-1]).read())
            return f"{400 + len(submitted)};cluster\n"
        if argv[0] == "squeue":
            return "".join(f"{400 + i + 1}|{'RUNNING' if len(submitted) and fake.polls < 1 else 'COMPLETED'}|j|n\n" for i in range(len(submitted)))
        if argv[0] == "sacct":
            return ""
        raise AssertionError(argv)
    fake.polls = 0
    sc = C.SlurmSchedulerClient("exp", "t0", partition="dev", run=fake)
    sc.submit_array("model_worker", C.remote_worker_cmd("exp", "t0", False, "model_worker"), count=10, cpu=16, gpu=1, mem=100000,
                    env_vars={"REAL_MODE": "SLURM"})
    sc.submit("master_worker", C.remote_worker_cmd("exp", "t0", False, "master_worker"), cpu=8, mem=32000)
    assert not submitted, "jobs are placed together and submitted on wait()/commit()"
    sc.commit()
    assert len(submitted) == 2
    mw = next(s for s in submitted if ":model_worker" in s)
    assert "#SBATCH --nodelist=b200-[01-02]" in mw and "#SBATCH --ntasks=10" in mw
    infos = sc.find_all()
    assert {i.name for i in infos} == {"model_worker", "master_worker"}
    fake.polls = 1
    sc.wait(timeout=5, poll=0.01)
