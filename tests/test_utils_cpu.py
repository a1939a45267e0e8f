"""Utilities and aux subsystems: logits warpers, stats tracker, in-flight batching, profile experiment, trace summary."""
import json

import pytest
import os

import torch

from realhf_b200.api.model import GenerationHyperparameters
from realhf_b200.base import monitor
from realhf_b200.models import generation as gen
from realhf_b200.models import hf_io
from realhf_b200.models.real_model import ReaLModel
from realhf_b200.utils import logits_warper as LW
from realhf_b200.utils.stats_tracker import StatsTracker


def test_logits_warpers_match_generation_filter():
    torch.manual_seed(0)
    x = torch.randn(6, 200) * 3
    g = GenerationHyperparameters(top_k=20, top_p=0.7, temperature=0.9)
    ref = gen._filter_logits((x / g.temperature).clone(), g)
    out = LW.chained_logits_wraper([LW.TemperatureLogitsWarper(0.9), LW.TopKLogitsWarper(20), LW.TopPLogitsWarper(0.7)])(None, x)
    kept_ref = ref > torch.finfo(ref.dtype).min
    kept = out > -float("inf")
    assert torch.equal(kept, kept_ref)
    assert (kept.sum(-1) >= 1).all() and (kept.sum(-1) <= 20).all()
    eps = LW.EpsilonLogitsWarper(0.01)(None, x)
    assert ((eps > -float("inf")).sum(-1) >= 1).all()


def test_stats_tracker_reduce_hook():
    t = StatsTracker()
    t.log("aux", torch.tensor(1.0))
    t.log("aux", torch.tensor(2.5), hook=lambda v: v * 2)
    assert t.pop("aux").item() == 7.0
    assert t.pop("aux") is None


def test_inflight_batching_matches_batched_generation():
    cfg = hf_io.family("llama").make_test_config()
    m = ReaLModel(cfg, dtype=torch.float32).instantiate(seed=3)
    g = GenerationHyperparameters(max_new_tokens=6, min_new_tokens=6, greedy=True)
    torch.manual_seed(0)
    prompts = [torch.randint(2, 128, (n,)) for n in (5, 9, 4, 7, 6)]
    res = gen.InflightBatchingGenerator(m, g, eos_id=1, pad_id=0, n_slots=2, max_prompt_len=16).generate(prompts)
    ids = torch.cat(prompts)
    cu = torch.zeros(len(prompts) + 1, dtype=torch.int32)
    cu[1:] = torch.tensor([p.numel() for p in prompts]).cumsum(0)
    out, _ = gen.generate(m, ids, cu, g, 1, 0)
    for i, (toks, lps, _ended) in enumerate(res):
        assert toks == out.tokens[i].tolist()
        torch.testing.assert_close(torch.tensor(lps), out.logprobs[i], atol=1e-4, rtol=1e-3)


def test_profile_experiment_cpu(tmp_path):
    from realhf_b200.apps.quickstart import build_experiment
    out = tmp_path / "prof.json"
    cfg = build_experiment(["profile", "device=cpu", "batch_sizes=[2]", "seqlens=[12]", "handles=[inference,train_step]", "repeats=1",
                            f"output_file={out}"])
    rows = cfg.run_local()
    assert {r["handle"] for r in rows} == {"inference", "train_step"} and all(r["secs"] > 0 for r in rows)
    assert json.load(open(out))[0]["bs"] == 2
    # the file is what `REAL_MFC_PROFILE` feeds to the allocation search: an exact (handle, layout, batch, length) hit returns its time
    import os

    from realhf_b200.api.config import ModelInterfaceAbstraction, ModelInterfaceType
    from realhf_b200.api.dfg import MFCDef
    from realhf_b200.api.quickstart import ParallelismConfig
    from realhf_b200.search.engine import MFCProfile
    os.environ["REAL_MFC_PROFILE"] = str(out)
    try:
        prof = MFCProfile.find()
    finally:
        del os.environ["REAL_MFC_PROFILE"]
    rpc = MFCDef("inf", 2, ModelInterfaceType.INFERENCE, ModelInterfaceAbstraction("ppo_actor"), "actor", input_keys=("packed_input_ids",),
                 output_keys=("x",))
    row = next(r for r in rows if r["handle"] == "inference")
    assert prof.time_us(rpc, ParallelismConfig(1, 1, 1), 4, 8) == pytest.approx(row["secs"] * 1e6)      # prompt 4 + generated 8 = 12 tokens
    assert prof.time_us(rpc, ParallelismConfig(1, 1, 2), 4, 8) is None                                  # another layout: no measurement


def test_profile_experiment_sweeps_layouts_in_one_launch(tmp_path):
    """`layouts=[...]` sweeps parallel layouts inside one launch: multi-device layouts run in their own process group (gloo here),
    a point's time is the maximum over its ranks, `bs` is the global batch."""
    from realhf_b200.apps.quickstart import build_experiment
    out = tmp_path / "prof.json"
    cfg = build_experiment(["profile", "device=cpu", "interface=sft", "batch_sizes=[4]", "seqlens=[12]", "handles=[train_step]", "repeats=1",
                            "layouts=[d1m1p1,d2m1p1,d1m2p1]", f"output_file={out}"])
    rows = cfg.run_local()
    assert [r["layout"] for r in rows] == ["d1m1p1", "d2m1p1", "d1m2p1"]
    assert all(r["secs"] > 0 and r["bs"] == 4 and r["handle"] == "train_step" for r in rows)
    assert len(json.load(open(out))) == 3


def test_kernel_trace_categorisation():
    ev = [dict(cat="kernel", name="ncclDevKernel_AllReduce_Sum_bf16", dur=10.0), dict(cat="kernel", name="gemm_2cta_kernel<256>", dur=30.0),
          dict(cat="gpu_memcpy", name="Memcpy DtoD", dur=5.0), dict(cat="cpu_op", name="aten::add", dur=99.0)]
    s = monitor.summarize_chrome_trace(ev)
    assert s == {"collective": 10.0, "compute": 30.0, "memory": 5.0}
    f = monitor.calculate_llama_forward_flops(2, [16, 16], 2, 64, 128, 100)
    assert monitor.calculate_llama_train_flops(3, 2, [16, 16], 2, 64, 128, 100) == 3 * f


def test_profile_experiment_mocks_cover_every_interface():
    """`quickstart profile` fabricates valid inputs for every built-in interface through their `_mock_*` hooks."""
    from realhf_b200.apps.quickstart import build_experiment
    for itf, handles, extra in (("sft", "[train_step]", []), ("paired_rw", "[inference,train_step]", ["model.type.is_critic=True"]),
                                ("dpo", "[inference,train_step]", []), ("ppo_critic", "[inference,train_step]", ["model.type.is_critic=True"]),
                                ("generation", "[generate]", ["gen.max_new_tokens=3", "gen.min_new_tokens=3"])):
        cfg = build_experiment(["profile", "device=cpu", "batch_sizes=[4]", "seqlens=[12]", f"handles={handles}", f"interface={itf}", "repeats=1"] + extra)
        rows = cfg.run_local()
        assert rows and all(r["secs"] > 0 for r in rows), (itf, rows)


def test_metric_sinks_jsonl_tensorboard_and_failing_sink(tmp_path, monkeypatch):
    import glob
    import json
    import sys
    import types
    from realhf_b200.system.metrics import MetricSinks
    monkeypatch.setenv("REAL_TENSORBOARD", "1")
    monkeypatch.setenv("WANDB_MODE", "offline")
    logged = []

    class _Run:
        def log(self, d, step):
            if step == 2:
                raise RuntimeError("quota")
            logged.append((step, d))

        def finish(self):
            logged.append("finished")
    monkeypatch.setitem(sys.modules, "wandb", types.SimpleNamespace(init=lambda **kw: _Run()))
    m = MetricSinks("exp", "trial", str(tmp_path))
    assert m.active == ["jsonl", "tensorboard", "wandb"]
    for step in (1, 2, 3):
        m.log({"rpc": "actor_train", "step": step, "epoch": 0, "time": 1.0 * step, "loss": 0.5 / step, "n_tokens": 100, "note": "x", "ok": True})
    assert m.active == ["jsonl", "tensorboard"]          # the failing sink was switched off, the others kept going
    m.close()
    rows = [json.loads(l) for l in open(tmp_path / "stats.jsonl")]
    assert [r["step"] for r in rows] == [1, 2, 3] and rows[0]["loss"] == 0.5
    assert logged[0] == (1, {"actor_train/loss": 0.5, "actor_train/n_tokens": 100.0, "epoch": 0})
    ev = glob.glob(str(tmp_path / "tensorboard" / "events.out.tfevents.*"))
    assert ev and os.path.getsize(ev[0]) > 0
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    acc = EventAccumulator(str(tmp_path / "tensorboard"))
    acc.Reload()
    assert [e.step for e in acc.Scalars("actor_train/loss")] == [1, 2, 3]


def test_group_padding_for_the_grouped_wgrad_kernel():
    """The dispatch-side contract of csrc/gemm_grouped_wgrad.cu: blocks start at multiples of 64, padding rows are zero, and
    the per-expert products over the PADDED blocks equal the products over the original ragged blocks."""
    import torch
    from realhf_b200.ops.gemm import grouped_wgrad_ref, pad_groups
    torch.manual_seed(0)
    for counts in ([5, 0, 64, 130, 1], [0, 0, 7], [128, 64], [1] * 9, [0, 0]):
        G, T = len(counts), sum(counts)
        offsets = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), dtype=torch.int32)
        dy, x = torch.randn(T, 16), torch.randn(T, 24)
        dest, off_pad, n_pad = pad_groups(offsets, T)
        assert n_pad % 64 == 0 and int(off_pad[-1]) <= n_pad and all(int(o) % 64 == 0 for o in off_pad)
        assert [int(b - a) for a, b in zip(off_pad[:-1], off_pad[1:])] == [(c + 63) // 64 * 64 for c in counts]
        assert dest.numel() == T and (T == 0 or dest.unique().numel() == T)
        dy_p = torch.zeros(n_pad, 16).index_copy_(0, dest, dy)
        x_p = torch.zeros(n_pad, 24).index_copy_(0, dest, x)
        for g in range(G):   # rows land inside their own block, in order
            a, b = int(offsets[g]), int(offsets[g + 1])
            torch.testing.assert_close(dy_p[int(off_pad[g]): int(off_pad[g]) + (b - a)], dy[a:b])
            assert float(dy_p[int(off_pad[g]) + (b - a): int(off_pad[g + 1])].abs().sum()) == 0.0
        got = torch.stack([dy_p[int(off_pad[g]): int(off_pad[g + 1])].t() @ x_p[int(off_pad[g]): int(off_pad[g + 1])] for g in range(G)])
        torch.testing.assert_close(got, grouped_wgrad_ref(dy, x, offsets, G), atol=1e-4, rtol=1e-4)


def test_group_padding_drops_rows_that_belong_to_no_group():
    """An expert-parallel rank hands the grouped wgrad the offsets of ITS experts inside the globally sorted rows (offsets[0] > 0,
    rows of foreign experts before and after), and a receive buffer has an unused tail: those rows must land in the spare row
    past the padded tensor, never inside a group's zero padding."""
    import torch
    from realhf_b200.ops.gemm import grouped_wgrad_ref, pad_groups
    torch.manual_seed(1)
    T = 300
    offsets = torch.tensor([37, 37 + 70, 37 + 70, 37 + 70 + 5], dtype=torch.int32)   # 3 groups inside rows [37, 112)
    dy, x = torch.randn(T, 8), torch.randn(T, 12)
    dest, off_pad, n_pad = pad_groups(offsets, T)
    outside = torch.cat([torch.arange(0, 37), torch.arange(112, T)])
    assert (dest[outside] == n_pad).all() and int(dest.min()) >= 0
    inside = dest[37:112]
    assert inside.unique().numel() == 75 and int(inside.max()) < int(off_pad[-1])
    dy_p = torch.zeros(n_pad + 1, 8).index_copy_(0, dest, dy)[:n_pad]
    x_p = torch.zeros(n_pad + 1, 12).index_copy_(0, dest, x)[:n_pad]
    got = torch.stack([dy_p[int(off_pad[g]): int(off_pad[g + 1])].t() @ x_p[int(off_pad[g]): int(off_pad[g + 1])] for g in range(3)])
    off = offsets.tolist()
    ref = torch.stack([dy[off[g]:off[g + 1]].t() @ x[off[g]:off[g + 1]] for g in range(3)])
    torch.testing.assert_close(got, ref, atol=1e-4, rtol=1e-4)


def test_bench_clock_sampler_queries_one_gpu_per_tick(monkeypatch):
    """The bench's clock sampler must touch ONE GPU per tick (round robin): NVML queries of all GPUs share a driver lock with
    CUDA-graph launches."""
    import importlib.util
    import os
    import sys
    import time
    import types
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    calls = []
    fake = types.SimpleNamespace(
        NVML_CLOCK_SM=0, nvmlInit=lambda: None, nvmlDeviceGetHandleByIndex=lambda i: i,
        nvmlDeviceGetClockInfo=lambda h, k: (calls.append(h), 1500 + 10 * h)[1], nvmlDeviceGetMaxClockInfo=lambda h, k: 1965,
        nvmlDeviceGetCurrentClocksEventReasons=lambda h: 0x4 if h == 2 else 0)
    monkeypatch.setitem(sys.modules, "pynvml", fake)
    monkeypatch.delenv("CUDA_VISIBLE_DEVICES", raising=False)
    s = bench.ClockSampler(8, period=0.05)
    s.start()
    time.sleep(0.33)
    out = s.stop()
    assert calls[:4] == [0, 1, 2, 3] and len(calls) == out["n_samples"] <= 8     # one GPU per tick, in order, first tick immediate
    assert out["gpus_sampled"] == len(set(calls)) and out["gpus"] == 8 and out["sm_max_mhz"] == 1965
    assert out["reasons"] == ["sw_power_cap"] and out["sm_mhz"] in [1500 + 10 * h for h in set(calls)]


def test_bench_clock_sampler_round_robin_with_a_fake_nvml(monkeypatch):
    """bench.py's sampler queries ONE GPU per tick (round robin) and folds clocks / throttle reasons into the JSON `clocks` entry."""
    import importlib.util
    import os
    import sys
    import time
    import types
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    calls = []
    fake = types.SimpleNamespace(
        NVML_CLOCK_SM=0, nvmlInit=lambda: None, nvmlDeviceGetHandleByIndex=lambda i: i,
        nvmlDeviceGetClockInfo=lambda h, k: calls.append(h) or 1900.0 - 10 * h, nvmlDeviceGetMaxClockInfo=lambda h, k: 1965.0,
        nvmlDeviceGetCurrentClocksEventReasons=lambda h: 0x4 if h == 2 else 0)
    monkeypatch.setitem(sys.modules, "pynvml", fake)
    monkeypatch.delenv("CUDA_VISIBLE_DEVICES", raising=False)
    s = bench.ClockSampler(4, period=0.02)
    s.start()
    deadline = time.time() + 10.0
    while len(calls) < 8 and time.time() < deadline:      # a loaded CI machine may starve the sampler thread for a while
        time.sleep(0.02)
    out = s.stop()
    assert calls[:8] == [0, 1, 2, 3, 0, 1, 2, 3]                      # one GPU per tick, in turn
    assert out["gpus"] == 4 and out["gpus_sampled"] == 4 and out["n_samples"] == len(calls) >= 8
    assert out["sm_max_mhz"] == 1965.0 and 1870.0 <= out["sm_mhz"] <= 1900.0 and out["reasons"] == ["sw_power_cap"]
    assert bench._NoSampler().stop() is None


def test_bench_driver_contract_without_a_gpu():
    """What the round driver relies on and what can be checked on CPU: `--impl reference` prints ONE JSON line with
    `"impl": "reference"` and exits 0 (from rank 0 only when launched through torchrun), the default flags are the documented ones,
    and the generation-layout choice is made by the allocation search's cost model for every node size the driver runs."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bench = os.path.join(root, "bench.py")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    p = subprocess.run([sys.executable, bench, "--impl", "reference", "--gpus", "8", "--steps", "2", "--warmup", "3"], capture_output=True, text=True,
                       timeout=300, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and ("unavailable" in out or "value" in out)
    p1 = subprocess.run([sys.executable, bench, "--impl", "reference", "--gpus", "8"], capture_output=True, text=True, timeout=300,
                        env=dict(env, RANK="3", LOCAL_RANK="3", WORLD_SIZE="8"))
    assert p1.returncode == 0 and not [l for l in p1.stdout.splitlines() if l.startswith("{")]
    h = subprocess.run([sys.executable, bench, "--help"], capture_output=True, text=True, timeout=300, env=env)
    assert h.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--runtime", "--allocation", "--optimizer"):
        assert flag in h.stdout
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_contract", bench)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for world in (1, 2, 4, 8):
        c = mod.choose_gen_tp(world, 128, 128, 512, 32)
        assert c["best"] in c["pred_s"] and all(world % tp == 0 for tp in c["pred_s"])


def test_graft_entry_points_exist_and_the_native_libraries_load_without_a_gpu():
    """`__graft_entry__.build()` is the driver's "does it build" check (nvcc cross-compiles sm_100a without a GPU) and `smoke()` its
    first GPU step: both must exist; the built libraries must load on a CPU-only machine and carry tcgen05 / TMA code."""
    import importlib.util
    import inspect
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("graft_entry_for_test", os.path.join(root, "__graft_entry__.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert callable(mod.build) and callable(mod.smoke)
    assert not inspect.signature(mod.build).parameters and not inspect.signature(mod.smoke).parameters
    from realhf_b200 import ops
    from realhf_b200.ops import build as B
    if not B.OPS_LIB.exists():
        pytest.skip("kernels not built in this checkout (run python -m realhf_b200.ops.build)")
    assert ops.lib() is not None and ops.host() is not None
    sass = B.sass_summary()
    assert sass["UTCHMMA"] > 0 and sass["UTMALDG"] > 0 and sass["LDTM"] > 0 and sass["HMMA"] == 0     # tcgen05 + TMA + TMEM loads, no legacy mma


def test_bench_json_line_carries_every_key_of_the_driver_contract():
    """Static check of the JSON line the SPMD arm prints (it needs a GPU to run): all keys the round driver reads are there."""
    import ast
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tree = ast.parse(open(os.path.join(root, "bench.py")).read())
    main = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "main")
    dicts = [n for n in ast.walk(main) if isinstance(n, ast.Dict)]
    keysets = [{k.value for k in d.keys if isinstance(k, ast.Constant)} for d in dicts]
    top = next(ks for ks in keysets if {"metric", "value", "unit"} <= ks)
    assert {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "clocks", "e2e", "gpu_launches", "impl"} <= top
    assert any({"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= ks for ks in keysets)             # e2e
    assert any({"model", "global_batch", "seq_len", "parallelism"} <= ks for ks in keysets)                        # config
    src = open(os.path.join(root, "bench.py")).read()
    assert "pin_memory()" in src and "non_blocking=True" in src          # inputs come from pinned host memory inside the timed region
    assert 'ap.add_argument("--warmup", type=int, default=3)' in src     # W >= 3 by default
