"""base/monitor.py: host time marks and their summary, overlap-aware kernel-time statistics from profiler traces, the NVML
sampler's no-driver behaviour.  Reference: realhf/base/monitor.py:32-274, :449-828."""
import json
import logging as pylogging
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from realhf_b200.base import monitor as M  # noqa: E402
from realhf_b200.base.monitor import CUDAKernelTimeCategory as C  # noqa: E402


def test_time_marks_round_trip_and_summary(tmp_path):
    M.enable_time_marks(True)
    try:
        lines = []
        h = pylogging.Handler()
        h.emit = lambda rec: lines.append(rec.getMessage())
        lg = pylogging.getLogger("benchmark")
        lg.addHandler(h)
        lg.setLevel(pylogging.DEBUG)
        ms = 1_000_000
        for w, off in (("model_worker/0", 0), ("model_worker/1", 5 * ms)):
            for step in range(3):
                base = step * 100 * ms + off
                M.time_mark("gen_start", w, step, t_ns=base)
                M.time_mark("gen_end", w, step, t_ns=base + 40 * ms)
                M.time_mark("train_start", w, step, t_ns=base + 50 * ms, note="#$&")
                M.time_mark("train_end", w, step, t_ns=base + 80 * ms)
        lg.removeHandler(h)
        assert len(lines) == 24
        log = tmp_path / "logs"
        log.mkdir()
        (log / "a.log").write_text("noise\n" + "\n".join("2026 INFO " + l for l in lines[:12]) + "\n")
        (log / "b.log").write_text("\n".join(lines[12:]) + "\nTIMEMARK {broken json\n")
        ident, t, rec = M.parse_time_mark_in_line("x " + lines[2], "train_start")
        assert (ident, t, rec["note"]) == ("model_worker/0", 50 * ms, "#$&")
        assert M.parse_time_mark_in_line(lines[2], "gen_start") is None
        assert M.parse_time_mark_in_line(lines[2], "train_start", step_range=(1, 3)) is None
        marks = M.parse_time_marks(str(log), "gen_start")
        assert marks == {"model_worker/0": [0, 100 * ms, 200 * ms], "model_worker/1": [5 * ms, 105 * ms, 205 * ms]}
        assert M.parse_time_marks(str(log / "a.log"), "gen_end", step_range=(0, 1)) == {"model_worker/0": [40 * ms]}
        fig = str(tmp_path / "gantt.png")
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            fig = None
        s = M.summary_time_points(["gen_start", "train_start"], ["gen_end", "train_end"], ["model_worker/0", "model_worker/1"], str(log),
                                  save_fig_path=fig)
        w0 = s["model_worker/0"]
        assert w0["window_ms"] == pytest.approx(285.0)
        assert w0["keys"]["gen_start"] == dict(n=3, sum_ms=120.0, avg_ms=40.0, min_ms=40.0, max_ms=40.0, percent=pytest.approx(100 * 120 / 285))
        assert w0["keys"]["train_start"]["sum_ms"] == 90.0
        assert w0["bubble_percent"] == pytest.approx(100 - 100 * 210 / 285)
        if fig:
            assert os.path.getsize(fig) > 0
        only1 = M.summary_time_points(["gen_start"], ["gen_end"], ["model_worker/1"], str(log), step_range=(1, 2))
        assert only1["model_worker/1"]["keys"]["gen_start"]["n"] == 1 and only1["model_worker/1"]["bubble_percent"] == pytest.approx(0.0)
        (log / "c.log").write_text(lines[0] + "\n")   # a start mark without its end
        with pytest.raises(ValueError, match="marks but"):
            M.summary_time_points(["gen_start"], ["gen_end"], ["model_worker/0"], str(log))
    finally:
        M.enable_time_marks(False)
    assert M.time_mark("x", "y") is None


def test_kernel_categories_cover_this_repos_kernels():
    expect = {"gemm_2cta_kernel<1,2>": C.COMPUTE, "attn_fwd_kernel": C.COMPUTE, "decode_attn_reduce_kernel": C.COMPUTE,
              "rmsnorm_bwd_kernel": C.COMPUTE, "adamw_kernel": C.COMPUTE, "sample_kernel": C.COMPUTE, "gae_1d_kernel": C.COMPUTE,
              "nvls_rs_sumsq_kernel": C.COLL_COMM, "nvls_adam_ag_kernel": C.COLL_COMM, "allreduce_1shot_kernel": C.COLL_COMM,
              "ncclDevKernel_AllGather_RING_LL": C.COLL_COMM, "ncclDevKernel_SendRecv": C.P2P_COMM, "ep_move_rows_kernel": C.COLL_COMM,
              "segcopy_kernel": C.MEM, "Memcpy DtoD (Device -> Device)": C.MEM, "at::native::vectorized_elementwise_kernel": C.COMPUTE,
              "something_else": C.MISC}
    for name, cat in expect.items():
        assert C.from_name(name) == cat, name


def test_kernel_stat_sweep_gives_overlap_to_the_higher_priority_category():
    E = M.KernelEventEntry
    ev = [E(10, 0, 30, C.COMPUTE),        # 10..40
          E(30, 1, 30, C.COLL_COMM),      # 30..60: 30..40 hidden under compute -> 20 us of exposed collective
          E(55, 2, 10, C.MEM),            # 55..65: 55..60 hidden under the collective -> 5 us
          E(70, 0, 5, C.MISC),
          E(72, 1, 10, C.P2P_COMM),       # 72..82, wins over misc
          E(95, 0, 50, C.COMPUTE)]        # clipped at the window end (100)
    st = M.kernel_stat_from_events(ev, 0, 100)
    assert st.as_dict() == {"compute": 35.0, "coll_comm": 20.0, "p2p_comm": 10.0, "memoryIO": 5.0, "misc": 2.0, "idle": 28.0}
    assert st.total == 100.0 and st.world_size == 1
    two = st + M.CUDAKernelTimeStat(1, compute=65.0, idle=35.0)
    assert two.world_size == 2 and two.compute == 100.0 and two.gpu_average().compute == 50.0 and two.gpu_average().world_size == 1
    assert two.percentage()["compute"] == pytest.approx(0.5)
    with pytest.raises(ValueError):
        (two + st) / 2
    text = repr(two)
    assert "2 GPU" in text and "coll_comm" in text
    assert M.kernelStatFromEvents is M.kernel_stat_from_events


def _trace(path, events):
    with open(path, "w") as f:
        json.dump({"traceEvents": events}, f)


def test_kernel_stat_from_trace_pairs_send_recv_and_sums_ranks(tmp_path):
    k = lambda name, ts, dur, cat="kernel": dict(name=name, ts=ts, dur=dur, cat=cat, tid=7, ph="X")  # noqa: E731
    # rank 0 computes 0..100 then sends for 10; rank 1 posted its recv at 20 and waited: its 90 us recv kernel is charged 10 us
    _trace(tmp_path / "actor_train_r0_c1.json", [
        k("gemm_2cta_kernel", 0, 100), k("ncclDevKernel_SendRecv(ncclDevKernelArgsStorage)", 100, 10),
        k("nccl:send 0->1", 100, 10, "gpu_user_annotation"), k("cpu_op", 0, 500, "cpu_op"),
        k("Memcpy DtoD", 112, 8, "gpu_memcpy")])
    _trace(tmp_path / "actor_train_r1_c1.json", [
        k("ncclDevKernel_SendRecv(ncclDevKernelArgsStorage)", 20, 90), k("nccl:recv 1<-0", 20, 90, "gpu_user_annotation"),
        k("attn_fwd_kernel", 110, 10)])
    _trace(tmp_path / "other_mfc_r0_c1.json", [k("gemm", 0, 1000)])
    per = M.kernel_stat_from_trace(str(tmp_path), "actor_train", per_rank=True)
    assert sorted(per) == [0, 1]
    assert per[0].as_dict() == {"compute": 100.0, "p2p_comm": 10.0, "coll_comm": 0.0, "memoryIO": 8.0, "idle": 2.0, "misc": 0.0}
    assert per[1].as_dict() == {"compute": 10.0, "p2p_comm": 10.0, "coll_comm": 0.0, "memoryIO": 0.0, "idle": 100.0, "misc": 0.0}
    both = M.kernel_stat_from_trace(str(tmp_path), "actor_train")
    assert both.world_size == 2 and both.total == 240.0 and both.gpu_average().idle == 51.0
    with pytest.raises(RuntimeError, match="no trace file"):
        M.kernel_stat_from_trace(str(tmp_path), "critic_train")


def test_gpu_utilization_monitor_without_a_driver_is_a_noop():
    m = M.GpuUtilizationMonitor(0, interval=0.01).start()
    if m.available:   # a machine with NVML: the sampler must produce sane numbers instead
        import time
        time.sleep(0.1)
        m.stop()
        s = m.summary()
        assert s["n"] >= 1 and 0 <= s["util_avg"] <= 100
    else:
        assert m.sample() is None and m.summary() == {}
        m.stop()
        assert M.gpu_utilization_monitor(3, 0.01, 0.05) == {}
