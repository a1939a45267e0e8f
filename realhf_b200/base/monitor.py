"""Timing / FLOP accounting utilities.

Parity: `realhf/base/monitor.py` — host time marks + their summary / Gantt chart (:32-252), the NVML utilisation sampler
(:255-274), analytic LLaMA FLOP formulas (:277-351), CUDA time marks (`REAL_CUDA_TMARK`, :354-445), kernel-time
categorisation and the overlap-aware per-category statistics of profiler traces (:449-828).  CUDA time marks use events on
the current stream instead of the reference's `cuda.synchronize()` pairs, so marking does not serialise the device.
"""

import contextlib
import dataclasses
import enum
import os
import pickle
import time
from typing import Dict, List

import torch


def calculate_llama_train_flops(checkpoint_activations_factor: int, batch_size: int, seqlens: List[int], num_layers: int,
                                hidden_size: int, intermediate_size: int, vocab_size: int) -> float:
    return checkpoint_activations_factor * calculate_llama_forward_flops(batch_size, seqlens, num_layers, hidden_size,
                                                                         intermediate_size, vocab_size)


def calculate_llama_forward_flops(batch_size: int, seqlens: List[int], num_layers: int, hidden_size: int,
                                  intermediate_size: int, vocab_size: int) -> float:
    T = sum(seqlens)
    attn = sum(2 * 2 * s * s * hidden_size for s in seqlens) / 2  # causal
    per_layer = 2 * T * hidden_size * (4 * hidden_size + 3 * intermediate_size) + attn
    return num_layers * per_layer + 2 * T * hidden_size * vocab_size


def calculate_llama_gen_flops(batch_size: int, prompt_lens: List[int], gen_len: int, num_layers: int, hidden_size: int,
                              intermediate_size: int, vocab_size: int) -> float:
    f = calculate_llama_forward_flops(batch_size, prompt_lens, num_layers, hidden_size, intermediate_size, vocab_size)
    for i in range(gen_len):
        prefix = [p + i for p in prompt_lens]
        f += num_layers * (2 * batch_size * hidden_size * (4 * hidden_size + 3 * intermediate_size)
                           + sum(2 * 2 * p * hidden_size for p in prefix)) + 2 * batch_size * hidden_size * vocab_size
    return f


class CUDATimeMarkType(enum.Enum):
    forward = "forward"
    backward = "backward"
    optim_step = "optim_step"
    comm = "comm"
    misc = "misc"
    mem_layout = "memory_layout"


@dataclasses.dataclass
class TimeMarkEntry:
    name: str
    model_name: str
    type_: CUDATimeMarkType
    start: "torch.cuda.Event"
    end: "torch.cuda.Event"


TIME_MARK_DB: List[TimeMarkEntry] = []
_ENABLED = os.environ.get("REAL_CUDA_TMARK", "0") == "1"


@contextlib.contextmanager
def cuda_tmarked(name: str, type_: CUDATimeMarkType = CUDATimeMarkType.misc, model_name: str = ""):
    if not _ENABLED or not torch.cuda.is_available():
        yield
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    try:
        yield
    finally:
        e.record()
        TIME_MARK_DB.append(TimeMarkEntry(name, model_name, type_, s, e))


def cuda_tmark(name: str, type_: CUDATimeMarkType = CUDATimeMarkType.misc):
    def deco(fn):
        def wrapped(*a, **k):
            with cuda_tmarked(name, type_):
                return fn(*a, **k)
        return wrapped
    return deco


def dump_tmark_db(path: str):
    torch.cuda.synchronize()
    base = TIME_MARK_DB[0].start if TIME_MARK_DB else None
    rows = [dict(name=t.name, model=t.model_name, type=t.type_.value, start_ms=base.elapsed_time(t.start),
                 dur_ms=t.start.elapsed_time(t.end)) for t in TIME_MARK_DB]
    with open(path, "wb") as f:
        pickle.dump(rows, f)
    TIME_MARK_DB.clear()


_KERNEL_CLASSES = {"p2p": ("sendrecv", "ncclDevKernel_SendRecv"),   # before "collective": these names contain "nccl" too
                   "collective": ("nccl", "allreduce", "all_gather", "allgather", "reduce_scatter", "rb_ar_", "symm_", "nvls_",
                                  "ep_move_rows", "ep_plan", "barrier_kernel", "spin_wait"),
                   "memory": ("memcpy", "memset", "segcopy"),
                   "compute": ("gemm", "tcgen05", "flash", "attn", "rmsnorm", "layernorm", "adamw", "elementwise", "logprob", "wgrad",
                               "gated_act", "rope_kernel", "sample_kernel", "gae_", "quant_rows", "sumsq", "colsum", "reduce_slabs",
                               "cutlass", "cublas", "nvjet", "sm90_", "sm100_", "triton")}


def categorize_kernel(name: str) -> str:
    n = name.lower()
    for cls, keys in _KERNEL_CLASSES.items():
        if any(k.lower() in n for k in keys):
            return cls
    return "misc"


def summarize_chrome_trace(events: List[dict]) -> Dict[str, float]:
    """Total device time (us) per category from a torch.profiler chrome trace's kernel events."""
    out: Dict[str, float] = {}
    for e in events:
        if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset"):
            c = categorize_kernel(e.get("name", ""))
            out[c] = out.get(c, 0.0) + float(e.get("dur", 0.0))
    return out


# ------------------------------------------------------------------------------------------- host time marks
# Parity: monitor.py:32-252 (`time_mark`, `parse_time_mark_in_{line,file,dir}`, `summary_time_points`).  The reference writes
# marks as decorated debug-log lines and parses them back with string splits; here a mark is one JSON object per line (written
# through the `benchmark` logger, so it lands in the worker's log file like every other line), which survives identifiers
# containing `#` / `$` and carries arbitrary extra fields.
_HOST_MARKS = os.environ.get("REAL_TIME_MARK", "0") == "1"
_MARK_TAG = "TIMEMARK "


def enable_time_marks(flag: bool = True):
    global _HOST_MARKS
    _HOST_MARKS = flag


def time_mark(name: str, identifier: str, step: int = 0, t_ns: "int | None" = None, **extra):
    """Record `name` (e.g. "actor_train_start") for `identifier` (e.g. "model_worker/3") at `t_ns` (now by default)."""
    if not _HOST_MARKS:
        return None
    import json

    from realhf_b200.base import logging
    rec = dict(extra, name=name, id=str(identifier), t=int(time.time_ns() if t_ns is None else t_ns), step=int(step))
    logging.getLogger("benchmark").info(_MARK_TAG + json.dumps(rec, sort_keys=True))
    return rec


def parse_time_mark_in_line(line: str, name: "str | None" = None, step_range=None):
    """(identifier, t_ns, record) of a mark line, or None (not a mark / another name / outside `[lo, hi)` steps)."""
    import json
    i = line.find(_MARK_TAG)
    if i < 0:
        return None
    try:
        rec = json.loads(line[i + len(_MARK_TAG):].strip())
    except ValueError:
        return None
    if name is not None and rec.get("name") != name:
        return None
    if step_range is not None and not (step_range[0] <= rec.get("step", 0) < step_range[1]):
        return None
    return rec["id"], rec["t"], rec


def parse_time_marks(path: str, name: "str | None" = None, step_range=None) -> Dict[str, List[int]]:
    """{identifier: [t_ns, ...]} of the marks called `name` in a log file, or in every file of a directory."""
    files = [path] if os.path.isfile(path) else sorted(os.path.join(path, f) for f in os.listdir(path)
                                                       if os.path.isfile(os.path.join(path, f)))
    out: Dict[str, List[int]] = {}
    for fn in files:
        try:
            with open(fn, "r", errors="replace") as f:
                for line in f:
                    if _MARK_TAG not in line:
                        continue
                    r = parse_time_mark_in_line(line, name, step_range)
                    if r is not None:
                        out.setdefault(r[0], []).append(r[1])
        except OSError:
            continue
    for v in out.values():
        v.sort()
    return out


def summary_time_points(start_keys: List[str], end_keys: List[str], identifiers: List[str], path: str, step_range=None,
                        start_time: "int | None" = None, end_time: "int | None" = None, save_fig_path: "str | None" = None,
                        figsize=(12, 4)) -> Dict[str, dict]:
    """Pair the i-th `start_keys[k]` mark with the i-th `end_keys[k]` mark of every identifier and report, per identifier and
    key: n, sum / avg / min / max (ms) and the share of the covered wall-clock window; what is left is the `bubble`.  With
    `save_fig_path` also draws the Gantt chart (one row per identifier, one colour per key; needs matplotlib)."""
    assert len(start_keys) == len(end_keys)
    marks = {k: parse_time_marks(path, k, step_range) for k in set(start_keys) | set(end_keys)}
    spans: Dict[str, Dict[str, List[tuple]]] = {}
    lo = hi = None
    for ident in identifiers:
        spans[ident] = {}
        for sk, ek in zip(start_keys, end_keys):
            st, en = marks[sk].get(ident, []), marks[ek].get(ident, [])
            if len(st) != len(en):
                raise ValueError(f"{ident}: {len(st)} `{sk}` marks but {len(en)} `{ek}` marks")
            pairs = [(s, e) for s, e in zip(st, en) if (start_time is None or s > start_time) and (end_time is None or s < end_time)]
            spans[ident][sk] = pairs
            for s, e in pairs:
                lo = s if lo is None else min(lo, s)
                hi = e if hi is None else max(hi, e)
    window = max((hi - lo) if lo is not None else 0, 1)
    out: Dict[str, dict] = {}
    for ident in identifiers:
        rows, busy = {}, 0.0
        for sk in start_keys:
            d = [e - s for s, e in spans[ident][sk]]
            pct = 100.0 * sum(d) / window
            busy += pct
            rows[sk] = dict(n=len(d), sum_ms=sum(d) / 1e6, avg_ms=(sum(d) / len(d) / 1e6) if d else None,
                            min_ms=min(d) / 1e6 if d else None, max_ms=max(d) / 1e6 if d else None, percent=pct)
        out[ident] = dict(keys=rows, bubble_percent=100.0 - busy, window_ms=window / 1e6)
    if save_fig_path:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        colors = plt.rcParams["axes.prop_cycle"].by_key()["color"]
        fig, ax = plt.subplots(1, 1, figsize=figsize)
        for row, ident in enumerate(identifiers):
            for ki, sk in enumerate(start_keys):
                for j, (s, e) in enumerate(spans[ident][sk]):
                    ax.barh(y=row, width=(e - s) / 1e9, left=(s - lo) / 1e9, height=0.8, color=colors[ki % len(colors)],
                            label=sk if (row == 0 and j == 0) else None)
        ax.set_yticks(list(range(len(identifiers))))
        ax.set_yticklabels(identifiers)
        ax.set_xlabel("seconds")
        handles, labels = ax.get_legend_handles_labels()
        if handles:
            ax.legend(loc=(1.01, 0.0))
        fig.tight_layout()
        fig.savefig(save_fig_path)
        plt.close(fig)
    return out


# ------------------------------------------------------------------------------------------- GPU utilisation sampler
class GpuUtilizationMonitor:
    """Samples NVML utilisation / memory / SM clock / power of ONE device on a daemon thread (parity: `gpu_utilization_monitor`,
    monitor.py:255-274, which logs util + memory of GPU `worker_idx % 8` every `interval` seconds for `ttl` seconds).  Differences:
    the samples are kept (bounded ring) and summarised on demand so that they can ride on an MFC reply, and the default
    interval is seconds, not sub-second — NVML queries take a driver lock that CUDA-graph launches also need."""

    def __init__(self, device_index: int, interval: float = 2.0, ttl: "float | None" = None, max_samples: int = 4096, log: bool = False):
        import threading
        self.device_index, self.interval, self.ttl, self.log = device_index, interval, ttl, log
        self.samples: List[dict] = []
        self._max = max_samples
        self._stop = threading.Event()
        self._thread = None
        self.available = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.available = True
        except Exception:  # no driver / no NVML in this environment: the monitor stays a no-op
            self._nvml = self._handle = None

    def sample(self) -> "dict | None":
        if not self.available:
            return None
        n = self._nvml
        util = n.nvmlDeviceGetUtilizationRates(self._handle)
        mem = n.nvmlDeviceGetMemoryInfo(self._handle)
        rec = dict(t=time.time(), gpu=self.device_index, util=float(util.gpu), mem_used_mb=mem.used / 2 ** 20, mem_total_mb=mem.total / 2 ** 20)
        for key, fn in (("sm_mhz", lambda: n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM)),
                        ("power_w", lambda: n.nvmlDeviceGetPowerUsage(self._handle) / 1000.0)):
            try:
                rec[key] = float(fn())
            except Exception:
                pass
        return rec

    def _run(self):
        t0 = time.time()
        while not self._stop.is_set() and (self.ttl is None or time.time() - t0 < self.ttl):
            rec = self.sample()
            if rec is not None:
                self.samples.append(rec)
                if len(self.samples) > self._max:
                    del self.samples[: len(self.samples) - self._max]
                if self.log:
                    from realhf_b200.base import logging
                    logging.getLogger("benchmark").debug(
                        f"GPU {rec['gpu']}: compute utilization {rec['util']:.0f}%, memory {rec['mem_used_mb']:.0f} / "
                        f"{rec['mem_total_mb']:.0f} MB ({100 * rec['mem_used_mb'] / max(rec['mem_total_mb'], 1):.1f}%)")
            self._stop.wait(self.interval)

    def start(self):
        import threading
        if self.available and self._thread is None:
            self._thread = threading.Thread(target=self._run, name=f"gpu-util-{self.device_index}", daemon=True)
            self._thread.start()
        return self

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=self.interval + 1)
            self._thread = None

    def summary(self, since: "float | None" = None) -> Dict[str, float]:
        rows = [s for s in self.samples if since is None or s["t"] >= since]
        if not rows:
            return {}
        out = dict(n=len(rows), util_avg=sum(r["util"] for r in rows) / len(rows), util_max=max(r["util"] for r in rows),
                   mem_used_mb_max=max(r["mem_used_mb"] for r in rows))
        if all("sm_mhz" in r for r in rows):
            out["sm_mhz_avg"] = sum(r["sm_mhz"] for r in rows) / len(rows)
        if all("power_w" in r for r in rows):
            out["power_w_max"] = max(r["power_w"] for r in rows)
        return out


def gpu_utilization_monitor(worker_idx: int, interval: float, ttl: float, gpus_per_node: int = 8):
    """The reference's blocking form: log utilisation of GPU `worker_idx % gpus_per_node` every `interval` s for `ttl` s."""
    m = GpuUtilizationMonitor(worker_idx % gpus_per_node, interval=interval, ttl=ttl, log=True)
    if m.available:
        m._run()
    return m.summary()


# ------------------------------------------------------------------------------------------- kernel-time statistics
# Parity: monitor.py:449-828 (`CUDAKernelTimeCategory`, `CUDAKernelTimeStat`, `kernelStatFromEvents`, `kernelStatFromTrace`).
class CUDAKernelTimeCategory(enum.Enum):
    COMPUTE = "compute"
    P2P_COMM = "p2p_comm"
    COLL_COMM = "coll_comm"
    MEM = "memoryIO"
    IDLE = "idle"
    MISC = "misc"

    @classmethod
    def from_name(cls, name: str) -> "CUDAKernelTimeCategory":
        return _CATEGORY_OF[categorize_kernel(name)]


# what the device is "doing" when kernels of several categories overlap: the most useful one wins
_PRIORITY = [CUDAKernelTimeCategory.COMPUTE, CUDAKernelTimeCategory.COLL_COMM, CUDAKernelTimeCategory.P2P_COMM,
             CUDAKernelTimeCategory.MEM, CUDAKernelTimeCategory.MISC]
_CATEGORY_OF = {"compute": CUDAKernelTimeCategory.COMPUTE, "collective": CUDAKernelTimeCategory.COLL_COMM,
                "p2p": CUDAKernelTimeCategory.P2P_COMM, "memory": CUDAKernelTimeCategory.MEM, "misc": CUDAKernelTimeCategory.MISC}


class CUDAKernelTimeStat:
    """Device time (µs) per category, summed over `world_size` GPUs; `a + b` merges ranks, `/ n` and `gpu_average()` average."""

    def __init__(self, world_size: int = 1, **us):
        self.world_size = world_size
        for c in CUDAKernelTimeCategory:
            setattr(self, c.value, float(us.get(c.value, 0.0)))

    def as_dict(self) -> Dict[str, float]:
        return {c.value: getattr(self, c.value) for c in CUDAKernelTimeCategory}

    @property
    def total(self) -> float:
        return sum(self.as_dict().values())

    def percentage(self) -> Dict[str, float]:
        t = self.total or 1.0
        return {k: v / t for k, v in self.as_dict().items()}

    def __add__(self, other: "CUDAKernelTimeStat"):
        return CUDAKernelTimeStat(self.world_size + other.world_size, **{k: v + getattr(other, k) for k, v in self.as_dict().items()})

    def __truediv__(self, n: int):
        if self.world_size % n != 0:
            raise ValueError(f"cannot split the statistics of {self.world_size} GPUs into {n} parts")
        return CUDAKernelTimeStat(self.world_size // n, **{k: v / n for k, v in self.as_dict().items()})

    def gpu_average(self):
        return self / self.world_size

    def __repr__(self):
        from tabulate import tabulate
        pct = self.percentage()
        cats = list(CUDAKernelTimeCategory)
        rows = [["time (s)", f"{self.total / 1e6:.3f}"] + [f"{getattr(self, c.value) / 1e6:.3f}" for c in cats],
                ["share", "-"] + [f"{pct[c.value]:.1%}" for c in cats]]
        return f"kernel time over {self.world_size} GPU(s)\n" + tabulate(rows, headers=["", "total"] + [c.value for c in cats])


@dataclasses.dataclass
class KernelEventEntry:
    ts: float
    tid: int
    dur: float
    category: CUDAKernelTimeCategory


def kernel_stat_from_events(entries: List[KernelEventEntry], global_start_ts: float, global_end_ts: float) -> CUDAKernelTimeStat:
    """Sweep over the kernel intervals of ONE device: every instant of `[global_start_ts, global_end_ts]` is attributed to the
    highest-priority category with a kernel in flight (compute > collective > p2p > memory > misc) or to `idle` — so overlapped
    communication counts as compute time and the categories add up to the window, which is what makes per-rank numbers
    comparable (`global_*` are the first / last timestamps over ALL ranks: waiting for a slower rank shows up as idle)."""
    bounds = []
    for e in entries:
        s, t = max(e.ts, global_start_ts), min(e.ts + e.dur, global_end_ts)
        if t > s:
            bounds.append((s, 1, e.category))
            bounds.append((t, -1, e.category))
    bounds.sort(key=lambda b: (b[0], b[1]))
    active = {c: 0 for c in CUDAKernelTimeCategory}
    times = {c: 0.0 for c in CUDAKernelTimeCategory}
    cur = global_start_ts

    def account(upto):
        nonlocal cur
        if upto > cur:
            cat = next((c for c in _PRIORITY if active[c] > 0), CUDAKernelTimeCategory.IDLE)
            times[cat] += upto - cur
            cur = upto
    for ts, delta, cat in bounds:
        account(ts)
        active[cat] += delta
    account(global_end_ts)
    assert all(v == 0 for v in active.values()), active
    return CUDAKernelTimeStat(1, **{c.value: v for c, v in times.items()})


kernelStatFromEvents = kernel_stat_from_events   # the reference's spelling


def _match_send_recv(annotations: Dict[int, List[dict]]) -> Dict[int, List[float]]:
    """NCCL p2p kernels include the time spent waiting for the peer.  The profiler annotates them `nccl:send a->b` /
    `nccl:recv b<-a`; the k-th send a->b pairs with the k-th recv b<-a and both are charged the SHORTER of the two durations
    (the transfer itself).  Returns the corrected durations per rank in that rank's own time order.  (The reference resolves
    the pairing recursively over per-rank queues, monitor.py:746-796; pairing by (src, dst, k) is the same matching because
    NCCL orders the operations of one (src, dst) pair.)"""
    import re
    sends: Dict[tuple, List[tuple]] = {}
    recvs: Dict[tuple, List[tuple]] = {}
    order: Dict[int, List[tuple]] = {}
    for pid, evs in annotations.items():
        for i, ev in enumerate(sorted(evs, key=lambda e: e["ts"])):
            m = re.match(r"nccl:send (\d+)->(\d+)", ev["name"])
            if m:
                key, book = (int(m.group(1)), int(m.group(2))), sends
            else:
                m = re.match(r"nccl:recv (\d+)<-(\d+)", ev["name"])
                if not m:
                    continue
                key, book = (int(m.group(2)), int(m.group(1))), recvs
            book.setdefault(key, []).append((pid, i, float(ev["dur"])))
            order.setdefault(pid, []).append((book is sends, key, len(book[key]) - 1))
    out: Dict[int, List[float]] = {pid: [] for pid in annotations}
    for pid, ops in order.items():
        for is_send, key, k in ops:
            mine = (sends if is_send else recvs)[key][k][2]
            other = (recvs if is_send else sends).get(key, [])
            out[pid].append(min(mine, other[k][2]) if k < len(other) else mine)
    return out


def kernel_stat_from_trace(root_dir: str, mfc_name: str, per_rank: bool = False):
    """Kernel-time statistics of one MFC from the chrome traces `REAL_DUMP_TRACE=1` leaves under `<log>/trace/`
    (`<mfc>_r<rank>_c<call>.json`; the reference's `<mfc>_r<rank>.json` is read too).  Ranks are summed (use `.gpu_average()`);
    `per_rank=True` returns `{rank: stat}` instead."""
    import json
    import re
    traces: Dict[int, List[dict]] = {}
    for fn in sorted(os.listdir(root_dir)):
        m = re.fullmatch(re.escape(mfc_name) + r"_r(\d+)(?:_c\d+)?\.json", fn)
        if not m:
            continue
        with open(os.path.join(root_dir, fn)) as f:
            evs = json.load(f)
        evs = evs["traceEvents"] if isinstance(evs, dict) else evs
        keep = [e for e in evs if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset", "gpu_user_annotation") and "ts" in e]
        traces.setdefault(int(m.group(1)), []).extend(keep)
    if not traces:
        raise RuntimeError(f"no trace file of MFC `{mfc_name}` under {root_dir}")
    annotations = {pid: [e for e in evs if e["cat"] == "gpu_user_annotation" and e.get("name", "").startswith(("nccl:send", "nccl:recv"))]
                   for pid, evs in traces.items()}
    sr_time = _match_send_recv(annotations)
    kernels = {pid: sorted((e for e in evs if e["cat"] != "gpu_user_annotation" and float(e.get("dur", 0)) > 0), key=lambda e: e["ts"])
               for pid, evs in traces.items()}
    all_k = [e for evs in kernels.values() for e in evs]
    if not all_k:
        raise RuntimeError(f"the traces of `{mfc_name}` contain no device activity (CPU-only run?)")
    g0 = min(float(e["ts"]) for e in all_k)
    g1 = max(float(e["ts"]) + float(e["dur"]) for e in all_k)
    stats: Dict[int, CUDAKernelTimeStat] = {}
    for pid, evs in kernels.items():
        pending = list(sr_time.get(pid, []))
        entries = []
        for e in evs:
            cat = CUDAKernelTimeCategory.MEM if e["cat"] in ("gpu_memcpy", "gpu_memset") else CUDAKernelTimeCategory.from_name(e.get("name", ""))
            dur = float(e["dur"])
            if cat == CUDAKernelTimeCategory.P2P_COMM and pending:
                dur = min(dur, pending.pop(0))
            entries.append(KernelEventEntry(float(e["ts"]), int(e.get("tid", 0)), dur, cat))
        stats[pid] = kernel_stat_from_events(entries, g0, g1)
    if per_rank:
        return stats
    total = None
    for pid in sorted(stats):
        total = stats[pid] if total is None else total + stats[pid]
    return total


kernelStatFromTrace = kernel_stat_from_trace
