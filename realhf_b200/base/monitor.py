"""Timing / FLOP accounting utilities.

Parity: `realhf/base/monitor.py` — analytic LLaMA FLOP formulas (:277-351), CUDA time marks (`REAL_CUDA_TMARK`,
:354-445) and kernel-time categorisation of profiler traces (:449-514).  Time marks use CUDA events on the current
stream instead of the reference's `cuda.synchronize()` pairs, so marking does not serialise the device.
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


_KERNEL_CLASSES = {"collective": ("nccl", "allreduce", "all_gather", "reduce_scatter", "rb_ar_", "symm_"),
                   "p2p": ("sendrecv", "ncclDevKernel_SendRecv"), "memory": ("memcpy", "memset", "segcopy"),
                   "compute": ("gemm", "tcgen05", "flash", "attn", "rmsnorm", "adamw", "elementwise", "logprob")}


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
