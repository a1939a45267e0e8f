"""Name-keyed CUDA graph registry (parity: `realhf/impl/model/utils/cuda_graph.py`: capture_func :83-165, destroy).

`capture_func(name, fn, input_buffers, ...)` warms `fn` up once on a side stream, captures it into a `torch.cuda.CUDAGraph`
and returns `(graph, input_buffers, output_buffers)`; later calls with the same name return the cached entry unless
`force_recapture` is set.  The decode loop (`models/generation.py`) keeps its graph inside `DecodeState`; this registry is
for user code that wants the same capture-once / replay-many behaviour for arbitrary functions (custom interfaces)."""

from __future__ import annotations

from typing import Any, Callable, Dict, Optional, Tuple

import torch

_REGISTRY: Dict[str, Tuple["torch.cuda.CUDAGraph", Dict[str, Any], Any]] = {}


def capture_func(name: str, fn: Callable[..., Any], input_buffers: Dict[str, torch.Tensor], force_recapture: bool = False,
                 no_grad: bool = True, warmup: int = 1):
    """fn(**input_buffers) must only touch device memory (no host syncs).  Inputs are updated in place before `replay`."""
    if name in _REGISTRY and not force_recapture:
        return _REGISTRY[name]
    if name in _REGISTRY:
        destroy(name)
    dev = next(iter(input_buffers.values())).device
    assert dev.type == "cuda", "CUDA graphs need CUDA tensors"
    ctx = torch.no_grad() if no_grad else torch.enable_grad()
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with ctx, torch.cuda.stream(side):
        for _ in range(warmup):
            fn(**input_buffers)
    torch.cuda.current_stream(dev).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with ctx, torch.cuda.graph(graph):
        out = fn(**input_buffers)
    _REGISTRY[name] = (graph, input_buffers, out)
    return _REGISTRY[name]


def replay(name: str, **new_inputs: torch.Tensor):
    graph, bufs, out = _REGISTRY[name]
    for k, v in new_inputs.items():
        bufs[k].copy_(v)
    graph.replay()
    return out


def get(name: str) -> Optional[Tuple["torch.cuda.CUDAGraph", Dict[str, Any], Any]]:
    return _REGISTRY.get(name)


def destroy(name: str):
    entry = _REGISTRY.pop(name, None)
    if entry is not None:
        entry[0].reset()


def destroy_all():
    for k in list(_REGISTRY):
        destroy(k)
