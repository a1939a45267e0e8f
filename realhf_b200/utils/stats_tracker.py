"""Process-wide side channel for values born deep inside a model (MoE aux / z losses, router statistics...).

Parity: `GLOBAL_STATS_TRACKER` in `realhf/base/constants.py:479-547`: `log(name, tensor, hook=...)` appends, interfaces
`pop(name)` at the end of a step; an optional reduce hook (e.g. a DP all-reduce) runs when the value is popped, so logging
never adds a collective inside the layer loop."""

from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch


class StatsTracker:
    def __init__(self):
        self._vals: Dict[str, List[torch.Tensor]] = {}
        self._hooks: Dict[str, Optional[Callable[[torch.Tensor], torch.Tensor]]] = {}

    def log(self, name: str, value: torch.Tensor, hook: Optional[Callable[[torch.Tensor], torch.Tensor]] = None):
        self._vals.setdefault(name, []).append(value.detach() if torch.is_tensor(value) else torch.as_tensor(value))
        if hook is not None:
            self._hooks[name] = hook

    def pop(self, name: str, reduce: str = "sum") -> Optional[torch.Tensor]:
        vals = self._vals.pop(name, None)
        if not vals:
            return None
        v = torch.stack([x.float() for x in vals])
        v = v.sum(0) if reduce == "sum" else v.mean(0)
        hook = self._hooks.pop(name, None)
        return hook(v) if hook is not None else v

    def pop_all(self, reduce: str = "sum") -> Dict[str, torch.Tensor]:
        return {k: self.pop(k, reduce) for k in list(self._vals)}

    def clear(self):
        self._vals.clear()
        self._hooks.clear()


GLOBAL_STATS_TRACKER = StatsTracker()
