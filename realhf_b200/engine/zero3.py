"""ZeRO-3 with per-layer parameter gather / release.

Parity: DeepSpeed ZeRO stage 3 as configured by the reference (`backend/deepspeed.py:276-322`): parameters, gradients and
optimizer state are all sharded over the data-parallel group; a layer's parameters exist in full only while that layer runs.

Design here (one flat buffer per model makes this compact):
  * every transformer block is a *bucket*: its slice of the flat layout is cut into dp equal slices (64-element aligned, the
    last ones possibly shorter) and rank r keeps slice r of EVERY block in one shard tensor (on the GPU, or in pinned host
    memory with `offload_param`); the optimizer state has the same shard shape;
  * forward: block i is all-gathered into one of two scratch buffers (the next block's gather is issued on a side stream
    while block i computes), its parameter views are re-pointed into the scratch, the block runs WITHOUT keeping
    activations, the views are dropped;
  * backward: the block is gathered again, recomputed under autograd, back-propagated; its parameter gradients land in a
    gradient scratch and are reduce-scattered into this rank's gradient shard at once; scratch memory is reused by the next
    block.  Peak memory = shards + 2 parameter scratches + 1 gradient scratch + one block's activations, independent of depth;
  * the embedding and the output head stay resident for the duration of an engine call (the head is used by the loss outside
    the layer loop) and are sharded between calls like the blocks;
  * `materialize()` / `release()` of the optimizer still build / drop the whole flat buffer for the paths that need all
    parameters at once (generation, checkpoint save, parameter reallocation out of a ZeRO-3 layout).
"""

from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from realhf_b200.ops import functional as OF


def _all_gather_into(dst: torch.Tensor, src: torch.Tensor, group, dp: int):
    if dst.is_cuda:
        dist.all_gather_into_tensor(dst, src, group=group)
    else:
        parts = [torch.empty_like(src) for _ in range(dp)]
        dist.all_gather(parts, src.contiguous(), group=group)
        dst.copy_(torch.cat(parts))


def _reduce_scatter_avg(dst: torch.Tensor, src: torch.Tensor, group, dp: int, rank: int):
    if src.is_cuda:
        dist.reduce_scatter_tensor(dst, src, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(src, group=group)
        per = src.numel() // dp
        dst.copy_(src[rank * per:(rank + 1) * per] / dp)


class Zero3Layers:
    def __init__(self, optim):
        self.optim = optim
        self.model = m = optim.model
        self.ctx = ctx = optim.ctx
        self.dp, self.rank, self.group = ctx.dp_size, ctx.dp_rank, ctx.dp_group
        dev = m.device
        cfg = m.config
        # layer -> [lo, hi) of the flat layout, names of its parameters
        spans: Dict[int, List[int]] = {}
        self.names: Dict[int, List[str]] = {}
        for name, slot in m.slots.items():
            li = int(name.split(".", 1)[0])
            end = slot.offset + (slot.numel + 63) // 64 * 64
            s = spans.setdefault(li, [slot.offset, end])
            s[0], s[1] = min(s[0], slot.offset), max(s[1], end)
            self.names.setdefault(li, []).append(name)
        self.layers = sorted(spans)
        self.block_layers = [li for li in self.layers if 1 <= li <= cfg.n_layers]
        self.resident_layers = [li for li in self.layers if li not in self.block_layers]
        self.span = {li: (spans[li][0], spans[li][1]) for li in self.layers}
        self.per = {li: ((spans[li][1] - spans[li][0] + self.dp - 1) // self.dp + 63) // 64 * 64 for li in self.layers}
        self.off: Dict[int, int] = {}
        acc = 0
        for li in self.layers:
            self.off[li] = acc
            acc += self.per[li]
        self.shard_n = acc
        pin = optim.cfg.offload_param and dev.type == "cuda"
        self.pshard = torch.zeros(acc, dtype=m.dtype, device="cpu" if pin else dev, pin_memory=pin)
        self.gshard = torch.zeros(acc, dtype=optim.grad_dtype, device=dev)
        flat = m.flat_param.data
        for li in self.layers:
            a, b = self._my_range(li)
            if b > a:
                self.pshard[self.off[li]: self.off[li] + (b - a)].copy_(flat[a:b])
        max_pad = max(self.per[li] * self.dp for li in self.block_layers) if self.block_layers else 64
        self.scratch_p = [torch.empty(max_pad, dtype=m.dtype, device=dev) for _ in range(2)]
        self.scratch_g = torch.empty(max_pad, dtype=optim.grad_dtype, device=dev)
        self.res_p: Dict[int, torch.Tensor] = {}
        self.res_g: Dict[int, torch.Tensor] = {}
        self._slot_of: Dict[int, int] = {}      # block currently held by each scratch buffer
        self._gather_ev: Dict[int, object] = {}
        self._side = torch.cuda.Stream(dev) if dev.type == "cuda" else None
        self.streaming = False                   # True while the model runs with per-layer gathers
        self.n_gathers = 0
        if optim.grad_dtype != m.dtype:  # fp32 gradient scratch: autograd's bf16 gradients are folded into `main_grad`
            for prm in m.p.values():
                if not getattr(prm, "_main_grad_hooked", False):
                    def _fold(param):
                        if param.grad is None or getattr(param, "main_grad", None) is None:
                            return
                        param.main_grad.add_(param.grad.view_as(param.main_grad))
                        param.grad = None
                    prm.register_post_accumulate_grad_hook(_fold)
                    prm._main_grad_hooked = True
        m.detach_params()
        m._zero3 = self

    # ---------------------------------------------------------------- layout helpers
    def _my_range(self, li: int) -> Tuple[int, int]:
        lo, hi = self.span[li]
        a = lo + self.rank * self.per[li]
        return a, max(a, min(hi, a + self.per[li]))

    def _my_pshard(self, li: int) -> torch.Tensor:
        return self.pshard[self.off[li]: self.off[li] + self.per[li]]

    def _point(self, li: int, buf: Optional[torch.Tensor], grads: Optional[torch.Tensor] = None):
        """Re-point the parameter views of layer li at `buf` (laid out like the layer's slice of the flat buffer)."""
        m = self.model
        lo, _ = self.span[li]
        for name in self.names[li]:
            slot = m.slots[name]
            p = m.p[name]
            if buf is None:
                p.data = torch.empty(0, dtype=m.dtype, device=m.device)
                p.grad = None
                if hasattr(p, "main_grad"):
                    p.main_grad = None
                continue
            p.data = buf[slot.offset - lo: slot.offset - lo + slot.numel].view(slot.shape)
            if grads is not None:
                g = grads[slot.offset - lo: slot.offset - lo + slot.numel].view(slot.shape)
                if grads.dtype == p.dtype:
                    p.grad = g
                    p._grad_in_flat_buffer = True
                else:
                    p.grad = None
                    p.main_grad = g

    # ---------------------------------------------------------------- gather / release
    def _gather_into(self, li: int, buf: torch.Tensor):
        """buf[: per * dp] <- all ranks' slices of layer li (this rank's slice comes from its shard, possibly in host memory)."""
        per, n = self.per[li], self.per[li] * self.dp
        mine = buf[self.rank * per:(self.rank + 1) * per]
        mine.copy_(self._my_pshard(li), non_blocking=True)
        if self.dp > 1:
            if buf.is_cuda:
                dist.all_gather_into_tensor(buf[:n], mine, group=self.group)  # in place: my slice already sits at its position
            else:
                parts = [torch.empty(per, dtype=buf.dtype) for _ in range(self.dp)]
                dist.all_gather(parts, mine.clone(), group=self.group)
                buf[:n].copy_(torch.cat(parts))
        self.n_gathers += 1

    def prefetch(self, li: Optional[int]):
        """Start gathering block li into the scratch buffer that is not in use (side stream on CUDA)."""
        if li is None or li in self._slot_of.values():
            return
        busy = set(self._slot_of)  # slots currently pointing at a live block
        slot = next(s for s in (0, 1) if s not in busy) if len(busy) < 2 else None
        if slot is None:
            return
        if self._side is not None:
            self._side.wait_stream(torch.cuda.current_stream(self.model.device))
            with torch.cuda.stream(self._side):
                self._gather_into(li, self.scratch_p[slot])
                ev = torch.cuda.Event()
                ev.record(self._side)
            self._gather_ev[li] = ev
        else:
            self._gather_into(li, self.scratch_p[slot])
        self._slot_of[slot] = li

    def acquire(self, li: int, with_grads: bool = False):
        """Make block li's parameters usable (gathering them now unless a prefetch already did)."""
        slot = next((s for s, l in self._slot_of.items() if l == li), None)
        if slot is None:
            self.prefetch(li)
            slot = next(s for s, l in self._slot_of.items() if l == li)
        ev = self._gather_ev.pop(li, None)
        if ev is not None:
            torch.cuda.current_stream(self.model.device).wait_event(ev)
        grads = None
        if with_grads:
            grads = self.scratch_g[: self.per[li] * self.dp]
            grads.zero_()
        self._point(li, self.scratch_p[slot], grads)
        return slot

    def release(self, li: int):
        slot = next((s for s, l in self._slot_of.items() if l == li), None)
        if slot is not None:
            del self._slot_of[slot]
        self._point(li, None)

    def reduce_grads(self, li: int):
        """Gradient scratch of block li -> this rank's gradient shard (accumulated over micro-batches)."""
        n = self.per[li] * self.dp
        dst = self.gshard[self.off[li]: self.off[li] + self.per[li]]
        if self.dp > 1:
            tmp = torch.empty(self.per[li], dtype=self.gshard.dtype, device=self.gshard.device)
            _reduce_scatter_avg(tmp, self.scratch_g[:n], self.group, self.dp, self.rank)
            dst.add_(tmp)
        else:
            dst.add_(self.scratch_g[:n])

    # ---------------------------------------------------------------- engine-call scope (resident layers)
    def begin_call(self, train: bool):
        """Drop the full flat buffer (if any), gather the resident layers (embedding / head) for the duration of the call."""
        m = self.model
        self.streaming = True
        m._zero3 = self
        for li in self.resident_layers:
            n = self.per[li] * self.dp
            buf = torch.empty(n, dtype=m.dtype, device=m.device)
            self._gather_into(li, buf)
            self.res_p[li] = buf
            g = None
            if train:
                g = self.res_g.get(li)
                if g is None:
                    g = self.res_g[li] = torch.zeros(n, dtype=self.gshard.dtype, device=m.device)
            self._point(li, buf, g)

    def end_call(self):
        for li in self.resident_layers:
            self._point(li, None)
        self.res_p.clear()
        for li in list(self._slot_of.values()):
            self.release(li)
        self.streaming = False

    def zero_grad(self):
        self.gshard.zero_()
        for g in self.res_g.values():
            g.zero_()

    def finish_grads(self):
        """Resident layers' gradients -> gradient shard (once per step, after the last micro-batch)."""
        for li in self.resident_layers:
            g = self.res_g.get(li)
            if g is None:
                continue
            dst = self.gshard[self.off[li]: self.off[li] + self.per[li]]
            if self.dp > 1:
                tmp = torch.empty(self.per[li], dtype=g.dtype, device=g.device)
                _reduce_scatter_avg(tmp, g, self.group, self.dp, self.rank)
                dst.add_(tmp)
            else:
                dst.add_(g)

    # ---------------------------------------------------------------- running a block
    def run_block(self, li: int, fn: Callable, x: torch.Tensor, d: Optional[torch.Tensor]):
        nxt = self._neighbour(li, +1)
        if not torch.is_grad_enabled():
            self.acquire(li)
            self.prefetch(nxt)
            out = fn(x, d)
            self.release(li)
            return out
        dd = d if d is not None else x.new_empty(0)
        x_out, d_out = _Zero3Block.apply(self, li, fn, d is not None, x, dd)
        return x_out, (d_out if d_out.numel() else None)

    def _neighbour(self, li: int, step: int) -> Optional[int]:
        j = self.block_layers.index(li) + step
        return self.block_layers[j] if 0 <= j < len(self.block_layers) else None

    # ---------------------------------------------------------------- whole-model materialisation (generation, save, realloc)
    def materialize_full(self) -> torch.Tensor:
        m = self.model
        flat = torch.empty(m.flat_numel, dtype=m.dtype, device=m.device)
        tmp = torch.empty(max(self.per[li] * self.dp for li in self.layers), dtype=m.dtype, device=m.device)
        for li in self.layers:
            lo, hi = self.span[li]
            self._gather_into(li, tmp)
            flat[lo:hi].copy_(tmp[: hi - lo])
        return flat

    def absorb_full(self, flat: torch.Tensor):
        """Refresh the parameter shard from a full flat buffer (after it was written by someone else, e.g. a checkpoint load)."""
        for li in self.layers:
            a, b = self._my_range(li)
            if b > a:
                self.pshard[self.off[li]: self.off[li] + (b - a)].copy_(flat[a:b])

    # ---------------------------------------------------------------- optimizer step over the shard space
    def step(self, lr: float) -> None:
        """AdamW on this rank's shard of every layer.  Parameter shard and optimizer state may each live on the GPU or in pinned
        host memory (`offload_param` / optimizer `offload`): host-resident pieces are streamed through the GPU, the update
        itself always runs in the fused kernel."""
        o, cfg = self.optim, self.optim.cfg
        dev = self.model.device
        p_on_host = dev.type == "cuda" and not self.pshard.is_cuda
        for li in self.layers:
            a, b = self._my_range(li)
            n_my = b - a
            if n_my <= 0:
                continue
            so = self.off[li]
            g = self.gshard[so: so + n_my]
            p = self.pshard[so: so + n_my]
            if p_on_host:
                p = p.to(dev, non_blocking=True)
            if o.offload:   # m / v / master in pinned host memory: chunked, double-buffered (FlatAdamW._offloaded_update)
                o._offloaded_update(p, g, lr, so)
            else:
                mst = o.master[so: so + n_my] if o.master is not None else None
                OF.adamw_step(p, g, o.m[so: so + n_my], o.v[so: so + n_my], mst, lr, cfg.beta1, cfg.beta2, cfg.eps, cfg.weight_decay,
                              o.step_count, o._scale, o._skip, stochastic=(o.state_dtype == torch.bfloat16),
                              seed=(o.step_count * 2654435761 + 7919 * li) % (2 ** 31))
            if p_on_host:
                self.pshard[so: so + n_my].copy_(p, non_blocking=True)
        if p_on_host:
            torch.cuda.current_stream(dev).synchronize()

    def grad_sumsq(self, stats: torch.Tensor):
        for li in self.layers:
            a, b = self._my_range(li)
            if b > a:
                OF.sumsq_accum(self.gshard[self.off[li]: self.off[li] + (b - a)], stats)


class _Zero3Block(torch.autograd.Function):
    """One transformer block under per-layer ZeRO-3: forward keeps only the block's inputs; backward re-gathers the parameters,
    recomputes the block, back-propagates and reduce-scatters the parameter gradients immediately."""

    @staticmethod
    def forward(ctx, z: Zero3Layers, li: int, fn, has_d: bool, x, d):
        z.acquire(li)
        z.prefetch(z._neighbour(li, +1))
        with torch.no_grad():
            x_out, d_out = fn(x, d if has_d else None)
        z.release(li)
        ctx.z, ctx.li, ctx.fn, ctx.has_d = z, li, fn, has_d
        ctx.save_for_backward(x, d)
        if d_out is None:
            d_out = x_out.new_empty(0)
        return x_out, d_out

    @staticmethod
    def backward(ctx, gx, gd):
        z, li, fn = ctx.z, ctx.li, ctx.fn
        x, d = ctx.saved_tensors
        z.acquire(li, with_grads=True)
        z.prefetch(z._neighbour(li, -1))
        with torch.enable_grad():
            xi = x.detach().requires_grad_(True)
            di = d.detach().requires_grad_(True) if ctx.has_d else None
            x_out, d_out = fn(xi, di)
            outs, grads = [x_out], [gx]
            if d_out is not None and gd is not None and gd.numel():
                outs.append(d_out)
                grads.append(gd)
        torch.autograd.backward(outs, grads)
        # parameters whose gradient came through autograd's .grad / main_grad folding already sit in the gradient scratch
        m = z.model
        for name in z.names[li]:
            p = m.p[name]
            mg = getattr(p, "main_grad", None)
            if mg is not None and p.grad is not None:  # fp32 scratch, bf16 autograd gradient: fold it
                mg.add_(p.grad.view_as(mg))
                p.grad = None
        z.reduce_grads(li)
        z.release(li)
        return None, None, None, None, xi.grad, (di.grad if di is not None else None)
