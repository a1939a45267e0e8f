"""Sharded flat AdamW (ZeRO-1) with device-side clipping / overflow skip, LR schedules and host offload.

Replaces what the reference borrows from Megatron-core (`DistributedOptimizer`, `clip_grad_norm_fp32`,
`OptimizerParamScheduler`; backend/megatron.py:158-520) and DeepSpeed (`FusedAdam`, `DeepSpeedCPUAdam`,
ZeRO offload; backend/deepspeed.py:276-475).  One step is:

    reduce-scatter grads over DP -> (SP: all-reduce norm grads over TP) -> sum-of-squares kernel ->
    ONE packed all-reduce of [sumsq, nonfinite] -> clip coefficient and skip flag computed on the device ->
    fused AdamW kernel on this rank's shard of the flat buffer -> all-gather updated params over DP.

No host sync anywhere on the path (the reference syncs for found_inf and grad norm).  With
`offload=True` the fp32 master weights and Adam moments live in pinned host memory and are streamed
through the GPU in double-buffered chunks (B200's answer to DeepSpeedCPUAdam: the update still runs in
the fused kernel at HBM speed, the host link only carries the state).
"""

from __future__ import annotations

import dataclasses
import math
import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from realhf_b200.base.topology import ParallelContext
from realhf_b200.ops import functional as OF
from realhf_b200.ops import lib


@dataclasses.dataclass
class OptimizerConfig:
    """Same knobs as the reference `api/quickstart/model.py:62-110` (defaults included)."""

    type: str = "adam"
    lr: float = 1e-5
    weight_decay: float = 0.05
    beta1: float = 0.9
    beta2: float = 0.95
    eps: float = 1e-5
    min_lr_ratio: float = 0.0
    lr_scheduler_type: str = "cosine"  # linear | cosine | constant
    warmup_steps_proportion: float = 0.02
    offload: bool = False
    initial_loss_scale: float = 2 ** 32
    min_loss_scale: float = 1.0
    loss_scale_window: float = 5
    hysteresis: int = 2
    gradient_clipping: float = 1.0
    # B200-native additions
    state_dtype: str = "fp32"      # fp32 | bf16 (bf16 moments, stochastic rounding, no master copy)
    use_master_weights: bool = True
    grad_dtype: str = "bf16"       # dtype of the flat gradient buffer: bf16 | fp32
    share_grad_buffer: bool = False  # trainable models on one GPU that never train concurrently share one grad buffer
    zero_stage: int = 1              # 1/2: optimizer state (+ reduced grads) sharded over DP; 3: parameters too, between calls
    offload_param: bool = False      # ZeRO-3 only: park the parameter shard in pinned host memory between calls
    bucket_numel: int = 1 << 28      # gradient bucket size (elements) of the overlapped reduce-scatter; 0: one bucket
    comm: str = "auto"               # auto | nvls | nccl: transport of the gradient reduce-scatter / parameter all-gather


class LRScheduler:
    """Warmup + {linear, cosine, constant} decay to `min_lr_ratio * lr`, indexed by absolute step."""

    def __init__(self, cfg: OptimizerConfig, total_steps: int):
        self.cfg = cfg
        self.total = max(1, int(total_steps))
        self.warmup = int(cfg.warmup_steps_proportion * self.total)
        self.step = 0

    def lr_at(self, step: int) -> float:
        c = self.cfg
        if self.warmup > 0 and step < self.warmup:
            return c.lr * (step + 1) / self.warmup
        if c.lr_scheduler_type == "constant":
            return c.lr
        prog = min(1.0, (step - self.warmup) / max(1, self.total - self.warmup))
        lo = c.lr * c.min_lr_ratio
        if c.lr_scheduler_type == "linear":
            return lo + (c.lr - lo) * (1.0 - prog)
        if c.lr_scheduler_type == "cosine":
            return lo + (c.lr - lo) * 0.5 * (1.0 + math.cos(math.pi * prog))
        raise ValueError(c.lr_scheduler_type)

    def step_absolute(self, step: int):
        self.step = int(step)

    def get_lr(self) -> float:
        return self.lr_at(self.step)

    def state_dict(self):
        return {"step": self.step}

    def load_state_dict(self, sd):
        self.step = sd["step"]


_DT = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}

_GRAD_POOL: Dict[Tuple, Tuple[torch.Tensor, object]] = {}


def _nvls_usable(ctx: ParallelContext, dev: torch.device, cfg: "OptimizerConfig") -> bool:
    """The in-switch (NVLS) ZeRO path needs: CUDA, a data-parallel group of 2..8 ranks on one NVSwitch domain, multicast
    support in the driver, and optimizer state resident on the GPU (`REAL_ZERO_COMM=nccl|nvls` overrides the choice)."""
    want = os.environ.get("REAL_ZERO_COMM", cfg.comm)
    if want == "nccl" or dev.type != "cuda" or ctx.dp_size < 2 or ctx.dp_size > 8 or cfg.zero_stage >= 3 or cfg.offload:
        return False
    if dist.get_backend(ctx.dp_group) != "nccl":
        return False
    from realhf_b200.parallel import symm_mem
    ok = symm_mem.multicast_supported(dev)
    flags: List = [None] * ctx.dp_size
    dist.all_gather_object(flags, (bool(ok), os.uname().nodename), group=ctx.dp_group)
    ok = all(f[0] for f in flags) and len({f[1] for f in flags}) == 1
    if want == "nvls" and not ok:
        raise RuntimeError("REAL_ZERO_COMM=nvls but the data-parallel group has no NVSwitch multicast support")
    return ok


def _grad_pool_get(numel: int, dtype, device, symm_group=None):
    """One gradient buffer per (dtype, device, group), grown to the largest request; every train_batch zeroes it first and
    consumes it in its own optimizer step, so sequentially-trained models (PPO actor / critic) can alias it.
    Returns (tensor, symmetric buffer or None)."""
    key = (dtype, str(device), id(symm_group) if symm_group is not None else None)
    ent = _GRAD_POOL.get(key)
    if ent is None or ent[0].numel() < numel:
        # A later, larger user (a bigger model, or the same model under a data-parallel degree with a coarser padding) gets a new
        # buffer that also serves everybody after it; earlier users keep the smaller one they were given (correct, just not shared)
        ent = _alloc_flat(numel, dtype, device, symm_group)
        _GRAD_POOL[key] = ent
    return ent[0][:numel], ent[1]


def _alloc_flat(numel: int, dtype, device, symm_group=None):
    """Zeroed flat buffer; with `symm_group` it lives in VMM symmetric memory with an NVSwitch multicast mapping."""
    if symm_group is None:
        return torch.zeros(numel, dtype=dtype, device=device), None
    from realhf_b200.parallel.symm_mem import VmmSymmetricBuffer
    sb = VmmSymmetricBuffer(numel * torch.tensor([], dtype=dtype).element_size(), group=symm_group, device=device)
    assert sb.mc_ptr != 0
    return sb.data(dtype=dtype)[:numel], sb


class _GradReady(torch.autograd.Function):
    """Identity on the activation entering layer `i`; its backward runs once every parameter gradient of layers >= i of this
    micro-batch has been enqueued, and tells the optimizer so (bucketed reduce-scatter overlapped with the rest of backward)."""

    @staticmethod
    def forward(ctx, x, cb, layer_idx):
        ctx.cb, ctx.layer_idx = cb, layer_idx
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        ctx.cb(ctx.layer_idx)
        return g, None, None


class FlatAdamW:
    """AdamW over a ReaLModel's flat parameter buffer, sharded across the data-parallel group (ZeRO-1).

    The flat buffer is cut into equal buckets (multiples of 64*dp elements, ~`bucket_numel`); rank r owns slice r of EVERY
    bucket.  During the backward pass of the last micro-batch a bucket is reduce-scattered as soon as the layers that write
    into it have finished (`grad_ready`), overlapped with the remaining backward compute (the reference overlaps through
    Megatron DDP's `overlap_grad_reduce`, backend/megatron.py:883-909).  Two transports:

      * NVLS (single NVSwitch domain): gradients and parameters live in VMM symmetric memory; the reduce-scatter is ONE kernel
        per bucket that pulls the in-switch sum of this rank's slice (`multimem.ld_reduce`) fused with the 1/dp average, the
        cast and the grad-norm statistics; after the global norm is known ONE kernel per bucket runs AdamW on the slice and
        stores the new weights with `multimem.st`, i.e. the parameter all-gather is the optimizer's own store (csrc/nvls.cu);
      * NCCL (any topology, also gloo on CPU): asynchronous in-place `reduce_scatter_tensor` per bucket during backward,
        fused AdamW per slice, in-place `all_gather_into_tensor` per bucket issued as soon as that bucket's update is queued.
    """

    def __init__(self, model, cfg: OptimizerConfig, total_steps: int = 1000):
        self.model, self.cfg = model, cfg
        self.ctx: ParallelContext = model.ctx
        self.sched = LRScheduler(cfg, total_steps)
        dev, n = model.device, model.flat_numel
        dp = self.ctx.dp_size
        assert n % 64 == 0
        align = 64 * dp  # every rank's slice of every bucket is 64-element aligned
        self.padded = (n + align - 1) // align * align
        self.shard_n = self.padded // dp
        self.grad_dtype = _DT[cfg.grad_dtype]
        self.state_dtype = _DT[cfg.state_dtype]
        pdt = model.dtype
        self.nvls = _nvls_usable(self.ctx, dev, cfg)
        # ---- buckets and this rank's slices
        bsz = self.padded
        if dp > 1 and cfg.zero_stage < 3 and cfg.bucket_numel > 0:
            bsz = max(align, cfg.bucket_numel // align * align)
        self.buckets: List[Tuple[int, int]] = []
        lo = 0
        while lo < self.padded:
            hi = min(self.padded, lo + bsz)
            if self.padded - hi < bsz // 4:  # no runt bucket at the end
                hi = self.padded
            self.buckets.append((lo, hi))
            lo = hi
        r = self.ctx.dp_rank
        self.ranges: List[Tuple[int, int]] = []   # my slice of every bucket (element offsets into the padded flat buffer)
        self.state_off: List[int] = []            # where that slice starts in m / v / master
        off = 0
        for (blo, bhi) in self.buckets:
            per = (bhi - blo) // dp
            self.ranges.append((blo + r * per, blo + (r + 1) * per))
            self.state_off.append(off)
            off += per
        assert off == self.shard_n
        self.lo, self.hi = self.ranges[0][0], min(n, self.ranges[-1][1])  # legacy names (exact with a single bucket)
        # ---- per-layer ZeRO-3: parameters / gradients / states exist only as per-layer shards (engine/zero3.py)
        self.z3 = None
        if cfg.zero_stage >= 3 and model.ctx.pp_size == 1 and os.environ.get("REAL_ZERO3_GRANULARITY", "layer") == "layer" \
                and model.config.resid_pdrop == 0 and model.config.attn_pdrop == 0 and model.config.embd_pdrop == 0 \
                and not model.sequence_parallel:
            from realhf_b200.engine.zero3 import Zero3Layers
            self.flat_grad, self._grad_symm = None, None
            self.z3 = Zero3Layers(self)
            self.shard_n = self.z3.shard_n
        else:
            # ---- gradient buffer
            sg = self.ctx.dp_group if self.nvls else None
            if cfg.share_grad_buffer:
                self.flat_grad, self._grad_symm = _grad_pool_get(self.padded, self.grad_dtype, dev, sg)
            else:
                self.flat_grad, self._grad_symm = _alloc_flat(self.padded, self.grad_dtype, dev, sg)
            model.attach_grad_buffer(self.flat_grad[:n])
        # ---- parameter storage: padded (in-place all-gather of the last bucket) and symmetric on the NVLS path
        self._param_symm = None
        self.param_store: Optional[torch.Tensor] = None
        if dp > 1 and cfg.zero_stage < 3 and (self.nvls or self.padded != n):
            self._rehome_params()
        self.use_master = cfg.use_master_weights and pdt != torch.float32 and self.state_dtype == torch.float32
        self.offload = cfg.offload and dev.type == "cuda"
        sdev = "cpu" if self.offload else dev
        pin = dict(pin_memory=True) if self.offload else {}
        self.m = torch.zeros(self.shard_n, dtype=self.state_dtype, device=sdev, **pin)
        self.v = torch.zeros(self.shard_n, dtype=self.state_dtype, device=sdev, **pin)
        self.master = None
        if self.use_master:
            self.master = torch.zeros(self.shard_n, dtype=torch.float32, device=sdev, **pin)
            if self.z3 is not None:
                self.master.copy_(self.z3.pshard.float())
            else:
                flat = model.flat_param.data
                for (a, b), so in zip(self.ranges, self.state_off):
                    b = min(b, n)
                    if b > a:
                        self.master[so: so + b - a].copy_(flat[a:b].float())
        self.step_count = 0
        self._stats = torch.zeros(2, dtype=torch.float32, device=dev)
        self._scale = torch.ones((), dtype=torch.float32, device=dev)
        self._skip = torch.zeros(1, dtype=torch.int32, device=dev)
        self.loss_scale = 1.0 if pdt != torch.float16 else float(min(cfg.initial_loss_scale, 2 ** 16))
        self._good_steps = 0
        self.last_grad_norm: Optional[torch.Tensor] = None
        # elements of replicated (non-TP-split) params in my slices: counted once (tp rank 0) in the global norm
        self._dup_idx: List[Optional[torch.Tensor]] = [None] * len(self.ranges)
        if self.ctx.tp_size > 1 and self.ctx.tp_rank != 0:
            for j, (lo_j, hi_j) in enumerate(self.ranges):
                idx = []
                for slot in model.slots.values():
                    if slot.spec.split_dim is None:
                        a, b = max(slot.offset, lo_j), min(slot.offset + slot.numel, hi_j)
                        if b > a:
                            idx.append(torch.arange(a - lo_j, b - lo_j))
                if idx:
                    self._dup_idx[j] = torch.cat(idx).to(dev)
        self._sp_sync = [s for s in model.slots.values() if s.spec.sp_grad_sync] if model.sequence_parallel else []
        # ---- overlap machinery
        self._layer_lo: Dict[int, int] = {}
        for slot in model.slots.values():
            li = int(slot.spec.name.split(".", 1)[0])
            self._layer_lo[li] = min(self._layer_lo.get(li, 1 << 62), slot.offset)
        self._next_bucket = len(self.buckets) - 1   # buckets are reduced from the end of the buffer (head) to the start
        self._pending: List = []                     # async NCCL works of this step
        self._comm_stream = torch.cuda.Stream(dev) if dev.type == "cuda" and dp > 1 else None
        self._overlap_ok = dp > 1 and len(self.buckets) > 1 and not self._sp_sync and not model.tied_embedding_params() \
            and os.environ.get("REAL_ZERO_OVERLAP", "1") == "1"
        self.n_overlapped = 0  # buckets whose reduce-scatter was issued from inside backward in the last step (diagnostics)

    # ------------------------------------------------------------------ storage
    def _rehome_params(self):
        """Move the model's flat parameters into optimizer-owned storage of `padded` elements (VMM symmetric memory with a
        multicast mapping on the NVLS path): the per-bucket all-gather then runs in place, without staging copies."""
        m = self.model
        n = m.flat_numel
        store, symm = _alloc_flat(self.padded, m.dtype, m.device, self.ctx.dp_group if self.nvls else None)
        with torch.no_grad():
            store[:n].copy_(m.flat_param.data)
        m.attach_flat(store[:n])
        if m.flat_grad is not None:
            m.attach_grad_buffer(self.flat_grad[:n])
        self.param_store, self._param_symm = store, symm

    def _store(self) -> torch.Tensor:
        """The padded parameter buffer the all-gather works on (the model's own buffer when no padding is needed)."""
        flat = self.model.flat_param.data
        if self.param_store is not None and self.param_store.data_ptr() == flat.data_ptr():
            return self.param_store
        if self.padded == flat.numel():
            return flat
        # somebody re-attached the model to a fresh buffer (recover / realloc into a trainable replica): adopt it again
        self.nvls_params_lost = True
        self._rehome_params()
        return self.param_store

    # ------------------------------------------------------------------ step
    def zero_grad(self):
        if self.z3 is not None:
            self.z3.zero_grad()
            return
        self.flat_grad.zero_()
        self._next_bucket = len(self.buckets) - 1
        self._pending = []
        self.n_overlapped = 0
        if self.nvls:
            self._stats.zero_()

    def scale_loss(self, loss: torch.Tensor) -> torch.Tensor:
        return loss * self.loss_scale if self.loss_scale != 1.0 else loss

    def boundary(self, x: torch.Tensor, layer_idx: int) -> torch.Tensor:
        """Called by the model on the activation entering layer `layer_idx` while `grad_ready` hooks are armed."""
        return _GradReady.apply(x, self.grad_ready, layer_idx)

    def arm(self, on: bool):
        """Arm / disarm the in-backward reduce-scatter (armed for the last micro-batch of a train_batch only)."""
        self.model._grad_boundary = self.boundary if (on and self._overlap_ok and self.model.ctx.pp_size == 1 and self.z3 is None) else None

    def grad_ready(self, layer_idx: int):
        """All parameter gradients of layers >= layer_idx are final (and enqueued): reduce every bucket that lies entirely
        at or above the first parameter of that layer."""
        ready_lo = self._layer_lo.get(layer_idx)
        if ready_lo is None:
            return
        while self._next_bucket >= 0 and self.buckets[self._next_bucket][0] >= ready_lo:
            self._reduce_bucket(self._next_bucket, overlapped=True)
            self._next_bucket -= 1

    def _reduce_bucket(self, k: int, overlapped: bool = False):
        ctx = self.ctx
        blo, bhi = self.buckets[k]
        lo, hi = self.ranges[k]
        self.n_overlapped += int(overlapped)
        if self.nvls:
            es = self.flat_grad.element_size()
            cs = self._comm_stream
            cs.wait_stream(torch.cuda.current_stream(self.flat_grad.device))
            with torch.cuda.stream(cs):
                self._grad_symm.reduce_scatter_(lo * es, (hi - lo) * es, self.grad_dtype, 1.0 / ctx.dp_size, self._stats)
            return
        bucket = self.flat_grad[blo:bhi]
        if self.flat_grad.is_cuda:
            w = dist.reduce_scatter_tensor(self.flat_grad[lo:hi], bucket, op=dist.ReduceOp.AVG, group=ctx.dp_group, async_op=True)
            self._pending.append(w)
        else:  # gloo has no AVG / reduce_scatter_tensor: all-reduce the bucket, keep my slice
            dist.all_reduce(bucket, group=ctx.dp_group)
            self.flat_grad[lo:hi].div_(ctx.dp_size)

    def _sync_grads(self):
        """Finish the gradient reduction: afterwards flat_grad[ranges[k]] holds the DP-averaged gradient of my slices."""
        ctx = self.ctx
        if self._sp_sync and ctx.tp_size > 1:  # norm params see only T/tp tokens under sequence parallelism
            bufs = [self.model.flat_grad[s.offset: s.offset + s.numel] for s in self._sp_sync]
            flat = torch.cat([b.float() for b in bufs])
            dist.all_reduce(flat, group=ctx.tp_group)
            off = 0
            for b in bufs:
                b.copy_(flat[off: off + b.numel()].to(b.dtype))
                off += b.numel()
        tied = self.model.tied_embedding_params()
        if tied and ctx.embedding_group is not None:  # the two copies of a tied embedding see different gradients
            for p in tied:
                g = p.grad if p.grad is not None else getattr(p, "main_grad", None)
                if g is not None:
                    dist.all_reduce(g, group=ctx.embedding_group)
        if ctx.dp_size == 1:
            return
        while self._next_bucket >= 0:  # whatever backward did not hand over itself (embedding bucket, pipeline runs, ...)
            self._reduce_bucket(self._next_bucket)
            self._next_bucket -= 1
        for w in self._pending:
            w.wait()
        self._pending = []
        if self.nvls:
            torch.cuda.current_stream(self.flat_grad.device).wait_stream(self._comm_stream)

    def _step_zero3(self, version_steps: Optional[int]) -> Dict[str, torch.Tensor]:
        cfg, ctx = self.cfg, self.ctx
        z = self.z3
        z.finish_grads()
        self._stats.zero_()
        z.grad_sumsq(self._stats)
        if ctx.model_group is not None and ctx.topo.world_size() > 1:
            dist.all_reduce(self._stats, group=ctx.model_group)
        inv_ls = 1.0 / self.loss_scale
        norm = self._stats[0].sqrt() * inv_ls
        self.last_grad_norm = norm
        clip = cfg.gradient_clipping
        coef = torch.clamp(clip / (norm + 1e-6), max=1.0) if clip and clip > 0 else torch.ones_like(norm)
        self._scale.copy_(coef * inv_ls)
        self._skip.copy_((self._stats[1:2] > 0).int())
        self.step_count += 1
        if version_steps is not None:
            self.sched.step_absolute(version_steps)
        lr = self.sched.get_lr()
        z.step(lr)  # the updated shard IS the parameter store: the next gather picks the new values up, no all-gather here
        if self.model.dtype == torch.float16:
            self._update_loss_scale()
        return {"grad_norm": norm, "lr": torch.tensor(lr), "skipped": self._skip.float()[0]}

    def step(self, version_steps: Optional[int] = None) -> Dict[str, torch.Tensor]:
        if self.z3 is not None:
            return self._step_zero3(version_steps)
        cfg, ctx = self.cfg, self.ctx
        self.arm(False)
        self._sync_grads()
        n = self.model.flat_numel
        # ---- global grad norm + overflow detection, all on device (the NVLS reduce-scatter kernels accumulated them already)
        if not self.nvls:
            self._stats.zero_()
            for (lo, hi) in self.ranges:
                OF.sumsq_accum(self.flat_grad[lo:hi], self._stats)
        for (lo, hi), dup in zip(self.ranges, self._dup_idx):
            if dup is not None:
                self._stats[0] -= self.flat_grad[lo:hi][dup].float().pow(2).sum()
        if ctx.model_group is not None and ctx.topo.world_size() > 1:
            dist.all_reduce(self._stats, group=ctx.model_group)
        inv_ls = 1.0 / self.loss_scale
        norm = self._stats[0].sqrt() * inv_ls
        self.last_grad_norm = norm
        clip = cfg.gradient_clipping
        coef = torch.clamp(clip / (norm + 1e-6), max=1.0) if clip and clip > 0 else torch.ones_like(norm)
        self._scale.copy_(coef * inv_ls)
        self._skip.copy_((self._stats[1:2] > 0).int())
        # ---- AdamW on my slices + parameter all-gather, bucket by bucket
        self.step_count += 1
        if version_steps is not None:
            self.sched.step_absolute(version_steps)
        lr = self.sched.get_lr()
        store = self._store() if ctx.dp_size > 1 and cfg.zero_stage < 3 else self.model.flat_param.data
        seed = self.step_count * 2654435761 % (2 ** 31)
        works = []
        for k, ((blo, bhi), (lo, hi), so) in enumerate(zip(self.buckets, self.ranges, self.state_off)):
            hi = min(hi, store.numel())  # an unpadded store (ZeRO-3 / dp == 1) ends before the last slice's padding
            n_my = max(0, hi - lo)
            g = self.flat_grad[lo:hi]
            mst = self.master[so: so + n_my] if self.master is not None else None
            stoch = self.state_dtype == torch.bfloat16
            if self.nvls:
                sb = self._param_symm
                lib().nvls_adam_allgather(sb.data_ptrs, sb.pad_ptrs, sb.mc_ptr, lo * store.element_size(), 0 if store.dtype == torch.float32 else 1,
                                          g, self.m[so: so + n_my], self.v[so: so + n_my], mst, n_my, lr, cfg.beta1, cfg.beta2, cfg.eps,
                                          cfg.weight_decay, self.step_count, self._scale, self._skip, stoch, (seed + 7919 * k) % (2 ** 31),
                                          sb.rank)
                continue
            if n_my > 0:
                if self.offload:
                    self._offloaded_update(store[lo:hi], g, lr, so)
                else:
                    OF.adamw_step(store[lo:hi], g, self.m[so: so + n_my], self.v[so: so + n_my], mst, lr, cfg.beta1, cfg.beta2, cfg.eps,
                                  cfg.weight_decay, self.step_count, self._scale, self._skip, stochastic=stoch,
                                  seed=(seed + 7919 * k) % (2 ** 31))
            if ctx.dp_size > 1 and cfg.zero_stage < 3:
                if store.is_cuda:  # in place: my slice already sits at its final position; overlaps the next bucket's update
                    works.append(dist.all_gather_into_tensor(store[blo:bhi], store[lo:hi], group=ctx.dp_group, async_op=True))
                else:
                    parts = [torch.empty(n_my, dtype=store.dtype) for _ in range(ctx.dp_size)]
                    dist.all_gather(parts, store[lo:hi].contiguous(), group=ctx.dp_group)
                    store[blo:bhi].copy_(torch.cat(parts))
        for w in works:
            w.wait()
        if self.model.dtype == torch.float16:
            self._update_loss_scale()
        return {"grad_norm": norm, "lr": torch.tensor(lr), "skipped": self._skip.float()[0]}

    def _offloaded_update(self, p_shard, g, lr, so: int = 0, chunk: int = 1 << 26):
        """Stream pinned-host optimizer state through the GPU in chunks, double-buffered on a side stream.
        `so`: offset of this slice in the state tensors."""
        cfg = self.cfg
        n = p_shard.numel()
        dev = p_shard.device
        if not hasattr(self, "_h2d"):
            self._h2d, self._d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        main = torch.cuda.current_stream(dev)
        offs = list(range(0, n, chunk))
        sm, sv = self.m[so: so + n], self.v[so: so + n]
        smaster = self.master[so: so + n] if self.master is not None else None

        def fetch(a):
            b = min(n, a + chunk)
            with torch.cuda.stream(self._h2d):
                bufs = [t[a:b].to(dev, non_blocking=True) for t in (sm, sv)]
                bufs.append(smaster[a:b].to(dev, non_blocking=True) if smaster is not None else None)
                ev = torch.cuda.Event()
                ev.record(self._h2d)
            return a, b, bufs, ev

        nxt = fetch(offs[0])
        for i, a in enumerate(offs):
            a, b, (m, v, ms), ev = nxt
            if i + 1 < len(offs):
                nxt = fetch(offs[i + 1])
            main.wait_event(ev)
            OF.adamw_step(p_shard[a:b], g[a:b], m, v, ms, lr, cfg.beta1, cfg.beta2, cfg.eps, cfg.weight_decay,
                          self.step_count, self._scale, self._skip)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(self._d2h):
                self._d2h.wait_event(done)
                sm[a:b].copy_(m, non_blocking=True)
                sv[a:b].copy_(v, non_blocking=True)
                if ms is not None:
                    smaster[a:b].copy_(ms, non_blocking=True)
                for t in (m, v, ms):
                    if t is not None:
                        t.record_stream(self._d2h)
        main.wait_stream(self._d2h)

    def _update_loss_scale(self):
        skipped = bool(self._skip.item())  # fp16 only: the one host sync dynamic loss scaling needs
        c = self.cfg
        if skipped:
            # `hysteresis` overflowing steps are tolerated (each one is skipped) before the scale is halved; the credit refills
            # when the scale grows again (DeepSpeed / Megatron dynamic loss scaler semantics)
            self._hysteresis_left = getattr(self, "_hysteresis_left", c.hysteresis) - 1
            if self._hysteresis_left <= 0:
                self.loss_scale = max(c.min_loss_scale, self.loss_scale / 2)
                self._hysteresis_left = c.hysteresis
            self._good_steps = 0
        else:
            self._good_steps += 1
            if self._good_steps >= c.loss_scale_window:
                self.loss_scale *= 2
                self._good_steps = 0
                self._hysteresis_left = c.hysteresis

    # ------------------------------------------------------------------ ZeRO-3: parameters sharded between calls
    # Between model function calls only this rank's 1/dp slice of the flat parameter buffer stays resident (on the GPU,
    # or in pinned host memory with `offload_param`); `materialize()` all-gathers the full buffer before a call and
    # `release()` drops it afterwards.  ZeRO-3 uses ONE bucket (ranges[0] is the classic contiguous shard).
    def begin_call(self, train: bool = False):
        """Start of an engine call on a ZeRO-3 model: per-layer streaming when available, whole-model gather otherwise."""
        if self.z3 is not None:
            if self.model.instantiated:   # somebody materialised the full buffer (generation / save): drop it first
                self.release()
            self.z3.begin_call(train)
        else:
            self.materialize()

    def end_call(self):
        if self.z3 is not None:
            self.z3.end_call()
        else:
            self.release()

    def release(self):
        if self.z3 is not None:
            if self.model.instantiated:
                if getattr(self, "_full_dirty", False):  # the full buffer was written (checkpoint load, realloc into this model)
                    self.z3.absorb_full(self.model.flat_param.data)
                    self._full_dirty = False
                self.model.detach_params()
            return
        if self.cfg.zero_stage < 3 or not self.model.instantiated:
            return
        flat = self.model.flat_param.data
        shard = flat[self.lo: self.hi]
        if self.cfg.offload_param and flat.is_cuda:
            if getattr(self, "_param_shard_host", None) is None:
                self._param_shard_host = torch.empty(self.hi - self.lo, dtype=flat.dtype, device="cpu", pin_memory=True)
            self._param_shard_host.copy_(shard, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            self._param_shard = None
        else:
            self._param_shard = shard.clone()
        self._dev = flat.device
        self.model.release_params()

    def materialize(self):
        if self.z3 is not None:
            if not self.model.instantiated:
                self.z3.end_call()
                self.model.attach_flat(self.z3.materialize_full())
            return
        if self.cfg.zero_stage < 3 or self.model.instantiated:
            return
        n = self.model.flat_numel
        buf = torch.empty(self.padded, dtype=self.model.dtype, device=self._dev)
        src = self._param_shard if self._param_shard is not None else self._param_shard_host
        buf[self.lo: self.hi].copy_(src, non_blocking=True)
        if self.ctx.dp_size > 1:
            mine = buf[self.lo: self.lo + self.shard_n]
            if buf.is_cuda:
                dist.all_gather_into_tensor(buf, mine, group=self.ctx.dp_group)  # in place
            else:
                parts = [torch.empty_like(mine) for _ in range(self.ctx.dp_size)]
                dist.all_gather(parts, mine.contiguous(), group=self.ctx.dp_group)
                buf.copy_(torch.cat(parts))
        self.model.attach_flat(buf[:n])
        self.model.attach_grad_buffer(self.flat_grad[:n])
        self._param_shard = None

    # ------------------------------------------------------------------ state (recover saves what the reference drops)
    def state_dict(self):
        return {"m": self.m.cpu(), "v": self.v.cpu(), "master": None if self.master is None else self.master.cpu(),
                "step": self.step_count, "sched": self.sched.state_dict(), "loss_scale": self.loss_scale,
                "shard": (self.lo, self.hi), "ranges": list(self.ranges)}

    def load_state_dict(self, sd):
        assert [tuple(r) for r in sd.get("ranges", [sd["shard"]])] == [tuple(r) for r in self.ranges] or \
            (len(self.ranges) == 1 and tuple(sd["shard"]) == (self.lo, self.hi)), "optimizer shard layout changed"
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
        if self.master is not None and sd["master"] is not None:
            self.master.copy_(sd["master"])
        self.step_count = sd["step"]
        self.sched.load_state_dict(sd["sched"])
        self.loss_scale = sd["loss_scale"]
