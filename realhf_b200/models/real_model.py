"""ReaLModel: a packed variable-length decoder-only transformer whose parameters live in ONE flat buffer.

Feature matrix of the reference model (`realhf/impl/model/nn/real_llm_api.py`, `real_llm_base.py`,
`modules/{attn,mlp,rotary,embedding}.py`): MHA/GQA, RoPE (+linear/dynamic scaling) or learned absolute
positions, LayerNorm / RMSNorm / Gemma-RMSNorm, GELU MLP / gated (SwiGLU, GeGLU) MLP / MoE, tied
embeddings, critic (scalar) head, per-block activation checkpointing, tensor + sequence parallelism and
pipeline stages (a stage holds a contiguous range of layer indices: 0 = embedding, 1..L = blocks,
L+1 = head).

Design: blocks are *functional* — they read parameter views out of `self.flat_param`, so parameter
reallocation, ZeRO sharding, host offload and checkpoint I/O all act on one contiguous tensor per shard.
All projections are fused (one QKV GEMM, one gate|up GEMM); RoPE is applied in place on the QKV output,
the gated activation reads the fused gate|up output once, and inference fuses residual-add into RMSNorm.
"""

from __future__ import annotations

import dataclasses
import math
import zlib
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint

from realhf_b200.api.model import ReaLModelConfig
from realhf_b200.base.topology import ParallelContext
from realhf_b200.models import sharding
from realhf_b200.models.sharding import ParamSpec
from realhf_b200.ops import attention as attn_ops
from realhf_b200.ops import functional as OF
from realhf_b200.parallel import tp as TP


def _numel(shape) -> int:
    n = 1
    for d in shape:
        n *= d
    return n


_ALIGN = 64  # every parameter starts on a 64-element boundary (128 B for bf16): vector kernels + TMA


@dataclasses.dataclass
class ParamSlot:
    spec: ParamSpec
    shape: Tuple[int, ...]  # local (TP-sharded) shape
    offset: int             # element offset in the flat buffer
    numel: int


def build_layout(cfg: ReaLModelConfig, layers: Sequence[int], tp_size: int) -> Tuple[Dict[str, ParamSlot], int]:
    """Flat-buffer layout of one shard: name -> slot, and the padded total element count."""
    slots: Dict[str, ParamSlot] = {}
    off = 0
    for spec in sharding.model_param_specs(cfg, layers):
        shp = sharding.shard_shape(spec, cfg, tp_size)
        n = _numel(shp)
        slots[spec.name] = ParamSlot(spec, shp, off, n)
        off += (n + _ALIGN - 1) // _ALIGN * _ALIGN
    return slots, off


@dataclasses.dataclass
class ModelOutput:
    """What a forward returns to a loss / post-processing function.

    `hidden` is the final-norm output [T, H].  Log-probs should be taken through `logprobs()` (fused LM-head +
    online softmax: the [T, V] logits are never materialised); `logits` builds them on demand for user code
    that wants them (and for critics, where it is the [T] value head output)."""

    hidden: torch.Tensor
    head_weight: Optional[torch.Tensor]
    ctx: Optional[ParallelContext]
    is_critic: bool = False
    _logits: Optional[torch.Tensor] = None

    @property
    def logits(self) -> torch.Tensor:
        if self._logits is None:
            if self.is_critic:
                self._logits = F.linear(self.hidden, self.head_weight).squeeze(-1)
            else:
                tp = self.ctx is not None and self.ctx.tp_size > 1
                lg = OF.linear(TP.copy_to_tp(self.hidden, self.ctx) if tp else self.hidden, self.head_weight)
                self._logits = TP.gather_last_dim(lg, self.ctx) if tp else lg
        return self._logits

    @property
    def values(self) -> torch.Tensor:
        assert self.is_critic
        return self.logits.float()

    def logprobs(self, labels: torch.Tensor, mask_bits: Optional[torch.Tensor] = None, temperature: float = 1.0,
                 rows: Optional[torch.Tensor] = None) -> torch.Tensor:
        """fp32 log p(labels[i] | hidden[rows[i]]) (rows defaults to all tokens)."""
        h = self.hidden if rows is None else self.hidden.index_select(0, rows)
        if self.ctx is not None and self.ctx.tp_size > 1:
            local = OF.linear(TP.copy_to_tp(h, self.ctx), self.head_weight)  # column-parallel head: d(hidden) is summed over TP
            return TP.vocab_parallel_logprobs(local, labels, self.ctx, temperature, mask_bits)
        return OF.lm_head_logprobs(h, self.head_weight, labels, mask_bits, temperature)


class ReaLModel(nn.Module):
    def __init__(self, config: ReaLModelConfig, ctx: Optional[ParallelContext] = None, dtype=torch.bfloat16,
                 device="cpu", layer_range: Optional[Tuple[int, int]] = None):
        super().__init__()
        if not config.do_layernorm_before:
            # the option exists in the reference's config for OPT-350m-style post-LN blocks; none of the registered model families
            # uses it, and silently running pre-LN instead would load such a checkpoint into the wrong network
            raise NotImplementedError("do_layernorm_before=False (post-LN blocks) is not supported")
        self.config = config
        self.ctx = ctx if ctx is not None else ParallelContext.single()
        self.dtype = dtype
        self.device = torch.device(device)
        if layer_range is None:
            layer_range = sharding.partition_pipeline_layers(config, self.ctx.pp_size)[self.ctx.pp_rank]
        self.layer_range = layer_range
        self.layers = list(range(*layer_range))
        self.slots, self.flat_numel = build_layout(config, self.layers, self.ctx.tp_size)
        self.flat_param: Optional[nn.Parameter] = None
        self.flat_grad: Optional[torch.Tensor] = None
        self.p: Dict[str, torch.Tensor] = {}
        self._rope: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
        self._rope_len = 0
        self.gradient_checkpointing = bool(self.ctx.gradient_checkpointing)
        # "auto": recompute only as many blocks as the free HBM of this GPU requires (see _n_unckpt_blocks)
        self.ckpt_auto = self.ctx.gradient_checkpointing == "auto"
        self.ckpt_margin_bytes: Optional[int] = None
        self.last_unckpt_blocks = 0
        self.sequence_parallel = bool(self.ctx.sequence_parallel) and self.ctx.tp_size > 1
        self._shared_gens: Dict[torch.device, torch.Generator] = {}
        self._offloaded: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ construction
    @property
    def is_first_stage(self) -> bool:
        return self.layer_range[0] == 0

    @property
    def is_last_stage(self) -> bool:
        return self.layer_range[1] == self.config.n_layers + 2

    @property
    def instantiated(self) -> bool:
        return self.flat_param is not None and not getattr(self, "_params_detached", False)

    def detach_params(self):
        """Drop the flat buffer but KEEP the parameter objects (per-layer ZeRO-3 re-points their `.data` at gathered scratch
        buffers layer by layer; hooks and requires_grad flags live on the objects)."""
        if self.flat_param is None:
            return
        empty = torch.empty(0, dtype=self.dtype, device=self.device)
        self.flat_param.data = empty
        for v in self.p.values():
            v.data = empty
            v.grad = None
        self.flat_grad = None
        self._params_detached = True

    def instantiate(self, init: str = "random", std: float = 0.02, seed: Optional[int] = None) -> "ReaLModel":
        """Allocate the flat buffer and carve parameter views.  init: random | empty."""
        flat = torch.empty(self.flat_numel, dtype=self.dtype, device=self.device)
        self.attach_flat(flat)
        if init == "random":
            gen = torch.Generator(device="cpu")
            gen.manual_seed(seed if seed is not None else 1)
            with torch.no_grad():
                flat.zero_()
                for name, slot in self.slots.items():
                    if slot.spec.init == "ones":
                        self.p[name].fill_(1.0)
                    elif slot.spec.init == "normal":
                        # draw the FULL tensor from a per-parameter stream, then shard it: every (pp, tp) layout
                        # of the same seed holds the same weights
                        seed_name = "0.wte.weight" if (name.endswith("head.weight") and self.config.tied_embedding) else name
                        gen.manual_seed(((seed if seed is not None else 1) * 1000003 + zlib.crc32(seed_name.encode())) % (2 ** 63))
                        full = torch.empty(slot.spec.shape, dtype=torch.float32).normal_(0.0, std, generator=gen)
                        sh = sharding.shard_tensor(slot.spec, self.config, full, self.ctx.tp_rank, self.ctx.tp_size)
                        self.p[name].copy_(sh.to(self.dtype))
        return self

    def init_random_fast(self, std: float = 0.02, seed: int = 1):
        """Device-side random init for benchmarks (no full-tensor host materialisation)."""
        flat = torch.empty(self.flat_numel, dtype=self.dtype, device=self.device)
        self.attach_flat(flat)
        g = torch.Generator(device=self.device)
        g.manual_seed(seed + 1000 * self.ctx.tp_rank + 77 * self.ctx.pp_rank)
        g_rep = torch.Generator(device=self.device)  # TP-replicated tensors (MoE router, position embeddings, biases of row-
        g_rep.manual_seed(seed + 77 * self.ctx.pp_rank + 13)  # parallel layers, critic head) must be EQUAL on the ranks of a TP group
        with torch.no_grad():
            flat.normal_(0.0, std, generator=g)
            for name, slot in self.slots.items():
                if slot.spec.init == "ones":
                    self.p[name].fill_(1.0)
                elif slot.spec.init == "zeros":
                    self.p[name].zero_()
                elif self.ctx.tp_size > 1 and slot.spec.split_dim is None:
                    self.p[name].normal_(0.0, std, generator=g_rep)
        return self

    def attach_flat(self, flat: torch.Tensor):
        """(Re)point every parameter view at `flat` (used by instantiate, realloc and offload reload)."""
        assert flat.numel() == self.flat_numel, (flat.numel(), self.flat_numel)
        self._params_detached = False
        if self.flat_param is None:
            self.flat_param = nn.Parameter(flat, requires_grad=False)
        else:
            self.flat_param.data = flat
        for name, slot in self.slots.items():
            view = flat[slot.offset: slot.offset + slot.numel].view(slot.shape)
            if name in self.p:
                self.p[name].data = view
            else:
                self.p[name] = nn.Parameter(view, requires_grad=True)
        self.device = flat.device

    def release_params(self):
        """Drop the flat buffer (reverse direction of a realloc / after offload)."""
        if self.flat_param is None:
            return
        empty = torch.empty(0, dtype=self.dtype, device=self.device)
        self.flat_param.data = empty
        for v in self.p.values():
            v.data = empty
        self.flat_param = None
        self.p = {}

    def parameters(self, recurse: bool = True):
        return iter(self.p.values())

    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True):
        return iter(self.p.items())

    def attach_grad_buffer(self, flat_grad: torch.Tensor):
        """Point every `.grad` into one flat gradient buffer (same layout as the params)."""
        assert flat_grad.numel() == self.flat_numel
        self.flat_grad = flat_grad
        for name, slot in self.slots.items():
            g = flat_grad[slot.offset: slot.offset + slot.numel].view(slot.shape)
            if flat_grad.dtype == self.p[name].dtype:
                self.p[name].grad = g
                self.p[name]._grad_in_flat_buffer = True  # lets the GEMM wgrad accumulate into it in place
            else:
                p = self.p[name]
                p.main_grad = g  # fp32 bucket: the GEMM wgrad kernel accumulates into it directly
                if not getattr(p, "_main_grad_hooked", False):
                    def _fold(param):  # params whose grad comes from autograd (norms, embeddings, biases)
                        if param.grad is None:  # the GEMM wgrad already accumulated into main_grad and returned no gradient
                            return
                        param.main_grad.add_(param.grad.view_as(param.main_grad))
                        param.grad = None
                    p.register_post_accumulate_grad_hook(_fold)
                    p._main_grad_hooked = True

    def state_dict(self, *a, **k) -> Dict[str, torch.Tensor]:
        return {n: t.data for n, t in self.p.items()}

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        missing = [k for k in self.p if k not in sd]
        unexpected = [k for k in sd if k not in self.p]
        if strict and (missing or unexpected):
            raise KeyError(f"load_state_dict: missing={missing} unexpected={unexpected}")
        with torch.no_grad():
            for k, v in sd.items():
                if k in self.p:
                    self.p[k].copy_(v.to(self.dtype))
        return missing, unexpected

    # ------------------------------------------------------------------ host offload (non-trainable roles)
    def offload(self, frozen: bool = False):
        """D2H of the whole shard into one pinned buffer, then drop the device copy (reference: real_llm_api.py:274-306).
        frozen=True: the weights never change (reference / reward models), so after the first call the host copy is still
        valid and offloading is just freeing the device buffer."""
        if self.flat_param is None or not self.flat_param.is_cuda:
            return
        if self._offloaded is None or not frozen:
            if self._offloaded is None:
                self._offloaded = torch.empty(self.flat_numel, dtype=self.dtype, device="cpu", pin_memory=True)
            self._offloaded.copy_(self.flat_param.data, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        dev = self.flat_param.device
        self.release_params()
        self._offload_dev = dev

    def reload(self, stream: Optional["torch.cuda.Stream"] = None):
        """H2D of an offloaded shard.  With `stream` the copy runs there (the caller makes its compute stream wait on it before
        the first use, e.g. `SPMDExecutor` hooks): the reload of an idle model overlaps another model's generation."""
        if self._offloaded is None or self.flat_param is not None:
            return
        if stream is None:
            flat = torch.empty(self.flat_numel, dtype=self.dtype, device=self._offload_dev)
            flat.copy_(self._offloaded, non_blocking=True)
        else:
            stream.wait_stream(torch.cuda.current_stream(self._offload_dev))  # blocks freed by the compute stream are reusable
            with torch.cuda.stream(stream):
                flat = torch.empty(self.flat_numel, dtype=self.dtype, device=self._offload_dev)
                flat.copy_(self._offloaded, non_blocking=True)
        self.attach_flat(flat)
        for p in self.p.values():
            p.requires_grad_(False)

    # ------------------------------------------------------------------ helpers
    def _w(self, name: str) -> Optional[torch.Tensor]:
        return self.p.get(name)

    def rope_tables(self, need_len: int):
        c = self.config
        if self._rope is None or self._rope_len < need_len or self._rope[0].device != self.device:
            n = max(need_len, c.n_positions or 0, 2048)
            n = 1 << (n - 1).bit_length()
            self._rope = OF.rope_tables(n, c.head_dim, c.rotary_base, self.device, c.rotary_scaling, c.rotary_scaling_type)
            self._rope_len = n
        return self._rope

    def _norm(self, x, prefix: str):
        c = self.config
        w = self.p[f"{prefix}.weight"]
        if c.layer_norm_type is None:
            return OF.layer_norm(x, w, self.p[f"{prefix}.bias"], c.layer_norm_epsilon)
        return OF.rmsnorm(x, w, c.layer_norm_epsilon, 1.0 if c.layer_norm_type == "gemma" else 0.0)

    def _local_heads(self) -> Tuple[int, int]:
        c, t = self.config, self.ctx.tp_size
        return c.n_q_heads // t, max(1, c.n_kv_heads // t)

    def _attn_scale(self, layer_idx: int) -> float:
        c = self.config
        s = 1.0 / math.sqrt(c.head_dim) if c.scale_attn_weights else 1.0
        if c.scale_attn_by_inverse_layer_idx:
            s /= float(layer_idx)
        return s

    # ------------------------------------------------------------------ layers
    def _embed(self, input_ids, position_ids):
        c = self.config
        x = TP.vocab_parallel_embedding(input_ids, self.p["0.wte.weight"], self.ctx, sp=self.sequence_parallel)
        if not c.apply_rotary:
            pe = F.embedding(position_ids.long() + c.abs_position_embedding_offset, self.p["0.wpe.weight"])
            if self.sequence_parallel:
                pe = TP._split_first_dim(pe, self.ctx)
            x = x + pe
        if c.normalize_embed:
            x = x * torch.tensor(c.hidden_dim ** 0.5, dtype=x.dtype, device=x.device)
        return self._dropout(x, c.embd_pdrop)

    # ------------------------------------------------------------------ dropout under tensor parallelism
    def shared_generator(self, device) -> torch.Generator:
        """A generator whose stream is identical on all TP ranks of this replica (same dp / pp coordinates) and different
        elsewhere: randomness applied to activations that are REPLICATED over the TP group (residual / embedding dropout
        without sequence parallelism, CPU sampling) must agree across the group, while each worker's global generator is
        seeded per rank.  (Reference: the model-parallel RNG tracker, impl/model/utils/random.py:76-286.)"""
        device = torch.device(device)
        g = self._shared_gens.get(device)
        if g is None:
            from realhf_b200.base import seeding
            g = torch.Generator(device=device)
            g.manual_seed(seeding.derive_seed("tp-shared", self.ctx.dp_rank, self.ctx.pp_rank))
            self._shared_gens[device] = g
        return g

    def _dropout(self, x, p: float):
        if not self.training or p <= 0:
            return x
        if self.ctx.tp_size == 1 or self.sequence_parallel:
            return F.dropout(x, p)  # x is rank-local (or token-sharded): rank-local randomness is what we want
        keep = torch.empty_like(x).bernoulli_(1.0 - p, generator=self.shared_generator(x.device))
        return x * keep * (1.0 / (1.0 - p))

    def _ckpt_block(self, i, x, d, position_ids, cu_seqlens, max_seqlen):
        """Activation checkpointing of one block.  torch's checkpoint replays the GLOBAL generators; the TP-shared generator
        is ours to replay: the recomputation must see the state the first pass saw, and must not disturb the live state."""
        uses_shared = self.training and self.ctx.tp_size > 1 and not self.sequence_parallel and self.config.resid_pdrop > 0
        if not uses_shared:
            return checkpoint(self._block_packed, i, x, d, position_ids, cu_seqlens, max_seqlen, use_reentrant=False)
        g = self.shared_generator(x.device)
        at_forward = g.get_state()
        calls = [0]

        def run(x_, d_):
            calls[0] += 1
            if calls[0] == 1:
                return self._block_packed(i, x_, d_, position_ids, cu_seqlens, max_seqlen)
            live = g.get_state()
            g.set_state(at_forward)
            try:
                return self._block_packed(i, x_, d_, position_ids, cu_seqlens, max_seqlen)
            finally:
                g.set_state(live)
        return checkpoint(run, x, d, use_reentrant=False)

    def _attention_packed(self, i: int, h, position_ids, cu_seqlens, max_seqlen, kv_sink: Optional[list]):
        """Attention branch of block i on the already-normalised input h."""
        c = self.config
        nq, nkv = self._local_heads()
        hd = c.head_dim
        qkv = TP.col_linear(h, self.p[f"{i}.attn.qkv.weight"], self._w(f"{i}.attn.qkv.bias"), self.ctx, self.sequence_parallel)
        if c.apply_rotary:
            cos, sin = self.rope_tables(max_seqlen)
            qkv = OF.apply_rope(qkv, cos, sin, position_ids, nq + nkv, hd, hd, c.rotary_interleaved)
        T = qkv.shape[0]
        if kv_sink is not None:
            kv_sink.append((qkv[:, nq * hd:(nq + nkv) * hd].view(T, nkv, hd), qkv[:, (nq + nkv) * hd:].view(T, nkv, hd)))
        o = attn_ops.varlen_attention_qkv(qkv, cu_seqlens, max_seqlen, nq, nkv, hd, self._attn_scale(i), True,
                                          c.attn_pdrop if self.training else 0.0)
        o = TP.row_linear(o.reshape(T, nq * hd), self.p[f"{i}.attn.o.weight"], self._w(f"{i}.attn.o.bias"), self.ctx,
                          self.sequence_parallel)
        return self._dropout(o, c.resid_pdrop)

    def _mlp(self, i: int, x, h=None):
        """h: the already-normalised input (decode path fuses residual add + norm), else computed from x."""
        c = self.config
        if h is None:
            h = self._norm(x, f"{i}.mlp.ln")
        sp = self.sequence_parallel
        if c.mlp_type == "llama":
            if not sp:  # gated activation fused into the gate|up GEMM epilogue (the SP path fuses the all-gather instead)
                hh = TP.copy_to_tp(h, self.ctx) if self.ctx.tp_size > 1 else h
                a = OF.gated_linear(hh, self.p[f"{i}.mlp.gate_up.weight"], c.activation_function)
            else:
                gu = TP.col_linear(h, self.p[f"{i}.mlp.gate_up.weight"], None, self.ctx, sp)
                a = OF.gated_act(gu, c.activation_function)
            o = TP.row_linear(a, self.p[f"{i}.mlp.down.weight"], None, self.ctx, sp)
        elif c.mlp_type == "moe":
            from realhf_b200.models import moe
            o = moe.moe_forward(self, i, h)
        else:
            a = TP.col_linear(h, self.p[f"{i}.mlp.fc.weight"], self._w(f"{i}.mlp.fc.bias"), self.ctx, sp)
            a = _ACT[c.activation_function](a)
            o = TP.row_linear(a, self.p[f"{i}.mlp.proj.weight"], self._w(f"{i}.mlp.proj.bias"), self.ctx, sp)
        return self._dropout(o, c.resid_pdrop)

    def _add_norm(self, x, d, prefix: str):
        """(normalised activations, residual stream) at a layer boundary.  `d` is the previous branch's output that has not been
        added to the residual stream yet: for RMSNorm models the add is fused into the norm kernel (forward AND backward,
        `OF.add_rmsnorm`), which removes every eager residual add / gradient-accumulation add from a training step."""
        c = self.config
        if d is None:
            return self._norm(x, prefix), x
        if c.layer_norm_type is not None:
            return OF.add_rmsnorm(d, x, self.p[f"{prefix}.weight"], c.layer_norm_epsilon, 1.0 if c.layer_norm_type == "gemma" else 0.0)
        x = x + d
        return self._norm(x, prefix), x

    def _block_packed(self, i: int, x, d, position_ids, cu_seqlens, max_seqlen, kv_sink=None):
        """One transformer block on the packed batch.  Takes and returns (residual stream, pending branch output)."""
        h, x = self._add_norm(x, d, f"{i}.attn.ln")
        o = self._attention_packed(i, h, position_ids, cu_seqlens, max_seqlen, kv_sink)
        h2, x = self._add_norm(x, o, f"{i}.mlp.ln")
        d = self._mlp(i, x, h2)
        if i == self.config.n_layers:
            x, _ = self._add_norm(x, d, f"{i}.ln_f")
            d = None
        return x, d

    def head_weight(self) -> Optional[torch.Tensor]:
        c = self.config
        L1 = c.n_layers + 1
        if f"{L1}.head.weight" in self.p:
            return self.p[f"{L1}.head.weight"]
        if c.tied_embedding:
            return self.p.get("0.wte.weight")
        return None

    # ---- W8A8 decode (ops/fp8.py): e4m3 copies of the local linear weights, used by `decode_step` / the LM head while set
    _fp8 = None
    _fp8_active = False  # the copies may outlive a call (kept CUDA graph) but are only READ between enable and disable

    def fp8_decode_supported(self) -> bool:
        c = self.config
        from realhf_b200.ops import fp8
        on_dev = self.device.type == "cuda" and self.dtype in (torch.bfloat16, torch.float16)
        return bool((on_dev or (self.device.type != "cuda" and fp8.emulate())) and self.ctx.tp_size == 1 and c.mlp_type == "llama"
                    and self.instantiated)

    def enable_fp8_decode(self) -> int:
        """Quantise (or, when the buffers exist already, re-quantise in place) every local block linear and the LM head.
        Returns the bytes held by the e4m3 copies.  ~4 ms for a 7B model: done at the start of every generation call."""
        from realhf_b200.ops import fp8
        names = []
        for i in self.layers:
            if 1 <= i <= self.config.n_layers:
                names += [f"{i}.attn.qkv.weight", f"{i}.attn.o.weight", f"{i}.mlp.gate_up.weight", f"{i}.mlp.down.weight"]
        ws = {n: self.p[n] for n in names if n in self.p}
        if self.is_last_stage and self.head_weight() is not None:
            ws["head"] = self.head_weight()
        old = self._fp8 or {}
        new = {}
        for n, w in ws.items():
            if not fp8.supported(w):
                continue
            f = old.get(n)
            if f is not None and (f.N, f.K) == tuple(w.shape) and f.qw.device == w.device:
                f.requantize(w)
            else:
                f = fp8.Fp8Linear(w)
            new[n] = f
        self._fp8 = new
        self._fp8_active = True
        return sum(f.nbytes() for f in new.values())

    def disable_fp8_decode(self, free: bool = True):
        self._fp8_active = False
        if free:
            self._fp8 = None

    def tied_embedding_params(self) -> List[torch.Tensor]:
        """Parameters whose gradients must be summed over the embedding group (tied embeddings with pp > 1)."""
        c = self.config
        if not c.tied_embedding or self.ctx.pp_size == 1:
            return []
        if self.is_first_stage:
            return [self.p["0.wte.weight"]]
        if self.is_last_stage:
            return [self.p[f"{c.n_layers + 1}.head.weight"]]
        return []

    # ------------------------------------------------------------------ forward (packed)
    def forward(self, input_ids: Optional[torch.Tensor] = None, cu_seqlens: Optional[torch.Tensor] = None,
                max_seqlen: Optional[int] = None, hidden: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.Tensor] = None, kv_sink: Optional[list] = None) -> ModelOutput:
        """Packed forward over this stage's layers.

        First stage: `input_ids` [T]; later stages: `hidden` [T(/tp), H].  cu_seqlens: int32 [B+1].
        With sequence parallelism the caller (engine) has padded T to a multiple of tp.
        Returns a ModelOutput on the last stage, else the hidden-state tensor for the next stage.
        """
        c = self.config
        if self.flat_param is None and self._offloaded is not None:
            self.reload()  # offloaded by an OffloadHook after its last call: bring the weights back on first use
        cu_seqlens = cu_seqlens.int()
        if max_seqlen is None:
            max_seqlen = int((cu_seqlens[1:] - cu_seqlens[:-1]).max())
        T = int(input_ids.shape[0]) if input_ids is not None else None
        if position_ids is None:
            position_ids = packed_position_ids(cu_seqlens, T if T is not None else int(cu_seqlens[-1]))
        x = hidden
        ckpt = self.gradient_checkpointing and torch.is_grad_enabled() and kv_sink is None
        n_keep = self._n_unckpt_blocks(T if T is not None else int(hidden.shape[0])) if ckpt else 0
        first_kept = c.n_layers + 1 - n_keep  # the LAST n_keep blocks keep their activations (any subset would do)
        # armed by the optimizer for the last micro-batch of a step: marks where the gradients of all later layers are final,
        # so their buckets are reduce-scattered while the earlier layers are still in backward (engine/optim.py)
        gb = getattr(self, "_grad_boundary", None) if torch.is_grad_enabled() else None
        d = None  # branch output not yet folded into the residual stream (see _add_norm)
        for i in self.layers:
            if i == 0:
                x = self._embed(input_ids, position_ids)
            elif i <= c.n_layers:
                if gb is not None and x.requires_grad:
                    x = gb(x, i)
                z3 = getattr(self, "_zero3", None)
                if z3 is not None and z3.streaming:
                    # per-layer ZeRO-3: gather this block's parameters, run it without keeping activations, release (engine/zero3.py)
                    x, d = z3.run_block(i, lambda x_, d_, i=i: self._block_packed(i, x_, d_, position_ids, cu_seqlens, max_seqlen), x, d)
                elif ckpt and i < first_kept:
                    x, d = self._ckpt_block(i, x, d, position_ids, cu_seqlens, max_seqlen)
                else:
                    x, d = self._block_packed(i, x, d, position_ids, cu_seqlens, max_seqlen, kv_sink)
        if d is not None:  # a pipeline stage that does not end with the final norm hands on the summed stream
            x = x + d
        if not self.is_last_stage:
            return x
        if gb is not None and x.requires_grad:
            x = gb(x, c.n_layers + 1)
        if self.sequence_parallel:
            x = TP.gather_from_sp(x, self.ctx, reduce_scatter_bwd=False)
        return ModelOutput(hidden=x, head_weight=self.head_weight(), ctx=self.ctx, is_critic=c.is_critic)

    def _n_unckpt_blocks(self, n_tokens: int) -> int:
        """How many transformer blocks of this forward pass may keep their activations instead of being recomputed.

        The reference checkpoints every block (`nn/real_llm_base.py:194-204`); with 180 GB per GPU that wastes a quarter of
        the training FLOPs whenever the model states leave room.  In "auto" mode the budget is the HBM that is free right
        now (driver-free + cached-but-unused allocator blocks) minus a safety margin (max(24 GB, 15%): backward transients, logits
        chunks, workspace), divided by the saved-tensor footprint
        of one block for this micro-batch: x, two normalised inputs, qkv (+ its rotated copy), attention output and LSE,
        gate_up and the gated activation = (11 + 3 F/H) * T * H * 2 bytes."""
        if not self.ckpt_auto or self.device.type != "cuda":
            return 0
        c = self.config
        free, total = torch.cuda.mem_get_info(self.device)
        cached = torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
        margin = self.ckpt_margin_bytes if self.ckpt_margin_bytes is not None else max(24 << 30, (total * 3) // 20)
        tp = max(1, self.ctx.tp_size)
        inter = c.intermediate_dim if c.mlp_type != "moe" else c.intermediate_dim * max(1, getattr(c.moe, "top_k", 1))
        per_block = int(1.1 * n_tokens * c.hidden_dim * 2 * (5 + (6 + 3 * inter / c.hidden_dim) / tp))  # +10%: allocator rounding
        n = max(0, int((free + cached - margin) // max(per_block, 1)))
        self.last_unckpt_blocks = min(n, self.n_local_blocks())
        return self.last_unckpt_blocks

    # ------------------------------------------------------------------ decode step (one token per sequence)
    @torch.no_grad()
    def decode_step(self, input_ids: Optional[torch.Tensor], k_caches: List[torch.Tensor], v_caches: List[torch.Tensor],
                    cache_lens: torch.Tensor, hidden: Optional[torch.Tensor] = None) -> torch.Tensor:
        """input_ids [B] -> final hidden [B, H] (last stage: after ln_f).  Appends K/V at `cache_lens` in place.
        Caches: one [B, S, nkv_local, hd] tensor per local block.  Capturable in a CUDA graph (no host sync)."""
        c = self.config
        nq, nkv = self._local_heads()
        hd = c.head_dim
        x = hidden
        cos = sin = None
        if c.apply_rotary:
            cos, sin = self.rope_tables(k_caches[0].shape[1] if k_caches else 2048)
        li = 0
        rms = c.layer_norm_type is not None
        w_off = 1.0 if c.layer_norm_type == "gemma" else 0.0
        eps = c.layer_norm_epsilon
        # Tensor-parallel decode over NVSwitch multicast memory: the row-parallel GEMMs (o, down) write their partial sums into
        # symmetric memory and the all-reduce happens INSIDE the next residual-add + RMSNorm kernel (`multimem.ld_reduce`), so a
        # TP layer has exactly the kernels of a single-GPU layer (parallel/fused_tp.py::ar_add_rmsnorm)
        fused = getattr(self.ctx, "symm", None) if self.ctx.tp_size > 1 else None
        tp_fused = fused is not None and getattr(fused, "nvls", False) and rms and c.mlp_type == "llama" and not c.use_attn_proj_bias \
            and self.dtype in (torch.bfloat16, torch.float16)
        d = None      # branch output not yet added to the residual stream: every add is fused into the next RMSNorm kernel
        d_sym = None  # same, but still a per-rank partial sum in symmetric memory (tensor-parallel decode)
        # W8A8 decode: e4m3 weights + per-token e4m3 activations on tcgen05 kind::f8f6f4 (`enable_fp8_decode`, <= 128 rows)
        fp8 = self._fp8 if (self._fp8_active and self._fp8 and self.ctx.tp_size == 1) else None

        def q8(h_, wname, bname=None):
            f = fp8.get(wname) if fp8 is not None and h_.shape[0] <= 128 else None
            return None if f is None else f(h_, self._w(bname) if bname else None)

        def norm_q8(x_, ln_name, wname, bname=None):
            """Layer boundary whose only consumer is an fp8 GEMM: residual add + RMSNorm + e4m3 quantisation in one kernel,
            then the GEMM.  Returns (GEMM output, new residual stream) or None (then the unfused path runs)."""
            nonlocal d
            f = fp8.get(wname) if fp8 is not None and rms and x_.shape[0] <= 128 and d_sym is None else None
            if f is None:
                return None
            from realhf_b200.ops import fp8 as F8
            r = F8.add_rmsnorm_quant(d, x_, self.p[ln_name], eps, w_off)
            if r is None:
                return None
            d = None
            return f.gemm_q(r[0], r[1], self._w(bname) if bname else None, out_dtype=x_.dtype), r[2]

        def add_norm(x_, wname):
            """(normalised input, new residual stream) at a layer boundary, consuming the pending branch output."""
            nonlocal d, d_sym
            w_ = self.p[wname]
            if d_sym is not None:
                h_, x_ = fused.ar_add_rmsnorm(d_sym, x_, w_, eps, w_off)
            elif rms and d is not None:
                h_, x_ = OF.add_rmsnorm(d, x_, w_, eps, w_off)
            else:
                if d is not None:
                    x_ = x_ + d
                h_ = self._norm(x_, wname[: -len(".weight")])
            d = d_sym = None
            return h_, x_

        for i in self.layers:
            if i == 0:
                x = self._embed(input_ids, cache_lens)
            elif i <= c.n_layers:
                fq = norm_q8(x, f"{i}.attn.ln.weight", f"{i}.attn.qkv.weight", f"{i}.attn.qkv.bias")
                if fq is not None:
                    qkv, x = fq
                else:
                    h, x = add_norm(x, f"{i}.attn.ln.weight")
                    qkv = TP.col_linear(h, self.p[f"{i}.attn.qkv.weight"], self._w(f"{i}.attn.qkv.bias"), self.ctx, False)
                o = attn_ops.decode_attention(qkv, k_caches[li], v_caches[li], cache_lens, nq, nkv, hd, self._attn_scale(i),
                                              cos, sin, hd, c.rotary_interleaved)
                li += 1
                if tp_fused:
                    d_sym = fused.gemm_partial(o, self.p[f"{i}.attn.o.weight"])
                if d_sym is None:
                    d = q8(o, f"{i}.attn.o.weight", f"{i}.attn.o.bias")
                    if d is None:
                        d = TP.row_linear(o, self.p[f"{i}.attn.o.weight"], self._w(f"{i}.attn.o.bias"), self.ctx, False)
                fq = norm_q8(x, f"{i}.mlp.ln.weight", f"{i}.mlp.gate_up.weight") if fp8 is not None and f"{i}.mlp.down.weight" in fp8 else None
                if fq is not None:  # W8A8 MLP: norm+quant -> gate|up GEMM -> act+quant -> down GEMM
                    from realhf_b200.ops import fp8 as F8
                    gu8, x = fq
                    aq = F8.gated_act_quant(gu8, c.activation_function)
                    if aq is not None:
                        d = fp8[f"{i}.mlp.down.weight"].gemm_q(aq[0], aq[1], out_dtype=gu8.dtype)
                    else:
                        d = q8(OF.gated_act(gu8, c.activation_function), f"{i}.mlp.down.weight")
                elif tp_fused:
                    h2, x = add_norm(x, f"{i}.mlp.ln.weight")
                    gu = TP.col_linear(h2, self.p[f"{i}.mlp.gate_up.weight"], None, self.ctx, False)
                    d_sym = fused.gemm_partial(OF.gated_act(gu, c.activation_function), self.p[f"{i}.mlp.down.weight"])
                    if d_sym is None:
                        d = TP.row_linear(OF.gated_act(gu, c.activation_function), self.p[f"{i}.mlp.down.weight"], None, self.ctx, False)
                else:
                    h2, x = add_norm(x, f"{i}.mlp.ln.weight")
                    d = self._mlp(i, x, h2)
                if i == c.n_layers:
                    x, _ = add_norm(x, f"{i}.ln_f.weight")
        if d_sym is not None:  # a pipeline stage that does not end with ln_f hands on the summed residual stream
            x = fused.ar_add_rmsnorm(d_sym, x, None)
        elif d is not None:
            x = x + d
        if fused is not None and hasattr(fused, "align_parity"):
            fused.align_parity()
        return x

    def n_local_blocks(self) -> int:
        return sum(1 for i in self.layers if 1 <= i <= self.config.n_layers)


def packed_position_ids(cu_seqlens: torch.Tensor, total: int) -> torch.Tensor:
    """int32 [T]: position of every packed token inside its sequence, computed without a host sync."""
    dev = cu_seqlens.device
    idx = torch.arange(total, device=dev, dtype=torch.int32)
    seq = torch.searchsorted(cu_seqlens[1:].contiguous(), idx, right=True)
    seq = seq.clamp_(max=cu_seqlens.numel() - 2)
    return (idx - cu_seqlens[seq.long()]).int()


_ACT = {
    "gelu": F.gelu,
    "gelu_new": lambda x: F.gelu(x, approximate="tanh"),
    "gelu_pytorch_tanh": lambda x: F.gelu(x, approximate="tanh"),
    "relu": F.relu,
    "silu": F.silu,
}
