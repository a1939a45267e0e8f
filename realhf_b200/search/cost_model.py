"""Cost model of the allocation search: time and memory of one MFC under one (dp, tp, pp) layout, and the time of a parameter
reallocation between two layouts.

Parity: `realhf/search_engine/estimate.py` (per-instruction costs from profiled layer statistics :57-360, memory :377-450) and
`realhf/search_engine/param_realloc.py` (cost of a realloc between two layouts :19-110, tables :111-307).  Differences, by design:

  * the time of a layer is read from the layer profiler's table (`search/layers.py`: device-timed embedding / block / head /
    decode / optimizer rows on this machine) by interpolation over the token count of the nearest profiled sequence length;
    without a table every quantity falls back to the roofline formulas of `search/engine.py::estimate`, fed by
    `MEASURED_PEAKS.json`.  The reference needs an exact (batch size, sequence length) hit in its statistics files;
  * collectives follow a latency + bytes / bandwidth model whose constants come from measurements of this repo's own kernels
    (`profiles/nvls_collectives_*.json`: NVSwitch-multicast all-reduce / reduce-scatter / all-gather) instead of a hard-coded table
    of NCCL numbers for another machine;
  * the reallocation cost is not a formula over layouts: the REAL planner (`parallel/realloc.py::derive_plan`) is run on the two
    layouts and its transfers are costed per GPU -- local segment copies at HBM speed, peer stores at NVLink speed, the slowest GPU
    decides.  Plans of equal geometry are cached.
"""

from __future__ import annotations

import bisect
import dataclasses
import json
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

from realhf_b200.api.config import ModelInterfaceType
from realhf_b200.api.model import ReaLModelConfig

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


# --------------------------------------------------------------------------------------------------- profiled layer times


class ProfileTable:
    """Rows `{layer, op, bs, seqlen, time_us}` of the layer profiler, queried by (layer, op, tokens, seqlen).

    layer: embedding | block | head | optimizer;  op: fwd | fwd_bwd | decode | step.
    For `decode` rows `bs` is the number of sequences and `seqlen` the context length; for the optimizer row `bs` is the number of
    parameters (in millions) the step was timed on."""

    def __init__(self, rows: Sequence[Dict]):
        self.rows = list(rows)
        self.meta: Dict = {}
        self.path: Optional[str] = None
        self._idx: Dict[Tuple[str, str], Dict[int, List[Tuple[float, float]]]] = {}
        for r in self.rows:
            n = float(r["bs"]) if r["op"] in ("decode", "step") else float(r["bs"]) * float(r["seqlen"])
            self._idx.setdefault((r["layer"], r["op"]), {}).setdefault(int(r.get("seqlen", 0)), []).append((n, float(r["time_us"])))
        for per_len in self._idx.values():
            for pts in per_len.values():
                pts.sort()

    @classmethod
    def load(cls, path: str) -> "ProfileTable":
        with open(path) as f:
            d = json.load(f)
        t = cls(d["rows"] if isinstance(d, dict) else d)
        t.meta = d.get("meta", {}) if isinstance(d, dict) else {}
        t.path = path
        return t

    @classmethod
    def find(cls, model_name: str) -> Optional["ProfileTable"]:
        """`$REAL_PROFILE_TABLE`, else the profiler's cache entry of this model (`apps/profile_layers.py` writes it), else a table
        shipped with the package (`search/tables/<model>.json`, calibrated on B200 from end-to-end measurements)."""
        from realhf_b200.base import constants
        for p in (os.environ.get("REAL_PROFILE_TABLE", ""), os.path.join(constants.PROFILER_CACHE_PATH, f"layers_{model_name}.json"),
                  os.path.join(os.path.dirname(os.path.abspath(__file__)), "tables", f"{model_name}.json")):
            if p and os.path.exists(p):
                return cls.load(p)
        return None

    def has(self, layer: str, op: str) -> bool:
        return (layer, op) in self._idx

    def time_us(self, layer: str, op: str, n: float, seqlen: int) -> Optional[float]:
        """Interpolated time for `n` tokens (sequences for decode rows) at the profiled sequence length nearest to `seqlen`."""
        per_len = self._idx.get((layer, op))
        if not per_len:
            return None
        sl = min(per_len, key=lambda s: abs(math.log(max(s, 1)) - math.log(max(seqlen, 1))))
        pts = per_len[sl]
        xs = [p[0] for p in pts]
        if n <= xs[0]:
            # below the smallest profiled size a kernel chain stops shrinking: never less than a quarter of the smallest measurement
            return pts[0][1] * max(n / xs[0], 0.25)
        if n >= xs[-1]:
            if len(pts) >= 2 and xs[-1] > xs[-2]:
                slope = (pts[-1][1] - pts[-2][1]) / (xs[-1] - xs[-2])
                slope = max(slope, 0.5 * pts[-1][1] / xs[-1])            # a flat tail (latency-bound points) must not extrapolate flat
                return pts[-1][1] + slope * (n - xs[-1])
            return pts[-1][1] * n / xs[-1]
        i = bisect.bisect_right(xs, n)
        (x0, y0), (x1, y1) = pts[i - 1], pts[i]
        return y0 + (y1 - y0) * (n - x0) / (x1 - x0)


# --------------------------------------------------------------------------------------------------------- collectives


@dataclasses.dataclass
class CommModel:
    """time = latency + bytes moved per GPU / bandwidth.  Defaults: measured on 8 x B200 with this repo's kernels
    (`profiles/nvls_collectives_8gpu.json`: 64 KiB all-reduce 16 us, 512 MiB reduce-scatter 760 us, all-gather 782 us)."""

    ar_latency_us: float = 16.0          # small all-reduce (one-shot multimem.ld_reduce), isolated
    decode_ar_us: float = 40.0           # the same all-reduce in situ, fused into the RMSNorm of a graph-replayed TP decode layer: the
                                         # two-way epoch barrier waits for the slowest peer's GEMM (profiles/bench_n8_r2_gentp8.json)
    coll_latency_us: float = 30.0        # reduce-scatter / all-gather launch + barrier
    p2p_latency_us: float = 12.0
    bus_bw: float = 6.2e11               # bytes/s per GPU, large reduce-scatter / all-gather through the switch
    p2p_bw: float = 7.7e11               # peer copy, per direction
    inter_node_bw: float = 5.0e10        # per GPU (400 Gb/s NIC)

    @classmethod
    def from_measured(cls, world: int = 8) -> "CommModel":
        cm = cls()
        p = os.path.join(_ROOT, "profiles", f"nvls_collectives_{world}gpu.json")
        if not os.path.exists(p):
            return cm
        try:
            ranks = json.load(open(p))["per_rank"]
            small = max(min(r["64"].get("nvls_1shot_us", 1e9), r["64"].get("nccl_us", 1e9)) for r in ranks)
            z = [r["zero_512MiB"] for r in ranks if "zero_512MiB" in r]
            if small < 1e9:
                cm.ar_latency_us = small
            if z:
                rs = max(min(x.get("nvls_rs_us", 1e9), x.get("nccl_rs_us", 1e9)) for x in z)
                moved = (512 << 20) * (world - 1) / world
                cm.bus_bw = moved / ((rs - cm.coll_latency_us) * 1e-6)
        except (OSError, ValueError, KeyError, TypeError):
            pass
        return cm

    def _bw(self, n: int, gpus_per_node: int) -> float:
        return self.bus_bw if n <= gpus_per_node else self.inter_node_bw

    def all_reduce_us(self, nbytes: float, n: int, gpus_per_node: int = 8) -> float:
        if n <= 1:
            return 0.0
        return self.ar_latency_us + 2.0 * nbytes * (n - 1) / n / self._bw(n, gpus_per_node) * 1e6

    def reduce_scatter_us(self, nbytes: float, n: int, gpus_per_node: int = 8) -> float:
        """`nbytes`: the full (un-scattered) buffer."""
        if n <= 1:
            return 0.0
        return self.coll_latency_us + nbytes * (n - 1) / n / self._bw(n, gpus_per_node) * 1e6

    all_gather_us = reduce_scatter_us

    def p2p_us(self, nbytes: float, same_node: bool = True) -> float:
        return self.p2p_latency_us + nbytes / (self.p2p_bw if same_node else self.inter_node_bw) * 1e6


# ------------------------------------------------------------------------------------------------------------- helpers


def config_from_shape(shape: Dict[str, float], is_critic: bool = False) -> ReaLModelConfig:
    """A LLaMA-style `ReaLModelConfig` for the {h, L, f, v} shapes the search works with (used to run the realloc planner)."""
    h = int(shape["h"])
    nh = max(1, h // 128)
    return ReaLModelConfig(n_layers=int(shape["L"]), n_kv_heads=int(shape.get("kv", nh)), n_q_heads=nh, hidden_dim=h,
                           intermediate_dim=int(shape["f"]), vocab_size=int(shape["v"]), n_positions=4096, embd_pdrop=0.0,
                           resid_pdrop=0.0, attn_pdrop=0.0, activation_function="silu", scale_attn_by_inverse_layer_idx=False,
                           use_attention_bias=False, use_attn_proj_bias=False, use_mlp_bias=False, layer_norm_type="rms",
                           mlp_type="llama", apply_rotary=True, is_critic=is_critic)


def _block_params(shape) -> float:
    return 4 * shape["h"] ** 2 + 3 * shape["h"] * shape["f"]


@dataclasses.dataclass
class MFCCost:
    time_us: float
    mem_static: float          # bytes / GPU while the role lives on this layout (weights, gradients, optimizer state)
    mem_active: float          # bytes / GPU during the call only (activations, KV cache, logits)
    breakdown: Dict[str, float] = dataclasses.field(default_factory=dict)


# ------------------------------------------------------------------------------------------------------ time of an MFC


def estimate_mfc(kind: ModelInterfaceType, n_seqs: int, shape: Dict[str, float], dp: int, tp: int, pp: int, hw, table: Optional[ProfileTable],
                 comm: Optional[CommModel] = None, seq_len: int = 128, gen_len: int = 256, n_minibatches: int = 1, n_mbs: int = 1,
                 trainable_role: bool = False, use_sequence_parallel: bool = False, zero_stage: int = 1, gpus_per_node: int = 8,
                 optimizer_bytes_per_param: float = 12.0, gradient_checkpointing: bool = True) -> MFCCost:
    """Time and memory of one MFC call over `n_seqs` sequences of `seq_len` prompt + `gen_len` generated tokens.

    The table holds tp = 1 times.  A TP rank executes 1/tp of every GEMM's flops on all of its dp group's tokens, which costs what
    tokens/tp cost at tp = 1 (the fixed per-kernel part is kept by the table's small-size behaviour), plus the layer-boundary
    collectives.  A pipeline of pp stages runs m micro-batches in (m + pp - 1) slots of one stage-micro-batch each."""
    comm = comm or CommModel()
    h, L, v, n = shape["h"], shape["L"], shape["v"], shape["n"]
    total_len = seq_len + gen_len
    bs = max(1.0, n_seqs / dp)                      # sequences per dp rank
    tokens = bs * total_len
    layers_per_stage = L / pp
    wbytes = 2.0 * n / (tp * pp)
    bd: Dict[str, float] = {}

    def blk(op: str, toks: float, sl: int) -> float:
        """One block on one GPU for `toks` tokens of the dp rank."""
        t = table.time_us("block", op, toks / tp, sl) if table is not None else None
        if t is None:
            mult = {"fwd": 2.0, "fwd_bwd": 6.0 + (2.0 if gradient_checkpointing else 0.0)}[op]
            flops = mult * _block_params(shape) * toks / tp + mult * 2 * toks * sl * h / tp      # GEMMs + attention scores/values
            t = flops / (hw.bf16_flops * hw.gemm_eff) * 1e6 + 8 * hw.launch_us
        return t

    def edge(op: str, toks: float, sl: int) -> float:
        """Embedding + head (first / last stage)."""
        t = None
        if table is not None:
            te = table.time_us("embedding", op if table.has("embedding", op) else "fwd", toks / tp, sl)
            th = table.time_us("head", op, toks / tp, sl) if table.has("head", op) else 0.0
            if te is not None:
                t = te * (1.0 if table.has("embedding", op) or op == "fwd" else 3.0) + (th or 0.0)
        if t is None:
            mult = 2.0 if op == "fwd" else 6.0
            t = mult * v * h * toks / tp / (hw.bf16_flops * hw.gemm_eff) * 1e6 + 4 * hw.launch_us
        return t

    def tp_coll(toks: float, n_per_layer: int) -> float:
        """Layer-boundary collectives of one stage: all-reduce of [toks, h] bf16 (reduce-scatter + all-gather under SP: same bytes)."""
        if tp == 1:
            return 0.0
        return layers_per_stage * n_per_layer * comm.all_reduce_us(2.0 * toks * h, tp, gpus_per_node)

    def pipeline(stage_mb_us: float, m: int, act_bytes: float) -> float:
        if pp == 1:
            return stage_mb_us * m
        return (m + pp - 1) * (stage_mb_us + comm.p2p_us(act_bytes, same_node=tp * pp <= gpus_per_node))

    if kind == ModelInterfaceType.TRAIN_STEP:
        mb_tokens = tokens / n_minibatches
        m = max(n_mbs, pp * 2 if pp > 1 else n_mbs)
        t_mb = tokens / n_minibatches / m
        stage = layers_per_stage * blk("fwd_bwd", t_mb, total_len) + edge("fwd_bwd", t_mb, total_len) / pp + tp_coll(t_mb, 4)
        compute = n_minibatches * pipeline(stage, m, 2.0 * t_mb * h / (tp if use_sequence_parallel else 1))
        shard = n / (tp * pp)
        opt_t = None
        if table is not None and table.has("optimizer", "step"):
            opt_t = table.time_us("optimizer", "step", shard / dp / 1e6, 0)
        if opt_t is None:
            opt_t = (optimizer_bytes_per_param + 4.0) * shard / dp / hw.hbm_bw * 1e6 + 3 * hw.launch_us
        grad_sync = comm.reduce_scatter_us(2.0 * shard, dp, gpus_per_node) + comm.all_gather_us(2.0 * shard, dp, gpus_per_node)
        # gradient buckets are reduced inside the backward of the last micro-batch: only the tail bucket and the parameter gather are exposed
        exposed = 0.15 * comm.reduce_scatter_us(2.0 * shard, dp, gpus_per_node) + comm.all_gather_us(2.0 * shard, dp, gpus_per_node)
        t = compute + n_minibatches * (opt_t + exposed)
        bd.update(compute=compute, optimizer=n_minibatches * opt_t, grad_sync_exposed=n_minibatches * exposed, grad_sync_total=n_minibatches * grad_sync)
        opt_shard = dp if zero_stage >= 1 else 1
        grad_shard = dp if zero_stage >= 2 else 1
        param_shard = dp if zero_stage >= 3 else 1
        static = wbytes / param_shard + 2.0 * shard / grad_shard + optimizer_bytes_per_param * shard / opt_shard
        per_layer_act = (2.0 if gradient_checkpointing else 34.0) * mb_tokens / m * h / (tp if use_sequence_parallel else 1)
        active = per_layer_act * layers_per_stage * (pp if pp > 1 else 1) + 34.0 * mb_tokens / m * h / tp + 8.0 * min(mb_tokens / m, 4096) * v / tp
    elif kind == ModelInterfaceType.GENERATE:
        m = pp                                              # the token ring keeps pp micro-batches in flight
        mb_bs = bs / m
        prefill_stage = layers_per_stage * blk("fwd", mb_bs * seq_len, seq_len) + edge("fwd", mb_bs * seq_len, seq_len) / pp + tp_coll(mb_bs * seq_len, 2)
        prefill = pipeline(prefill_stage, m, 2.0 * mb_bs * seq_len * h)
        kv_per_tok = 2.0 * 2.0 * shape.get("kv_h", h) * L / (tp * pp)
        ctx = seq_len + gen_len / 2.0
        dec = table.time_us("block", "decode", mb_bs, int(ctx)) if (table is not None and table.has("block", "decode")) else None
        if dec is not None:
            # profiled at tp = 1: the weight-streaming part shrinks with tp, the per-kernel fixed part (~45 us per layer on B200) does not
            fixed = min(dec, 45.0)
            dec = fixed + (dec - fixed) / tp
        else:
            dec = (2.0 * _block_params(shape) / tp) / (0.85 * hw.hbm_bw) * 1e6 + mb_bs * ctx * kv_per_tok / L * pp / (0.87 * hw.hbm_bw) * 1e6 + 45.0
        head = table.time_us("head", "decode", mb_bs, int(ctx)) if (table is not None and table.has("head", "decode")) else None
        if head is None:
            head = 2.0 * v * h / tp / (0.85 * hw.hbm_bw) * 1e6 + 30.0
        step_stage = layers_per_stage * dec + head / pp
        if tp > 1:
            step_stage += layers_per_stage * 2 * max(comm.decode_ar_us, comm.all_reduce_us(2.0 * mb_bs * h, tp, gpus_per_node)) \
                + comm.all_gather_us(2.0 * mb_bs * v, tp, gpus_per_node)
        step = m * step_stage + (pp * comm.p2p_us(2.0 * mb_bs * h) if pp > 1 else 0.0)
        t = prefill + gen_len * step
        bd.update(prefill=prefill, decode_step=step)
        static = 0.0 if trainable_role else wbytes
        active = bs * total_len * kv_per_tok + (wbytes if trainable_role else 0.0) + 4.0 * bs * v / tp
    else:
        m = max(n_mbs, pp)
        t_mb = tokens / m
        stage = layers_per_stage * blk("fwd", t_mb, total_len) + edge("fwd", t_mb, total_len) / pp + tp_coll(t_mb, 2)
        t = pipeline(stage, m, 2.0 * t_mb * h)
        bd.update(compute=t)
        static = 0.0 if trainable_role else wbytes
        active = 6.0 * t_mb * h / tp + 8.0 * min(t_mb, 4096) * v / tp + (wbytes if trainable_role else 0.0)
    return MFCCost(t, static, active, bd)


# ------------------------------------------------------------------------------------- parameter reallocation between layouts


_PLAN_BYTES_CACHE: Dict[Tuple, Dict[str, Dict[int, float]]] = {}


def _plan_bytes(cfg: ReaLModelConfig, src: Tuple[int, int, int], src_ranks: Sequence[int], dst: Tuple[int, int, int],
                dst_ranks: Sequence[int]) -> Dict[str, Dict[int, float]]:
    """Per-GPU bytes of the planner's transfers: local copies, bytes sent to peers, bytes received from peers (bf16)."""
    from realhf_b200.base.topology import ProcessTopology
    from realhf_b200.parallel.realloc import derive_plan
    key = (cfg.n_layers, cfg.hidden_dim, cfg.intermediate_dim, cfg.vocab_size, cfg.n_q_heads, cfg.n_kv_heads, cfg.is_critic, src,
           tuple(src_ranks), dst, tuple(dst_ranks))
    hit = _PLAN_BYTES_CACHE.get(key)
    if hit is not None:
        return hit
    (s_dp, s_tp, s_pp), (d_dp, d_tp, d_pp) = src, dst
    vols = derive_plan(cfg, ProcessTopology(s_pp, s_dp, s_tp), list(src_ranks), ProcessTopology(d_pp, d_dp, d_tp), list(dst_ranks),
                       volumes_only=True)
    out = {"local": {}, "send": {}, "recv": {}}
    for (sw, dw), numel in vols.items():
        b = 2.0 * numel
        if sw == dw:
            out["local"][sw] = out["local"].get(sw, 0.0) + b
        else:
            out["send"][sw] = out["send"].get(sw, 0.0) + b
            out["recv"][dw] = out["recv"].get(dw, 0.0) + b
    _PLAN_BYTES_CACHE[key] = out
    return out


def realloc_time_us(cfg: ReaLModelConfig, src: Tuple[int, int, int], src_ranks: Sequence[int], dst: Tuple[int, int, int],
                    dst_ranks: Sequence[int], hw, comm: Optional[CommModel] = None, gpus_per_node: int = 8) -> float:
    """Time of `src` (dp, tp, pp on GPUs `src_ranks`) -> `dst` reallocation of one model: the planner's transfers, costed per GPU.
    A local segment copy reads and writes HBM; a peer store leaves through the sender's NVLink port and enters through the
    receiver's; GPUs on different nodes share the NIC bandwidth.  The slowest GPU decides."""
    comm = comm or CommModel()
    if tuple(src) == tuple(dst) and list(src_ranks) == list(dst_ranks):
        return 0.0
    b = _plan_bytes(cfg, tuple(src), src_ranks, tuple(dst), dst_ranks)
    if not b["send"] and not b["recv"]:
        d_dp, d_tp, d_pp = dst
        if d_tp == src[1] and d_pp == src[2]:
            # every destination GPU already holds exactly its shard (same tp / pp position, only the dp degree or the set of GPUs
            # differs): the runtime aliases the source's flat buffer, nothing is copied (system/model_worker.py::_param_realloc)
            return comm.coll_latency_us
    gpus = set(b["local"]) | set(b["send"]) | set(b["recv"])
    cross_node = len({g // gpus_per_node for g in list(src_ranks) + list(dst_ranks)}) > 1
    link = comm.inter_node_bw if cross_node else comm.p2p_bw
    worst = 0.0
    for g in gpus:
        t_local = 2.0 * b["local"].get(g, 0.0) / hw.hbm_bw
        t_link = max(b["send"].get(g, 0.0), b["recv"].get(g, 0.0)) / link
        worst = max(worst, t_local + t_link)
    return comm.coll_latency_us + worst * 1e6
