"""Allocation search driver: enumerate (device mesh, dp/tp/pp) candidates per MFC, cost them with the B200 model,
let the native MCMC + simulator (`_C/host_ext`: `multi_mcmc_search`) pick one per MFC.

Parity: `realhf/search_engine/{search,enumerate,estimate,param_realloc}.py` + `csrc/search`.  The reference's Python
driver is broken against its own DFG API; this one is exercised by `tests/test_search.py`.
Costs come from the layer profiler's table (`search/layers.py` -> `search/cost_model.py::estimate_mfc`) when the model has one,
else from the analytic roofline model below fed by the measured peaks (`MEASURED_PEAKS.json`).  The MCMC's best allocations are
re-ranked with parameter-reallocation times computed by the real planner (`refine_with_planned_realloc`).
"""

from __future__ import annotations

import dataclasses
import json
import os
from typing import Dict, List, Optional, Tuple

from realhf_b200.api.config import ModelInterfaceType
from realhf_b200.api.dfg import MFCDef, build_graph
from realhf_b200.api.quickstart import DeviceMesh, ModelTrainEvalConfig, ParallelismConfig, RPCAllocation, find_parallel_strategies
from realhf_b200.ops import host

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


@dataclasses.dataclass
class HardwareModel:
    bf16_flops: float = 1.43e15      # sustained dense bf16 (MEASURED_PEAKS.json: bf16_tflops_sustained)
    hbm_bw: float = 6.58e12          # bytes/s
    link_bw: float = 7.7e11          # NVLink 5 peer copy, per direction per GPU
    mem_cap: float = 180e9
    gemm_eff: float = 0.75           # achieved fraction of peak for training-sized GEMMs
    launch_us: float = 4.0

    @classmethod
    def from_measured(cls) -> "HardwareModel":
        hw = cls()
        from realhf_b200.base import cluster
        hw.mem_cap = cluster.spec().gpu_memory_gb * 1e9
        p = os.path.join(_ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(p):
            try:
                d = json.load(open(p))
                hw.bf16_flops = d.get("bf16_tflops_sustained", hw.bf16_flops / 1e12) * 1e12
                hw.hbm_bw = d.get("hbm_gbs", hw.hbm_bw / 1e9) * 1e9
            except (OSError, ValueError, TypeError) as e:  # unreadable / malformed file: keep the documented defaults
                import warnings
                warnings.warn(f"ignoring {p}: {e}")
        return hw


def model_shape(mcfg: ModelTrainEvalConfig) -> Dict[str, float]:
    """Approximate LLaMA-family shapes by nominal size in billions (path-less configs), or read config.json."""
    if mcfg.path and os.path.exists(os.path.join(mcfg.path, "config.json")):
        c = json.load(open(os.path.join(mcfg.path, "config.json")))
        h = c.get("hidden_size", c.get("n_embd"))
        L = c.get("num_hidden_layers", c.get("n_layer"))
        f = c.get("intermediate_size", 4 * h)
        v = c.get("vocab_size", 32000)
        n = L * (4 * h * h + 3 * h * f) + 2 * v * h
        return dict(h=h, L=L, f=f, v=v, n=n)
    table = {7: (4096, 32, 11008), 13: (5120, 40, 13824), 34: (8192, 48, 22016), 70: (8192, 80, 28672)}
    size = mcfg.type.size or 7
    h, L, f = table.get(size, (4096, max(1, int(32 * size / 7)), 11008))
    v = 32000
    return dict(h=h, L=L, f=f, v=v, n=L * (4 * h * h + 3 * h * f) + 2 * v * h)


def estimate(rpc: MFCDef, shape: Dict[str, float], par: ParallelismConfig, hw: HardwareModel, seq_len: int, gen_len: int,
             n_minibatches: int, trainable_role: bool) -> Tuple[float, float, float]:
    """(time_us, static bytes / GPU, active bytes / GPU) of one MFC under one layout."""
    dp, tp, pp = par.data_parallel_size, par.model_parallel_size, par.pipeline_parallel_size
    n, h, L, v = shape["n"], shape["h"], shape["L"], shape["v"]
    T = rpc.n_seqs * (seq_len + gen_len)
    per_gpu_tokens = T / dp
    wbytes = 2 * n / (tp * pp)
    flops_fwd = 2 * n * per_gpu_tokens / (tp * pp)
    tp_comm = 0.0
    if tp > 1:  # 2 collectives per layer fwd (4 with bwd), [tokens, h] bf16 each
        vol = 2 * per_gpu_tokens * h * (tp - 1) / tp
        tp_comm = (L / pp) * 2 * vol / hw.link_bw
    bubble = (pp - 1) / max(1, 2 * pp) if pp > 1 else 0.0
    if rpc.interface_type == ModelInterfaceType.TRAIN_STEP:
        t = 4 * flops_fwd / (hw.bf16_flops * hw.gemm_eff) + 3 * tp_comm
        t *= 1 + bubble
        opt = n_minibatches * (14 * n / (tp * pp * dp)) / hw.hbm_bw  # fused AdamW over the shard
        zero_comm = n_minibatches * 2 * (2 * n / (tp * pp)) * (dp - 1) / dp / hw.link_bw if dp > 1 else 0.0
        t += opt + zero_comm
        static = wbytes + 12 * n / (tp * pp * dp)
        active = wbytes + 34 * per_gpu_tokens / n_minibatches * h * (L / pp) / tp / 8
    elif rpc.interface_type == ModelInterfaceType.GENERATE:
        bs = rpc.n_seqs / dp
        prefill = 2 * n * bs * seq_len / (tp * pp) / (hw.bf16_flops * hw.gemm_eff)
        kv_per_tok = 2 * 2 * h * L / (tp * pp)
        # one token round: every stage runs its pp micro-batches, re-reading its weight shard for each of them
        m = pp
        # calibrated on B200 decode profiles (profiles/decode_step_breakdown_*, bench logs): weight streaming reaches ~85% and
        # the KV-cache reads ~87% of the HBM copy peak; every layer adds ~45 us of per-kernel fixed cost (4 small-M GEMMs,
        # attention, norms / activation) and, under TP, two all-reduces that cost ~40 us each end to end inside the graph
        step = wbytes * m / (0.85 * hw.hbm_bw) + bs * (seq_len + gen_len / 2) * kv_per_tok / (0.87 * hw.hbm_bw) + m * (L / pp) * 45e-6
        if tp > 1:
            step += m * (L / pp) * 2 * 40e-6
        if pp > 1:
            step += pp * 10e-6                 # p2p hops of the token ring
        t = prefill + gen_len * step
        static = 0.0 if trainable_role else wbytes
        active = bs * (seq_len + gen_len) * kv_per_tok + (wbytes if trainable_role else 0.0)
    else:
        t = flops_fwd / (hw.bf16_flops * hw.gemm_eff) + tp_comm
        t *= 1 + bubble
        static = 0.0 if trainable_role else wbytes
        active = 6 * per_gpu_tokens * h / tp + (wbytes if trainable_role else 0.0)
    return t * 1e6, static, active


def profile_table_for(mcfg: ModelTrainEvalConfig):
    """The layer-profile table of a model family / size (`search/cost_model.py::ProfileTable.find`), or None."""
    from realhf_b200.search.cost_model import ProfileTable
    name = f"{mcfg.type._class}-{mcfg.type.size or 7}"
    return ProfileTable.find(name)


class MFCProfile:
    """Measured whole-MFC times from `quickstart profile` (`experiments/profile.py`: rows {handle, interface, layout "d2m2p1", bs,
    seqlen, n_mbs, secs}).  An exact hit -- same handle, layout, global batch and sequence length -- replaces the cost model's
    time for that candidate (the reference's search consumes its profile statistics the same way, by exact key:
    search_engine/estimate.py:263-360); everything else stays estimated."""

    _HANDLE = {ModelInterfaceType.GENERATE: "generate", ModelInterfaceType.INFERENCE: "inference", ModelInterfaceType.TRAIN_STEP: "train_step"}

    def __init__(self, rows: List[dict]):
        self.rows = [r for r in rows if "secs" in r and "layout" in r]
        self._idx = {}
        for r in self.rows:
            key = (r["handle"], r["layout"], int(r["bs"]), int(r["seqlen"]))
            best = self._idx.get(key)
            if best is None or r["secs"] < best["secs"]:      # several n_mbs for one point: the launcher can pick the fastest
                self._idx[key] = r

    @classmethod
    def find(cls) -> Optional["MFCProfile"]:
        p = os.environ.get("REAL_MFC_PROFILE", "")
        if p and os.path.exists(p):
            with open(p) as f:
                return cls(json.load(f))
        return None

    def time_us(self, rpc: MFCDef, par: ParallelismConfig, seq_len: int, gen_len: int) -> Optional[float]:
        handle = self._HANDLE[rpc.interface_type]
        sl = seq_len if rpc.interface_type == ModelInterfaceType.GENERATE else seq_len + gen_len
        lay = f"d{par.data_parallel_size}m{par.model_parallel_size}p{par.pipeline_parallel_size}"
        r = self._idx.get((handle, lay, int(rpc.n_seqs), int(sl)))
        if r is None or (r.get("interface") and rpc.interface_impl is not None and r["interface"] != rpc.interface_impl.type_):
            return None
        return float(r["secs"]) * 1e6


def build_problem(mesh: DeviceMesh, rpcs: List[MFCDef], models: Dict[str, ModelTrainEvalConfig], seq_len: int, gen_len: int,
                  n_ppo_minibatches: int, hw: HardwareModel, max_cands: int = 1000, use_profile_tables: bool = True,
                  mfc_profile: Optional[MFCProfile] = None):
    """Candidates per MFC = every (sub-mesh, dp x tp x pp) that fits the batch, costed by the table-driven model
    (`search/cost_model.py::estimate_mfc`) when the role's model has a layer-profile table, else by the roofline formulas above."""
    from realhf_b200.search.cost_model import CommModel, estimate_mfc
    G = build_graph(rpcs)
    roles = sorted({r.role for r in rpcs})
    role_idx = {r: i for i, r in enumerate(roles)}
    shapes = {r: model_shape(models[r]) for r in roles}
    tables = {r: (profile_table_for(models[r]) if use_profile_tables else None) for r in roles}
    comm = CommModel.from_measured(min(8, mesh.n_gpus_per_node))
    trainable = {r.role for r in rpcs if r.interface_type == ModelInterfaceType.TRAIN_STEP}
    sub = [mesh] + [m for m in mesh.sub_device_meshes() if m != mesh]
    mesh_ranks = [m.global_ranks() for m in sub]
    kind = {ModelInterfaceType.GENERATE: 0, ModelInterfaceType.INFERENCE: 1, ModelInterfaceType.TRAIN_STEP: 2}
    prpcs, table = [], []
    for r in rpcs:
        cands = []
        for mi, m in enumerate(sub):
            for par in find_parallel_strategies(m):
                if r.n_seqs < par.data_parallel_size * par.pipeline_parallel_size:
                    continue
                if getattr(r, "balanced_dp", False) and r.n_seqs % par.data_parallel_size:
                    continue   # equal shares per DP rank are impossible under this layout
                if shapes[r.role]["v"] % par.model_parallel_size or shapes[r.role]["L"] < par.pipeline_parallel_size:
                    continue
                n_mini = n_ppo_minibatches if r.interface_type == ModelInterfaceType.TRAIN_STEP else 1
                if tables[r.role] is not None:
                    c = estimate_mfc(r.interface_type, r.n_seqs, shapes[r.role], par.data_parallel_size, par.model_parallel_size,
                                     par.pipeline_parallel_size, hw, tables[r.role], comm, seq_len, gen_len, n_minibatches=n_mini,
                                     n_mbs=max(1, getattr(r, "n_mbs", 1) or 1), trainable_role=r.role in trainable,
                                     use_sequence_parallel=par.model_parallel_size > 1, gpus_per_node=mesh.n_gpus_per_node)
                    t, st, ac = c.time_us, c.mem_static, c.mem_active
                else:
                    t, st, ac = estimate(r, shapes[r.role], par, hw, seq_len, gen_len, n_mini, r.role in trainable)
                if mfc_profile is not None:
                    measured = mfc_profile.time_us(r, par, seq_len, gen_len)
                    if measured is not None:
                        t = measured
                cands.append((mi, par.data_parallel_size, par.model_parallel_size, par.pipeline_parallel_size, t, st, ac, par))
        cands.sort(key=lambda c: c[4])
        cands = cands[:max_cands]
        table.append(cands)
        prpcs.append(dict(name=r.name, role=role_idx[r.role], kind=kind[r.interface_type], cands=[c[:7] for c in cands]))
    name_idx = {r.name: i for i, r in enumerate(rpcs)}
    edges = [(name_idx[u], name_idx[v]) for u, v in G.edges()]
    prob = dict(n_gpus=mesh.n_nodes * mesh.n_gpus_per_node, mem_cap=hw.mem_cap, link_bw=hw.link_bw, n_iters=2,
                role_bytes=[2 * shapes[r]["n"] for r in roles], meshes=mesh_ranks, edges=edges, rpcs=prpcs)
    prob["_shapes"] = [shapes[r] for r in roles]      # python-side only (the native parser ignores unknown keys)
    return prob, table, sub


def refine_with_planned_realloc(prob: dict, results: List[dict], hw: HardwareModel, gpus_per_node: int = 8) -> List[dict]:
    """Re-rank the MCMC's best allocations with the parameter-reallocation times of the REAL planner.

    The search itself uses the simulator's closed form (destination shard bytes / link bandwidth).  For each of its top results the
    (train layout -> other layout) pairs are planned with `parallel/realloc.py::derive_plan`, costed per GPU
    (`cost_model.realloc_time_us`), handed to the simulator as an override table, and the results are re-simulated and re-sorted."""
    from realhf_b200.search.cost_model import CommModel, config_from_shape, realloc_time_us
    h = host()
    comm = CommModel.from_measured(min(8, gpus_per_node))
    rpcs = prob["rpcs"]
    train_of = {r["role"]: i for i, r in enumerate(rpcs) if r["kind"] == 2}
    table: Dict[Tuple[int, ...], float] = {}
    for res in results:
        for i, r in enumerate(rpcs):
            t = train_of.get(r["role"])
            if t is None or t == i:
                continue
            src, dst = rpcs[t]["cands"][res["choice"][t]], r["cands"][res["choice"][i]]
            key = (r["role"],) + tuple(int(x) for x in src[:4]) + tuple(int(x) for x in dst[:4])
            if key in table or tuple(src[:4]) == tuple(dst[:4]):
                continue
            cfg = config_from_shape(prob["_shapes"][r["role"]])
            table[key] = realloc_time_us(cfg, (src[1], src[2], src[3]), prob["meshes"][src[0]], (dst[1], dst[2], dst[3]),
                                         prob["meshes"][dst[0]], hw, comm, gpus_per_node)
    p2 = dict(prob, realloc_table=[(list(k), v) for k, v in table.items()])
    out = []
    for res in results:
        sim = h.simulate_allocation(p2, list(res["choice"]))
        out.append(dict(res, cost=sim["cost"], time_us=sim["time_us"], max_mem=sim["max_mem"], time_us_closed_form=res["time_us"]))
    out.sort(key=lambda d: d["cost"])
    return out, p2


def search_rpc_allocations(device_mesh: DeviceMesh, rpcs: List[MFCDef], models: Dict[str, ModelTrainEvalConfig], seq_len: int = 128,
                           num_gen_tokens: int = 256, n_ppo_minibatches: int = 4, time_limit_s: float = 5.0,
                           hw: Optional[HardwareModel] = None, return_details: bool = False, refine_realloc: bool = True,
                           mfc_profile: Optional["MFCProfile"] = None, cross_step_overlap: bool = True):
    h = host()
    if h is None:
        raise RuntimeError("allocation search needs the native host extension: run `python -m realhf_b200.ops.build`")
    if hw is None:
        hw = HardwareModel.from_measured()
        hw.mem_cap = min(hw.mem_cap, float(device_mesh.gpu_memory_capacity))   # a mesh may declare less than the cluster spec
    prob, table, sub = build_problem(device_mesh, rpcs, models, seq_len, num_gen_tokens, n_ppo_minibatches, hw,
                                     mfc_profile=mfc_profile if mfc_profile is not None else MFCProfile.find())
    # the master walks the graph with a look-ahead of one step (`exp_ctrl.max_inflight_steps=2`): the simulator then scores the
    # steady state of two overlapping iterations; with a barrier after every step (`=1`) one traversal is the whole story
    prob["n_iters"] = 2 if cross_step_overlap else 1
    results = h.multi_mcmc_search(prob, [0.5, 2.0, 8.0, 32.0], time_limit_s, 1, 10)
    if refine_realloc:
        results, prob = refine_with_planned_realloc(prob, results, hw, device_mesh.n_gpus_per_node)
    best = results[0]
    allocs = []
    for r, ci, cands in zip(rpcs, best["choice"], table):
        c = cands[ci]
        par: ParallelismConfig = dataclasses.replace(c[7])
        par.use_sequence_parallel = (r.interface_type == ModelInterfaceType.TRAIN_STEP and par.model_parallel_size > 1)
        allocs.append(RPCAllocation(r, sub[c[0]], par))
    if return_details:
        return allocs, dict(best=best, results=results, problem=prob)
    return allocs
