#!/usr/bin/env python
"""Headline benchmark: PPO iteration throughput (tokens/s, generation + inference + training) for
LLaMA-7B actor + 7B critic + 7B reference + 7B reward model, synthetic prompts, random-init weights, bf16.

    python bench.py --gpus 1 --steps 2 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        bench.py --gpus 8 --steps 2 --warmup 3

Config = the reference's quickstart PPO script (`examples/scripts/local/ppo.sh`): 128 prompts x 128 tokens,
512 new tokens (min = max), top_p 0.9, top_k 1000, 4 PPO minibatches, CUDA-graph generation.  One step = one full
traversal of the 6-MFC PPO dataflow graph (actor_gen -> rew_inf, ref_inf, critic_inf -> actor_train, critic_train).
Total work is fixed as N grows (strong scaling): tokens per step = 128 x 640.

Allocation on B200: every MFC is data-parallel over all N GPUs (four 7B models are 54 GB in bf16; 180 GB HBM
holds them all), trainable models use the ZeRO-1 sharded flat AdamW.  Timing: CUDA events on the launching stream
bracketed by barrier + synchronize, MAX over ranks.  Inputs (> L2: weights alone are 54 GB) are copied from pinned
host memory every step and the step's statistics are read back to the host inside the e2e region.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "PPO tokens/sec (gen+train, device-timed max-over-ranks) LLaMA-7B x4"
BASELINE_TOKENS_PER_S = 128 * 640 / (574.312 / 39)  # reference quickstart log: 39 steps in 574.312 s on 8 GPUs


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons of the node's GPUs during the timed region.

    One sampler per node (local rank 0) through NVML in-process (`nvidia_ml_py`), ONE GPU per tick (round robin), one tick every
    3 s: driver queries contend with CUDA-graph launches -- sampling one GPU every 0.5 s made its generation MFC 10-25% slower
    inside the timed region than in the (unsampled) warm-up steps, and the queries of all GPUs of a node go through the same
    driver lock, so querying 8 GPUs per tick costs every rank what an 8x shorter period costs a single GPU.  The first tick is
    immediate; a region of T seconds yields 1 + T/3 samples from as many different GPUs (`gpus_sampled`).  Falls back to one
    `nvidia-smi` query per tick when NVML is unavailable."""

    _REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4))

    def __init__(self, n_gpus: int, period: float = 3.0):
        super().__init__(daemon=True)
        self.n_gpus, self.period = n_gpus, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.sampled_gpus = set()
        self._tick = 0
        self._stop_ev = threading.Event()

    def _sample_nvml(self, nv, handles):
        i = self._tick % len(handles)
        self._tick += 1
        self.sampled_gpus.add(i)
        for h in handles[i: i + 1]:
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            if self.max_mhz is None:
                self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            try:
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for name, bit in self._REASONS:
                if mask & bit:
                    self.reasons.add(name)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True,
                             timeout=10).stdout.strip().splitlines()
        self.sampled_gpus.update(range(min(self.n_gpus, len(out))))
        for line in out[: self.n_gpus]:
            f = line.split(",")
            self.samples.append(float(f[0]))
            self.max_mhz = float(f[1])
            for n, v in zip(names, f[2:]):
                if v.strip().lower().startswith("active"):
                    self.reasons.add(n)

    def run(self):
        nv = handles = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = [int(x) for x in vis.split(",")][: self.n_gpus] if vis and all(x.strip().isdigit() for x in vis.split(",")) else list(range(self.n_gpus))
            handles = [nv.nvmlDeviceGetHandleByIndex(i) for i in idx]
        except Exception:
            nv = None
        while not self._stop_ev.is_set():
            try:
                if nv is not None:
                    self._sample_nvml(nv, handles)
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop_ev.wait(self.period)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=10)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "n_samples": len(s), "gpus_sampled": len(self.sampled_gpus), "gpus": self.n_gpus}


class _NoSampler:
    def start(self):
        pass

    def stop(self):
        return None


def run_reference(args):
    """The unmodified reference cannot run in this image: see DESIGN.md ("Reference arm")."""
    why = ("reference installs into baseline/_ref only with --no-deps --ignore-requires-python; unmodified it cannot be imported on "
           "this image's Python 3.12: realhf/api/core/config.py declares a dataclass instance as a field default, which dataclasses "
           "reject since 3.11 (ValueError: mutable default ... use default_factory) -- every entry point imports that module; beyond "
           "it the PPO path needs megatron-core 0.6, deepspeed 0.14 and hydra-core, none of which exist offline "
           "(python baseline/probe_reference.py prints the evidence)")
    if os.environ.get("RANK", "0") == "0":  # launched like the other arm (torchrun for N > 1): one line, from rank 0
        print(json.dumps({"impl": "reference", "unavailable": why}))
    return 0


def run_master_runtime(args):
    """`--runtime master`: the SAME PPO config driven through the production runtime -- quickstart config -> launcher
    (`apps/main.py::main_start`) -> one master worker (asyncio DFG walker, metadata only) + one model worker process per GPU over
    ZMQ / NCCL, dataset on disk, tokenizer, data transfer between MFCs -- instead of the in-process SPMD executor.  The step time
    is the master's host clock around one full DFG traversal (what the reference reports as e2e time,
    system/master_worker.py:1353-1358); warm-up steps are dropped.  Random-init weights (`init_from_scratch`), synthetic prompts
    padded to `prompt_len`, `min_new_tokens == max_new_tokens` as in the reference's benchmark recipe (docs quickstart.rst:287-297)."""
    import re
    import tempfile
    import uuid

    import torch
    n = args.gpus
    root = tempfile.mkdtemp(prefix="realhf_b200_bench_")
    # BEFORE the package is imported: `base.constants` resolves the log / checkpoint roots from the environment at import time,
    # and the worker processes (which inherit the environment) must resolve the same paths as this launcher
    os.environ.setdefault("REAL_FILEROOT", os.path.join(root, "fileroot"))
    os.environ["PYTHONPATH"] = ROOT + os.pathsep + os.environ.get("PYTHONPATH", "")  # the worker processes import this checkout
    os.environ["REAL_FAST_INIT"] = "1"  # device-side random init (a host-side draw of 4 x 6.7e9 normals takes minutes)
    assert "realhf_b200.base.constants" not in sys.modules, "run_master_runtime must set REAL_FILEROOT before realhf_b200 is imported"
    from realhf_b200.api.model import ReaLModelConfig
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    from realhf_b200.models import hf_io
    # tokenizer + config-only "checkpoint" directories
    from tokenizers import Tokenizer, models, pre_tokenizers, trainers
    from transformers import PreTrainedTokenizerFast
    import random
    rng = random.Random(0)
    words = ["".join(rng.choice("abcdefghijklmnop") for _ in range(rng.randint(2, 6))) for _ in range(400)]
    tk = Tokenizer(models.WordLevel(unk_token="[UNK]"))
    tk.pre_tokenizer = pre_tokenizers.Whitespace()
    tk.train_from_iterator([" ".join(words)], trainers.WordLevelTrainer(vocab_size=1000, special_tokens=["[PAD]", "[EOS]", "[UNK]"]))
    fast = PreTrainedTokenizerFast(tokenizer_object=tk, pad_token="[PAD]", eos_token="[EOS]", unk_token="[UNK]")
    dirs = {}
    for role, critic in (("actor", False), ("critic", True)):
        d = os.path.join(root, role)
        os.makedirs(d)
        cfg = ReaLModelConfig(n_layers=args.layers, n_kv_heads=32, n_q_heads=32, hidden_dim=4096, intermediate_dim=11008, vocab_size=32000,
                              n_positions=4096, embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, layer_norm_epsilon=1e-5,
                              activation_function="silu", scale_attn_by_inverse_layer_idx=False, use_attention_bias=False,
                              use_attn_proj_bias=False, use_mlp_bias=False, layer_norm_type="rms", mlp_type="llama", apply_rotary=True,
                              is_critic=critic)
        if args.tiny:
            cfg.hidden_dim, cfg.intermediate_dim, cfg.n_q_heads, cfg.n_kv_heads, cfg.head_dim, cfg.vocab_size = 256, 512, 4, 4, 64, 1024
        fam = hf_io.family("llama")
        hf_cfg = fam.config_to_hf(cfg)
        hf_cfg.architectures = [fam.hf_cls_name]
        hf_cfg.save_pretrained(d)
        fast.save_pretrained(d)
        dirs[role] = d
    data = os.path.join(root, "prompts.jsonl")
    total = args.warmup + args.steps
    with open(data, "w") as f:
        for i in range(args.prompts * total):
            f.write(json.dumps(dict(id=i, prompt=" ".join(rng.choice(words) for _ in range(2 * args.prompt_len)))) + "\n")
    dev = "cpu" if args.tiny and not torch.cuda.is_available() else "cuda"
    exp_name = f"bench-{uuid.uuid4().hex[:6]}"
    qs = ["ppo", f"experiment_name={exp_name}", "trial_name=t0", f"device={dev}", f"dtype={'fp32' if dev == 'cpu' else 'bf16'}",
          f"n_gpus_per_node={n}", f"allocation_mode={'manual' if args.allocation == 'dp' else args.allocation}", f"dataset.path={data}", f"dataset.train_bs_n_seqs={args.prompts}",
          f"dataset.max_prompt_len={args.prompt_len}", "dataset.pad_to_max_length=true",
          f"ppo.gen.max_new_tokens={args.new_tokens}", f"ppo.gen.min_new_tokens={args.new_tokens}", "ppo.gen.top_p=0.9", "ppo.gen.top_k=1000",
          "ppo.gen.use_cuda_graph=true", "ppo.gen.force_cudagraph_recapture=true", "ppo.ppo_n_minibatches=4",
          "exp_ctrl.total_train_epochs=1", f"exp_ctrl.benchmark_steps={total}"]
    inf_mbs = 2 if (args.prompts // n) * (args.prompt_len + args.new_tokens) > 48 * 1024 else 1
    for role, path in (("actor", dirs["actor"]), ("ref", dirs["actor"]), ("critic", dirs["critic"]), ("rew", dirs["critic"])):
        qs += [f"{role}.type._class=llama", f"{role}.path={path}", f"{role}.init_from_scratch=true",
               f"{role}.gradient_checkpointing={'auto' if dev == 'cuda' else 'false'}"]
    for role in ("actor", "critic"):
        qs += [f"{role}.optimizer.state_dtype=bf16", f"{role}.optimizer.use_master_weights=false", f"{role}.optimizer.grad_dtype=bf16",
               f"{role}.optimizer.share_grad_buffer=true", f"{role}.optimizer.lr_scheduler_type=constant", f"{role}.optimizer.warmup_steps_proportion=0.0"]
        if dev == "cpu":
            qs += [f"{role}.optimizer.state_dtype=fp32", f"{role}.optimizer.grad_dtype=fp32"]
    if args.allocation == "dp":
        for mfc in ("actor_gen", "actor_train", "critic_train", "critic_inf", "ref_inf", "rew_inf"):
            qs += [f"{mfc}.parallel.data_parallel_size={n}"]
    for mfc in ("critic_inf", "ref_inf", "rew_inf"):
        qs += [f"{mfc}.n_mbs={inf_mbs}"]
    exp = build_experiment(qs)
    layouts = None
    if args.allocation != "dp":   # resolve once here to report what the mode chose (the launcher resolves the same thing again)
        try:
            layouts = {a.rpc.name if hasattr(a.rpc, "name") else str(a.rpc):
                       f"gpus{a.device_mesh.global_ranks()} d{a.parallel.data_parallel_size}"
                       f"m{a.parallel.model_parallel_size}p{a.parallel.pipeline_parallel_size}" for a in exp._get_rpc_allocations()}
        except Exception as e:  # reporting only
            layouts = {"unresolved": repr(e)}
    t0 = time.perf_counter()
    main_start(exp, timeout=3600)
    wall_total = time.perf_counter() - t0
    log = open(os.path.join(os.environ["REAL_FILEROOT"], "logs", exp_name, "t0", "master_worker-0")).read()
    steps = [(float(m.group(1)), m.group(2)) for m in re.finditer(r"step \d+ \(epoch[^)]*\) e2e ([0-9.]+)s; (.*)", log)]
    assert len(steps) >= total, f"master log has {len(steps)} steps, expected {total}:\n{log[-3000:]}"
    timed = steps[args.warmup: total]
    secs = sum(t for t, _ in timed)
    mfc = {}
    for _, line in timed:
        for part in line.split(", "):
            k, v = part.rsplit(" ", 1)
            mfc[k] = mfc.get(k, 0.0) + float(v.rstrip("s")) * 1e3 / len(timed)
    tokens_per_step = args.prompts * (args.prompt_len + args.new_tokens)
    value = tokens_per_step * len(timed) / secs
    headline = args.layers == 32 and args.prompts == 128 and args.prompt_len == 128 and args.new_tokens == 512 and not args.tiny
    print(json.dumps({
        "metric": METRIC, "value": round(value, 1), "unit": "tokens/s", "n_gpus": n, "steps": len(timed), "warmup": args.warmup,
        "ms_per_step": round(secs * 1e3 / len(timed), 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": round(value / BASELINE_TOKENS_PER_S, 3) if headline else None, "dtype": "bf16" if dev == "cuda" else "fp32",
        "data": "synthetic prompts (random words, padded to prompt_len), random-init weights",
        "runtime": "master/worker (quickstart -> launcher -> master worker + one model worker process per GPU, ZMQ control plane)",
        "timer": "master worker host clock around each DFG traversal (includes dataset fetch, data transfer, all six MFCs, replies)",
        "config": {"model": "LLaMA-7B actor + 7B critic + 7B ref + 7B reward" + ("" if headline else " [DEBUG shapes]"),
                   "global_batch": args.prompts, "seq_len": args.prompt_len + args.new_tokens,
                   "parallelism": f"dp{n} (all 6 MFCs), ZeRO-1 flat AdamW" if layouts is None else f"allocation_mode={args.allocation}: {layouts}",
                   "mfc_ms": {k: round(v, 1) for k, v in mfc.items()}, "launch_to_exit_s": round(wall_total, 1)},
        "impl": "ours"}), flush=True)
    return 0


def choose_gen_tp(world: int, n_prompts: int, prompt_len: int, new_tokens: int, n_layers: int) -> dict:
    """Generation layout for this node size, picked by the allocation search's cost model (`search/engine.py::estimate`, the
    function the `search` / `heuristic` allocation modes use) over every tp x dp factorisation of the node: decode streams
    1/tp of the weights per GPU per token, pays the layer-boundary all-reduces (fused into the RMSNorm kernels over NVSwitch
    multicast) and keeps the same KV bytes per GPU.  Returns {tp: predicted seconds} and the arg-min."""
    from realhf_b200.api.config import ModelInterfaceAbstraction, ModelInterfaceType
    from realhf_b200.api.dfg import MFCDef
    from realhf_b200.api.quickstart import ParallelismConfig
    from realhf_b200.search.engine import HardwareModel, estimate
    hw = HardwareModel.from_measured()
    h, f, v = 4096, 11008, 32000
    shape = dict(h=h, L=n_layers, f=f, v=v, n=n_layers * (4 * h * h + 3 * h * f) + 2 * v * h)
    rpc = MFCDef("actor_gen", n_prompts, ModelInterfaceType.GENERATE, ModelInterfaceAbstraction("ppo_actor"), "actor",
                 input_keys=("packed_prompts",), output_keys=("packed_input_ids",))
    pred = {}
    for tp in (1, 2, 4, 8):
        if world % tp or 32 % tp:
            continue
        par = ParallelismConfig(data_parallel_size=world // tp, model_parallel_size=tp, pipeline_parallel_size=1)
        pred[tp] = estimate(rpc, shape, par, hw, prompt_len, new_tokens, 1, True)[0] / 1e6
    best = min(pred, key=pred.get)
    return dict(pred_s={k: round(t, 3) for k, t in pred.items()}, best=best)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torchlib"],
                    help="ours; reference (the unmodified reference: unavailable offline, see DESIGN.md); torchlib: an on-box A/B arm that is "
                         "NOT the reference -- this repo's executor with library kernels (cuBLAS GEMMs, flash-attn varlen attention, "
                         "F.layer_norm, NCCL reduce-scatter / all-gather / all-reduce) in place of the sm_100a kernels")
    ap.add_argument("--layers", type=int, default=32, help="debug only: fewer layers (result is then NOT the headline config)")
    ap.add_argument("--prompts", type=int, default=128)
    ap.add_argument("--prompt-len", type=int, default=128)
    ap.add_argument("--new-tokens", type=int, default=512)
    ap.add_argument("--gemm", default=os.environ.get("REAL_GEMM", "tcgen05"), choices=["tcgen05", "cublas"])
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--runtime", default="spmd", choices=["spmd", "master"],
                    help="spmd (default): every rank walks the DFG in-process (launched by torchrun for N > 1); master: the production "
                         "master/worker runtime launched through the quickstart + local scheduler (run WITHOUT torchrun: it spawns its own workers)")
    ap.add_argument("--allocation", default="dp", choices=["dp", "search", "heuristic"],
                    help="--runtime master only: `dp` = every MFC data-parallel over all GPUs (the SPMD arm's allocation); `search` / `heuristic` "
                         "= the quickstart allocation modes (per-MFC device meshes and layouts, parameter reallocation, MFCs of different "
                         "steps overlapping on disjoint GPUs)")
    ap.add_argument("--profile-mfc", default="", help="debug only: comma-separated MFC names to wrap in torch.profiler during the LAST timed step; "
                                                     "prints kernel-time totals and the top kernels per MFC to stderr (the run's numbers are then not a bench value)")
    ap.add_argument("--tiny", action="store_true", help="debug only: toy model shapes for --runtime master (CPU smoke test of the arm)")
    ap.add_argument("--optimizer", default="lean", choices=["lean", "fp32"],
                    help="lean (default at every N, so the scaling curve compares like with like): bf16 Adam moments + stochastic "
                         "rounding, no master copy -- what fits four 7B models + optimizer states on ONE GPU; fp32: fp32 master weights + "
                         "fp32 moments, the reference's Megatron optimizer precision (fits from N >= 2; reported in profiles/)")
    ap.add_argument("--ckpt", default="auto", type=lambda v: {"auto": "auto", "all": True, "none": False}[v],
                    help="activation checkpointing: auto (recompute only what the free HBM requires), all (every block, the reference default), none")
    ap.add_argument("--offload-frozen", default="auto", choices=["auto", "on", "off"],
                    help="drop the idle reference / reward weights during training and stream them back during generation "
                         "(auto: on for a single GPU with --ckpt auto, where the freed HBM buys un-recomputed blocks)")
    ap.add_argument("--gen-tp", type=int, default=int(os.environ.get("REAL_BENCH_GEN_TP", "0")),
                    help="tensor-parallel degree of the generation replica (0: default for this N; 1: generate on the dp layout)")
    ap.add_argument("--gen-fp8", action="store_true",
                    help="NOT the headline precision: decode with e4m3 weights / activations (GenerationHyperparameters.fp8_weights); "
                         "the JSON line is labelled and carries no vs_baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "torchlib":
        args.gemm = "cublas"
        os.environ.update(REAL_ATTN="flash", REAL_ATTN_BWD="flash", REAL_LAYERNORM="torch", REAL_ZERO_COMM="nccl", REAL_TP_NVLS="0",
                          REAL_ZERO_OVERLAP="0")
    if args.runtime == "master":
        return run_master_runtime(args)

    # four 7B models + optimizer state + a 43 GB KV cache leave little slack on one GPU: expandable segments keep the caching
    # allocator from fragmenting (a failed 43 GB request makes it free and re-cudaMalloc its whole cache inside the timed region)
    os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
    import torch
    import torch.distributed as dist

    from realhf_b200.api.config import ModelInterfaceAbstraction, ModelInterfaceType, ModelName
    from realhf_b200.api.data import SequenceSample
    from realhf_b200.api.dfg import MFCDef
    from realhf_b200.api.model import FinetuneSpec, Model, ReaLModelConfig
    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.engine.engine import InferenceBackend, TrainBackend
    from realhf_b200.interfaces import basic, ppo
    from realhf_b200.models.real_model import ReaLModel
    from realhf_b200.ops import functional as OF
    from realhf_b200.ops import launches
    from realhf_b200.system.spmd import SPMDExecutor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        ctx = ParallelContext.build(ProcessTopology(1, world, 1), list(range(world)), rank, backend="nccl",
                                    gradient_checkpointing=args.ckpt)
    else:
        ctx = ParallelContext.single()
        ctx.gradient_checkpointing = args.ckpt
    if args.gemm == "tcgen05":
        from realhf_b200.ops import gemm as G
        OF.set_gemm_impl(G.linear)
    else:
        OF.set_gemm_impl(False)  # library matmul, for A/B runs only

    def llama7b(is_critic):
        return ReaLModelConfig(n_layers=args.layers, n_kv_heads=32, n_q_heads=32, hidden_dim=4096, intermediate_dim=11008,
                               vocab_size=32000, n_positions=4096, embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0,
                               layer_norm_epsilon=1e-5, activation_function="silu", scale_attn_by_inverse_layer_idx=False,
                               use_attention_bias=False, use_attn_proj_bias=False, use_mlp_bias=False, layer_norm_type="rms",
                               mlp_type="llama", apply_rotary=True, is_critic=is_critic)

    tok = types.SimpleNamespace(eos_token_id=2, pad_token_id=0)
    spec = FinetuneSpec(total_train_epochs=1, total_train_steps=1000, steps_per_epoch=1000)
    lean = dict(lr=1e-5, weight_decay=0.05, state_dtype="bf16", use_master_weights=False, grad_dtype="bf16",
                share_grad_buffer=True, warmup_steps_proportion=0.0, lr_scheduler_type="constant")
    if args.optimizer == "fp32":  # the reference's precision: fp32 master weights + fp32 Adam moments (Megatron distributed optimizer)
        lean.update(state_dtype="fp32", use_master_weights=True)
    models = {}
    for role, critic, train in (("actor", False, True), ("critic", True, True), ("ref", False, False), ("reward", True, False)):
        m = ReaLModel(llama7b(critic), ctx, dtype=torch.bfloat16, device=dev).init_random_fast(seed=11 + len(models))
        model = Model(ModelName(role, 0), m, tok, dev)
        be = TrainBackend(optimizer=dict(lean)) if train else InferenceBackend()
        models[role] = be.initialize(model, spec)
    torch.cuda.synchronize()

    per_rank = args.prompts // world
    assert per_rank * world == args.prompts
    gen_mbs = 1  # one decode pass over all local sequences: weights are streamed once per token (KV cache 43 GB at N=1)
    inf_mbs = 2 if per_rank * (args.prompt_len + args.new_tokens) > 48 * 1024 else 1
    # On one GPU the 43 GB KV cache must be freed before training, so the decode graph is re-captured every step (~0.15-0.3 s per
    # generate call: profiles/bench_n1_r2_gen_phases.log); at N=2 keeping the 21.5 GB cache resident costs 7 un-recomputed blocks of
    # the activation budget and loses more in training than it gains in generation (measured 9.9 k vs 10.1 k tokens/s).  From 4 GPUs
    # on the cache (<= 10.7 GB per GPU) stays resident together with the captured graph, like the reference's default
    # (`force_cudagraph_recapture=False`).
    recapture = world <= 2
    gcfg = dict(max_new_tokens=args.new_tokens, min_new_tokens=args.new_tokens, greedy=False, top_p=0.9, top_k=1000,
                temperature=1.0, use_cuda_graph=True, force_cudagraph_recapture=recapture, fp8_weights=args.gen_fp8)
    ppo_kw = dict(n_minibatches=4, kl_ctl=0.1, discount=1.0, gae_lambda=1.0, eps_clip=0.2, value_eps_clip=0.2,
                  max_reward_clip=20.0, adaptive_kl_ctl=False, value_norm=True)
    A = lambda t, **a: ModelInterfaceAbstraction(t, a)
    T = ModelInterfaceType
    n = args.prompts
    rpcs = [
        MFCDef("actor_gen", n, T.GENERATE, A("ppo_actor"), "actor", input_keys=("packed_prompts",),
               output_keys=("seq_no_eos_mask", "packed_input_ids", "packed_logprobs", "prompt_mask", "packed_logits_mask"), n_mbs=gen_mbs),
        MFCDef("rew_inf", n, T.INFERENCE, A("paired_rw"), "reward", input_keys=("packed_input_ids",), output_keys=("rewards",), n_mbs=inf_mbs),
        MFCDef("ref_inf", n, T.INFERENCE, A("ppo_actor"), "ref", input_keys=("packed_input_ids", "packed_logits_mask"),
               output_keys=("packed_ref_logprobs",), n_mbs=inf_mbs),
        MFCDef("critic_inf", n, T.INFERENCE, A("ppo_critic"), "critic", input_keys=("packed_input_ids", "seq_no_eos_mask"),
               output_keys=("values",), n_mbs=inf_mbs),
        MFCDef("actor_train", n, T.TRAIN_STEP, A("ppo_actor"), "actor",
               input_keys=("packed_input_ids", "packed_logprobs", "packed_ref_logprobs", "rewards", "values", "prompt_mask",
                           "seq_no_eos_mask", "packed_logits_mask"), n_mbs=1, log_return_value=True),
        MFCDef("critic_train", n, T.TRAIN_STEP, A("ppo_critic"), "critic",
               input_keys=("packed_input_ids", "packed_logprobs", "packed_ref_logprobs", "rewards", "values", "prompt_mask",
                           "seq_no_eos_mask"), n_mbs=1, log_return_value=True),
    ]
    actor_itf = ppo.PPOActorInterface(generation_config=gcfg, **ppo_kw)
    critic_itf = ppo.PPOCriticInterface(**{k: v for k, v in ppo_kw.items() if k not in ("eps_clip",)})
    rw_itf = basic.PairedRewardInterface()
    interfaces = {"actor_gen": actor_itf, "ref_inf": actor_itf, "actor_train": actor_itf, "critic_inf": critic_itf,
                  "critic_train": critic_itf, "rew_inf": rw_itf}
    ex = SPMDExecutor(rpcs, models, interfaces, dev)
    if args.offload_frozen == "on" or (args.offload_frozen == "auto" and world == 1 and args.ckpt == "auto"):
        # The reference and reward models idle from the end of their inference MFC until the next step's: their device copies
        # are dropped there (the pinned host copy of frozen weights stays valid) and streamed back on a side stream while the
        # actor generates.  The ~27 GB this frees during training go to the activation budget (fewer recomputed blocks); the
        # reference does the same with its OffloadHook (experiments/common/utils.py:182-198).
        side = torch.cuda.Stream(dev)
        frozen = {"ref": models["ref"].module.module, "reward": models["reward"].module.module}
        for mfc, role in (("ref_inf", "ref"), ("rew_inf", "reward")):
            ex.post_hooks.setdefault(mfc, []).append(lambda m=frozen[role]: m.offload(frozen=True))
            ex.hooks.setdefault(mfc, []).append(lambda: torch.cuda.current_stream(dev).wait_stream(side))
        ex.hooks.setdefault("actor_gen", []).append(lambda: [m.reload(stream=side) for m in frozen.values()])
    gen_choice = choose_gen_tp(world, args.prompts, args.prompt_len, args.new_tokens, args.layers)
    gen_tp = args.gen_tp if args.gen_tp > 0 else gen_choice["best"]
    if world > 1 and gen_tp > 1:
        # generation on a tp x dp replica of the actor: decode streams 1/tp of the weights per GPU per token; the replica is
        # refreshed from the (dp-replicated) training layout by local segment copies before every generation
        assert world % gen_tp == 0
        gen_topo = ProcessTopology(1, world // gen_tp, gen_tp)
        ctx_gen = ParallelContext.build(gen_topo, list(range(world)), rank, backend="nccl")
        ex.add_layout_replica("actor_gen", models["actor"], ctx_gen, ProcessTopology(1, world, 1), gen_topo, list(range(world)), rank)

    # synthetic prompts of the dataset's padded shape, in pinned host memory
    gcpu = torch.Generator().manual_seed(1234 + rank)
    host_prompts = torch.randint(3, 32000, (args.warmup + args.steps, per_rank * args.prompt_len), generator=gcpu).pin_memory()
    h2d_bytes = per_rank * args.prompt_len * host_prompts.element_size()
    result_host = torch.zeros(8, dtype=torch.float32).pin_memory()

    def one_step(i):
        ids = host_prompts[i].to(dev, non_blocking=True)
        batch = SequenceSample.from_default(seqlens=[args.prompt_len] * per_rank,
                                            ids=[f"s{i}r{rank}p{j}" for j in range(per_rank)], data=dict(packed_prompts=ids))
        rec = ex.run_step(batch)
        st_a, st_c = rec["actor_train"].result, rec["critic_train"].result
        res = torch.tensor([st_a["actor_loss"], st_a["task_reward"], st_a["importance_weight"], st_c["value_loss"],
                            float(st_a["n_tokens"]), float(st_a["grad_norm"]), float(st_c["grad_norm"]), 0.0])
        result_host.copy_(res)
        return rec, st_a, st_c

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    prof_state = {}
    if args.profile_mfc and rank == 0:
        from torch.profiler import ProfilerActivity, profile

        def _start(name):
            if prof_state.get("armed"):
                torch.cuda.synchronize()
                prof_state[name] = (profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]), time.perf_counter())
                prof_state[name][0].__enter__()

        def _stop(name):
            if name in prof_state and prof_state.get("armed"):
                torch.cuda.synchronize()
                pr, t0 = prof_state.pop(name)
                wall = time.perf_counter() - t0
                pr.__exit__(None, None, None)
                rows = []
                for e in pr.key_averages():
                    t = getattr(e, "self_device_time_total", 0) or 0
                    if t > 0:
                        rows.append((t, e.count, e.key))
                rows.sort(reverse=True)
                tot = sum(r[0] for r in rows)
                cpu_top = sorted(((e.self_cpu_time_total, e.count, e.key) for e in pr.key_averages()), reverse=True)[:12]
                print(f"[profile {name}] wall {wall * 1e3:.1f} ms, sum of kernel time {tot / 1e3:.1f} ms, {sum(r[1] for r in rows)} launches", file=sys.stderr)
                for t, c, k in rows[:18]:
                    print(f"    {t / 1e3:9.2f} ms  x{c:<5d} {k[:110]}", file=sys.stderr)
                print("    -- host side (self CPU time):", file=sys.stderr)
                for t, c, k in cpu_top:
                    print(f"    {t / 1e3:9.2f} ms  x{c:<5d} {k[:110]}", file=sys.stderr)
        for name in args.profile_mfc.split(","):
            ex.hooks.setdefault(name, []).append(lambda n=name: _start(n))
            ex.post_hooks.setdefault(name, []).append(lambda n=name: _stop(n))
    for i in range(args.warmup):
        rec, st_a, st_c = one_step(i)
        if args.verbose and rank == 0:
            from realhf_b200.models import generation as _gen
            print(f"[warmup {i}] " + " ".join(f"{k}={v.device_ms:.0f}ms" for k, v in rec.items())
                  + (f" gen_phases={_gen.LAST_TIMING}" if _gen.LAST_TIMING else ""), file=sys.stderr, flush=True)
    sampler = ClockSampler(world) if local_rank == 0 else _NoSampler()
    barrier()
    sampler.start()
    launches.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    mfc_ms = {}
    n_tokens_total = 0.0
    for i in range(args.steps):
        prof_state["armed"] = bool(args.profile_mfc) and i == args.steps - 1
        rec, st_a, st_c = one_step(args.warmup + i)
        for k, v in rec.items():
            mfc_ms[k] = mfc_ms.get(k, 0.0) + v.device_ms / args.steps
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    dev_s = e0.elapsed_time(e1) / 1e3
    clocks = sampler.stop()
    n_launch = launches.total
    unckpt = {r: getattr(getattr(models[r].module, "module", None), "last_unckpt_blocks", 0) for r in ("actor", "critic")}
    _o = models["actor"].module.optim
    zero_comm = ("none (dp=1)" if world == 1 else
                 (f"NVLS: multimem.ld_reduce reduce-scatter fused with avg/cast/grad-norm + AdamW with multimem.st all-gather, "
                  f"{len(_o.buckets)} buckets, {_o.n_overlapped} reduced inside backward" if _o.nvls else
                  f"NCCL in-place reduce-scatter / all-gather, {len(_o.buckets)} buckets, {_o.n_overlapped} reduced inside backward"))
    pool = ex.last_pool
    tokens_this_rank = float(sum(pool.flat_seqlens("packed_input_ids")))
    t = torch.tensor([dev_s, wall, tokens_this_rank, float(n_launch)], dtype=torch.float64, device=dev)
    tmax = t.clone()
    # per-MFC device time: rank 0's own clock, plus min / max over ranks (collective-free MFCs such as DP generation finish at
    # different times on different GPUs; the skew is absorbed by the first collective of the next training MFC)
    names = sorted(mfc_ms)
    mv = torch.tensor([mfc_ms[k] for k in names], dtype=torch.float64, device=dev)
    mv_max, mv_min = mv.clone(), mv.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(mv_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(mv_min, op=dist.ReduceOp.MIN)
    mfc_minmax = {k: [round(float(a), 1), round(float(b), 1)] for k, a, b in zip(names, mv_min, mv_max)}
    dev_s, wall = float(tmax[0]), float(tmax[1])
    tokens_per_step = float(t[2]) if world > 1 else tokens_this_rank
    value = tokens_per_step * args.steps / dev_s
    e2e = tokens_per_step * args.steps / wall
    headline = args.layers == 32 and args.prompts == 128 and args.prompt_len == 128 and args.new_tokens == 512 and not args.gen_fp8
    if rank == 0:
        out = {
            "metric": METRIC, "value": round(value, 1), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dev_s * 1e3 / args.steps, 1), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": round(value / BASELINE_TOKENS_PER_S, 3) if headline else None,
            "dtype": "bf16" if not args.gen_fp8 else "bf16 (training, inference, prefill) + e4m3 W8A8 decode [NOT the headline precision]",
            "data": "synthetic prompts (uniform random token ids), random-init weights",
            "config": {"model": "LLaMA-7B actor + 7B critic + 7B ref + 7B reward" + ("" if headline or args.gen_fp8 else f" [DEBUG layers={args.layers}]"),
                       "global_batch": args.prompts, "seq_len": args.prompt_len + args.new_tokens,
                       "prompt_len": args.prompt_len, "new_tokens": args.new_tokens, "ppo_minibatches": 4,
                       "parallelism": (f"dp{world} (all 6 MFCs)" if not (world > 1 and gen_tp > 1) else
                                       f"actor_gen tp{gen_tp}xdp{world // gen_tp} (realloc'd replica), other MFCs dp{world}") + ", ZeRO-1 flat AdamW", "tokens_per_step": tokens_per_step,
                       "optimizer": ("AdamW, fp32 master weights + fp32 moments (reference precision), bf16 grads" if args.optimizer == "fp32" else
                                     "AdamW, bf16 moments + stochastic rounding (no fp32 master), bf16 grads"),
                       "zero_comm": zero_comm, "gen_layout_search": gen_choice,
                       "decode_graph": "re-captured every step (KV cache freed before training)" if recapture else "captured once, KV cache resident",
                       "frozen_model_offload": args.offload_frozen == "on" or (args.offload_frozen == "auto" and world == 1 and args.ckpt == "auto"),
                       "activation_checkpointing": (f"auto: {unckpt} of {args.layers} blocks keep activations (free-HBM budget)" if args.ckpt == "auto"
                                                    else ("every block" if args.ckpt else "none")),
                       "gemm": args.gemm, "attention": "own tcgen05 varlen fwd/bwd + own split-KV decode kernel",
                       "l2": "working set >> L2 (54 GB weights per GPU); fresh inputs every step",
                       "mfc_ms": {k: round(v, 1) for k, v in mfc_ms.items()}, "mfc_ms_min_max_over_ranks": mfc_minmax},
            "clocks": clocks,
            "e2e": {"value": round(e2e, 1), "unit": "tokens/s", "h2d_bytes_per_step": h2d_bytes * world,
                    "d2h_bytes_per_step": result_host.numel() * 4 * world,
                    "timer": "host wall clock (perf_counter) around the same K steps driven through SPMDExecutor.run_step: "
                             "pinned-host prompt H2D, all six MFCs, loss/reward statistics D2H; `value` is the CUDA-event time",
                    "wall_ms_per_step": round(wall * 1e3 / args.steps, 2), "device_ms_per_step": round(dev_s * 1e3 / args.steps, 2)},
            "gpu_launches": int(float(t[3]) if world > 1 else n_launch),
            "memory": {"peak_allocated_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
                       "peak_reserved_gb": round(torch.cuda.max_memory_reserved() / 2 ** 30, 1),
                       "alloc_retries": int(torch.cuda.memory_stats().get("num_alloc_retries", 0)),
                       "allocator": os.environ.get("PYTORCH_CUDA_ALLOC_CONF", "")},
            "impl": "ours" if args.impl == "ours" else "torchlib (cuBLAS + flash-attn + NCCL on this repo's executor; NOT the reference)",
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
