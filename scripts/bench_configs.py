"""The secondary BASELINE.json configurations through the production runtime (quickstart -> launcher -> master + model workers),
at reduced depth, with random-init weights and synthetic data; prints ONE JSON line per run.

    python scripts/bench_configs.py 13b-tp2pp2  --gpus 4    # LLaMA-13B shapes, PPO, every MFC tp2 x pp2, CUDA-graph pipelined decode
    python scripts/bench_configs.py mixtral-ep   --gpus 2    # Mixtral-8x7B shapes, PPO, experts partitioned over the TP group
    python scripts/bench_configs.py dpo-zero3    --gpus 2    # LLaMA-7B shapes, DPO, ZeRO-3 (per-layer) + optimizer / parameter offload
    (add --tiny --device cpu for a CPU smoke run of the same plumbing)

"Reduced depth": `--layers` transformer blocks (default 8) of the named architecture; every other shape (hidden, FFN, heads, experts,
vocabulary, batch, sequence lengths) is the named one.  The numbers are recorded under profiles/ as *reduced-depth* runs.
"""
import argparse
import json
import os
import random
import re
import sys
import tempfile
import time
import uuid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _tokenizer(words):
    from tokenizers import Tokenizer, models, pre_tokenizers, trainers
    from transformers import PreTrainedTokenizerFast
    tk = Tokenizer(models.WordLevel(unk_token="[UNK]"))
    tk.pre_tokenizer = pre_tokenizers.Whitespace()
    tk.train_from_iterator([" ".join(words)], trainers.WordLevelTrainer(vocab_size=1000, special_tokens=["[PAD]", "[EOS]", "[UNK]"]))
    return PreTrainedTokenizerFast(tokenizer_object=tk, pad_token="[PAD]", eos_token="[EOS]", unk_token="[UNK]")


def _model_dir(root, name, family, cfg, tok):
    from realhf_b200.models import hf_io
    d = os.path.join(root, name)
    os.makedirs(d)
    fam = hf_io.family(family)
    hf_cfg = fam.config_to_hf(cfg)
    hf_cfg.architectures = [fam.hf_cls_name]
    hf_cfg.save_pretrained(d)
    tok.save_pretrained(d)
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["13b-tp2pp2", "mixtral-ep", "dpo-zero3"])
    ap.add_argument("--gpus", type=int, required=True)
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--prompts", type=int, default=128)
    ap.add_argument("--prompt-len", type=int, default=128)
    ap.add_argument("--new-tokens", type=int, default=512)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--gen-graph", choices=["true", "false"], default="true", help="decode loop inside a CUDA graph")
    args = ap.parse_args()

    root = tempfile.mkdtemp(prefix="realhf_b200_cfgbench_")
    os.environ.setdefault("REAL_FILEROOT", os.path.join(root, "fileroot"))  # before the package import: constants are resolved then
    os.environ["PYTHONPATH"] = ROOT + os.pathsep + os.environ.get("PYTHONPATH", "")
    os.environ["REAL_FAST_INIT"] = "1"
    from realhf_b200.api.model import ReaLModelConfig, ReaLMoEConfig
    from realhf_b200.apps.main import main_start
    from realhf_b200.apps.quickstart import build_experiment
    rng = random.Random(0)
    words = ["".join(rng.choice("abcdefghijklmnop") for _ in range(rng.randint(2, 6))) for _ in range(400)]
    tok = _tokenizer(words)
    n, L = args.gpus, args.layers
    dev = args.device
    dtype = "bf16" if dev == "cuda" else "fp32"
    base = dict(n_positions=4096, embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, layer_norm_epsilon=1e-5, activation_function="silu",
                scale_attn_by_inverse_layer_idx=False, use_attention_bias=False, use_attn_proj_bias=False, use_mlp_bias=False,
                layer_norm_type="rms", apply_rotary=True, vocab_size=32000)
    if args.config == "13b-tp2pp2":
        shape = dict(n_layers=L, hidden_dim=5120, intermediate_dim=13824, n_q_heads=40, n_kv_heads=40, mlp_type="llama")
        family, desc = "llama", f"LLaMA-13B shapes x{L} layers"
    elif args.config == "mixtral-ep":
        shape = dict(n_layers=L, hidden_dim=4096, intermediate_dim=14336, n_q_heads=32, n_kv_heads=8, mlp_type="moe",
                     moe=ReaLMoEConfig(num_experts=8, top_k=2, aux_loss_coeff=0.0))
        family, desc = "mixtral", f"Mixtral-8x7B shapes x{L} layers"
    else:
        shape = dict(n_layers=L, hidden_dim=4096, intermediate_dim=11008, n_q_heads=32, n_kv_heads=32, mlp_type="llama")
        family, desc = "llama", f"LLaMA-7B shapes x{L} layers"
    if args.tiny:
        shape.update(hidden_dim=256, intermediate_dim=512, n_q_heads=4, n_kv_heads=4, head_dim=64)
        base["vocab_size"] = 1024
    dirs = {}
    for role, critic in (("actor", False), ("critic", True)):
        cfg = ReaLModelConfig(**base, **shape, is_critic=critic)
        dirs[role] = _model_dir(root, role, family, cfg, tok)
    total = args.warmup + args.steps
    exp_name = f"cfg-{args.config.replace('-', '')}-{uuid.uuid4().hex[:5]}"
    common = [f"experiment_name={exp_name}", "trial_name=t0", f"device={dev}", f"dtype={dtype}", f"n_gpus_per_node={n}",
              "exp_ctrl.total_train_epochs=1", f"exp_ctrl.benchmark_steps={total}"]
    tokens_per_step = args.prompts * (args.prompt_len + args.new_tokens)

    def model_args(role, path, train):
        a = [f"{role}.type._class={family}", f"{role}.path={path}", f"{role}.init_from_scratch=true",
             f"{role}.gradient_checkpointing=true"]
        if train:
            a += [f"{role}.optimizer.lr_scheduler_type=constant", f"{role}.optimizer.warmup_steps_proportion=0.0",
                  f"{role}.optimizer.grad_dtype={'bf16' if dev == 'cuda' else 'fp32'}"]
            if dev == "cuda":
                a += [f"{role}.optimizer.state_dtype=bf16", f"{role}.optimizer.use_master_weights=false"]
        if args.config == "mixtral-ep":
            a += [f"{role}.expert_parallel=true"]
        return a

    if args.config in ("13b-tp2pp2", "mixtral-ep"):
        data = os.path.join(root, "prompts.jsonl")
        with open(data, "w") as f:
            for i in range(args.prompts * total):
                f.write(json.dumps(dict(id=i, prompt=" ".join(rng.choice(words) for _ in range(2 * args.prompt_len)))) + "\n")
        qs = ["ppo"] + common + [f"dataset.path={data}", f"dataset.train_bs_n_seqs={args.prompts}", f"dataset.max_prompt_len={args.prompt_len}",
                                 "dataset.pad_to_max_length=true", f"ppo.gen.max_new_tokens={args.new_tokens}",
                                 f"ppo.gen.min_new_tokens={args.new_tokens}", "ppo.gen.top_p=0.9", "ppo.gen.top_k=1000",
                                 # MoE decode is not graph-capturable yet: the expert bucketing (bincount / nonzero) syncs the host
                                 f"ppo.gen.use_cuda_graph={args.gen_graph}",
                                 "ppo.gen.force_cudagraph_recapture=true", "ppo.ppo_n_minibatches=4"]
        for role, path, train in (("actor", dirs["actor"], True), ("ref", dirs["actor"], False), ("critic", dirs["critic"], True),
                                  ("rew", dirs["critic"], False)):
            qs += model_args(role, path, train)
        if args.config == "13b-tp2pp2":
            assert n % 4 == 0
            qs += [f"allocation_mode=d{n // 4}m2p2"]
            par = f"every MFC dp{n // 4} x tp2 x pp2 (SP in training), CUDA-graph decode per (stage, micro-batch)"
        else:
            qs += ["allocation_mode=manual"]
            for mfc in ("actor_gen", "actor_train", "critic_train", "critic_inf", "ref_inf", "rew_inf"):
                qs += [f"{mfc}.parallel.model_parallel_size={n}", f"{mfc}.parallel.data_parallel_size=1"]
            for mfc in ("actor_train", "critic_train"):
                qs += [f"{mfc}.parallel.use_sequence_parallel=true"]
            par = f"every MFC tp{n} with the 8 experts partitioned over the {n} ranks (EP); training with SP: peer-store dispatch / combine"
        metric = "PPO tokens/sec (gen + inference + train, master host clock)"
    else:
        data = os.path.join(root, "pairs.jsonl")
        half = (args.prompt_len + args.new_tokens) // 2
        with open(data, "w") as f:
            for i in range(args.prompts * total):
                mk = lambda k: " ".join(rng.choice(words) for _ in range(k))
                f.write(json.dumps(dict(id=i, prompt=mk(args.prompt_len) + " ", pos_answers=[mk(half)], neg_answers=[mk(half)])) + "\n")
        qs = ["dpo"] + common + ["allocation_mode=manual", f"dataset.train_path={data}", f"dataset.train_bs_n_seqs={args.prompts}",
                                 f"dataset.max_seqlen={args.prompt_len + args.new_tokens}", "dataset.max_pairs_per_prompt=1"]
        qs += model_args("actor", dirs["actor"], True) + model_args("ref", dirs["actor"], False)
        qs += ["actor.zero_stage=3", "actor.offload=true", "actor.optimizer.offload_param=true", "actor.optimizer.state_dtype=fp32",
               "actor.optimizer.use_master_weights=true", "actor.optimizer.grad_dtype=fp32", "ref.offload=true",
               f"actor_train.parallel.data_parallel_size={n}", f"ref_inf.parallel.data_parallel_size={n}"]
        par = f"dp{n}, ZeRO-3 per-layer gather / release, fp32 master + moments and the parameter shard in pinned host memory, reference offloaded"
        tokens_per_step = 2 * args.prompts * (args.prompt_len + args.new_tokens)  # one positive + one negative answer per prompt (upper bound)
        metric = "DPO tokens/sec (ref inference + train, master host clock)"
    exp = build_experiment(qs)
    t0 = time.perf_counter()
    main_start(exp, timeout=3000)
    wall = time.perf_counter() - t0
    log = open(os.path.join(os.environ["REAL_FILEROOT"], "logs", exp_name, "t0", "master_worker-0")).read()
    steps = [(float(m.group(1)), m.group(2)) for m in re.finditer(r"step \d+ \(epoch[^)]*\) e2e ([0-9.]+)s; (.*)", log)]
    assert len(steps) >= total, f"master log has {len(steps)} steps, expected {total}:\n{log[-3000:]}"
    timed = steps[args.warmup: total]
    secs = sum(t for t, _ in timed)
    mfc = {}
    for _, line in timed:
        for part in line.split(", "):
            k, v = part.rsplit(" ", 1)
            mfc[k] = round(mfc.get(k, 0.0) + float(v.rstrip("s")) * 1e3 / len(timed), 1)
    tok_m = re.findall(r"throughput: (\d+) tokens", log)
    if tok_m:
        tokens_per_step = int(tok_m[-1])
    print(json.dumps({"config": args.config, "model": desc + (" [tiny debug shapes]" if args.tiny else ""), "reduced_depth": True, "metric": metric,
                      "value": round(tokens_per_step * len(timed) / secs, 1), "unit": "tokens/s", "n_gpus": n, "steps": len(timed), "warmup": args.warmup,
                      "ms_per_step": round(secs * 1e3 / len(timed), 1), "tokens_per_step": tokens_per_step, "parallelism": par, "dtype": dtype,
                      "mfc_ms": mfc, "data": "synthetic, random-init weights", "runtime": "master/worker (quickstart -> launcher)",
                      "launch_to_exit_s": round(wall, 1)}), flush=True)


if __name__ == "__main__":
    main()
