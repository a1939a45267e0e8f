"""TP / SP / PP / DP(ZeRO-1) equivalence on CPU with gloo: a sharded model must reproduce the single-process loss,
gradients' effect (updated weights) and greedy generation."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.distributed
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _batch(bs=8, seed=0, vocab=128):
    from realhf_b200.api.data import SequenceSample
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(5, 14, (bs,), generator=g).tolist()
    ids = torch.randint(2, vocab, (sum(lens),), generator=g)
    pm = torch.zeros(sum(lens), dtype=torch.bool)
    off = 0
    for l in lens:
        pm[off:off + 2] = True
        off += l
    return SequenceSample.from_default(seqlens=lens, ids=list(range(bs)), data=dict(packed_input_ids=ids, prompt_mask=pm))


def _worker(rank, world, layout, fam, n_steps, n_mbs=1, device="cpu", dtype=torch.float32, pg_backend="gloo", opt_extra=None):
    import types

    from realhf_b200.api.config import ModelName
    from realhf_b200.api.model import FinetuneSpec, GenerationHyperparameters, Model
    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.engine.engine import TrainBackend
    from realhf_b200.interfaces import basic
    from realhf_b200.models import hf_io
    from realhf_b200.models.real_model import ReaLModel
    from realhf_b200.api.data import SequenceSample
    pp, dp, tp, sp = layout
    cfg = hf_io.family(fam).make_test_config()
    cfg.n_layers = 4
    if device == "cuda":
        torch.cuda.set_device(rank)
        device = torch.device("cuda", rank)
    ctx = ParallelContext.build(ProcessTopology(pp, dp, tp), list(range(world)), rank, backend=pg_backend, sequence_parallel=sp)
    m = ReaLModel(cfg, ctx, dtype=dtype, device=torch.device(device)).instantiate(seed=7)
    tok = types.SimpleNamespace(eos_token_id=1, pad_token_id=0)
    model = TrainBackend(optimizer=dict(lr=1e-2, weight_decay=0.0, warmup_steps_proportion=0.0, lr_scheduler_type="constant",
                                        grad_dtype="fp32", gradient_clipping=1.0, **(opt_extra or {}))).initialize(Model(ModelName("m", 0), m, tok, device),
                                                                                               FinetuneSpec(1, 10, 10))
    full = _batch(8)
    mine = full.split(dp)[ctx.dp_rank] if dp > 1 else full
    mine.to_device(device)
    itf = basic.SFTInterface()
    losses = [itf.train_step(model, mine, n_mbs=n_mbs)["loss"] for _ in range(n_steps)]
    # greedy generation from the updated weights
    g = GenerationHyperparameters(max_new_tokens=5, min_new_tokens=5, greedy=True)
    plens = [4, 6, 5, 3][: max(2, 4)]
    prompts = SequenceSample.from_default(seqlens=plens, ids=list(range(len(plens))),
                                          data=dict(packed_input_ids=(torch.arange(2, 2 + sum(plens)) % cfg.vocab_size).to(device)))
    outs = model.module.generate(prompts, tok, g, num_micro_batches=1)
    gen_tokens = torch.cat([o.tokens for o in outs]).tolist() if outs is not None else None
    opt = model.module.optim
    return dict(losses=losses, gen=gen_tokens, coord=tuple(ctx.coord), n_buckets=len(opt.buckets), n_overlapped=opt.n_overlapped)


def _reference(fam, n_steps, n_mbs=1):
    """Single process with the same micro-batch partition as the sharded run (losses are means per micro-batch)."""
    from realhf_b200.base.testing import run_distributed
    return run_distributed(_worker, 1, layout=(1, 1, 1, False), fam=fam, n_steps=n_steps, n_mbs=n_mbs)[0]


@pytest.mark.parametrize("layout", [(1, 1, 2, False), (1, 1, 2, True), (2, 1, 1, False), (1, 2, 1, False), (2, 1, 2, True), (2, 2, 1, False),
                                    (2, 2, 2, True), (1, 1, 4, True), (4, 1, 1, False), (1, 2, 2, False)])
def test_layout_matches_single_process(layout):
    from realhf_b200.base.testing import run_distributed
    fam, n_steps = "llama", 3
    pp, dp, tp, sp = layout
    ref = _reference(fam, n_steps, n_mbs=dp * (2 * pp if pp > 1 else 1))
    res = run_distributed(_worker, pp * dp * tp, layout=layout, fam=fam, n_steps=n_steps)
    for r in res:
        for a, b in zip(r["losses"], ref["losses"]):
            assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (layout, r["losses"], ref["losses"])
    gens = [r["gen"] for r in res if r["gen"] is not None]
    assert gens and all(g == ref["gen"] for g in gens), (gens, ref["gen"])


@pytest.mark.parametrize("cfg", [(2, 1, 1), (2, 2, 1), (4, 1, 1), (2, 1, 2)])
def test_bucketed_overlapped_reduce_scatter_matches_single_process(cfg):
    """ZeRO-1 with many small gradient buckets: buckets are reduced from inside the backward pass of the last micro-batch
    (`FlatAdamW.grad_ready`), every rank owns a slice of every bucket, and the result equals the single-process run."""
    from realhf_b200.base.testing import run_distributed
    dp, n_mbs, tp = cfg
    fam, n_steps = "llama", 3
    ref = _reference(fam, n_steps, n_mbs=dp * n_mbs)
    res = run_distributed(_worker, dp * tp, layout=(1, dp, tp, False), fam=fam, n_steps=n_steps, n_mbs=n_mbs, opt_extra=dict(bucket_numel=2048))
    for r in res:
        assert r["n_buckets"] >= 4, r
        assert r["n_overlapped"] >= r["n_buckets"] // 2, r   # everything above the embedding starts inside backward
        for a, b in zip(r["losses"], ref["losses"]):
            assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (cfg, r["losses"], ref["losses"])
    gens = [r["gen"] for r in res if r["gen"] is not None]
    assert gens and all(g == ref["gen"] for g in gens), (gens, ref["gen"])


def test_gpt2_tied_embedding_pp2():
    from realhf_b200.base.testing import run_distributed
    ref = _reference("gpt2", 2, n_mbs=4)
    res = run_distributed(_worker, 2, layout=(2, 1, 1, False), fam="gpt2", n_steps=2)
    for r in res:
        for a, b in zip(r["losses"], ref["losses"]):
            assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (r["losses"], ref["losses"])


def _moe_worker(rank, world, layout, expert_parallel, device="cpu", dtype=torch.float32, pg_backend="gloo"):
    import types

    from realhf_b200.api.config import ModelName
    from realhf_b200.api.model import FinetuneSpec, Model
    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.engine.engine import TrainBackend
    from realhf_b200.interfaces import basic
    from realhf_b200.models import hf_io
    from realhf_b200.models.real_model import ReaLModel
    pp, dp, tp, sp = layout
    cfg = hf_io.family("mixtral").make_test_config()
    cfg.moe.expert_parallel = expert_parallel
    cfg.moe.aux_loss_coeff = 0.0
    if device == "cuda":
        torch.cuda.set_device(rank)
        device = torch.device("cuda", rank)
    ctx = ParallelContext.build(ProcessTopology(pp, dp, tp), list(range(world)), rank, backend=pg_backend, sequence_parallel=sp)
    m = ReaLModel(cfg, ctx, dtype=dtype, device=torch.device(device)).instantiate(seed=7)
    tok = types.SimpleNamespace(eos_token_id=1, pad_token_id=0)
    model = TrainBackend(optimizer=dict(lr=1e-2, weight_decay=0.0, warmup_steps_proportion=0.0, lr_scheduler_type="constant",
                                        grad_dtype="fp32")).initialize(Model(ModelName("m", 0), m, tok, "cpu"), FinetuneSpec(1, 10, 10))
    itf = basic.SFTInterface()
    return [itf.train_step(model, _batch(8), n_mbs=1)["loss"] for _ in range(3)]


@pytest.mark.parametrize("cfg", [((1, 1, 2, False), False), ((1, 1, 2, False), True), ((1, 1, 2, True), True), ((1, 1, 4, True), True)])
def test_moe_tp_and_expert_parallel_match_single_process(cfg):
    from realhf_b200.base.testing import run_distributed
    layout, ep = cfg
    ref = run_distributed(_moe_worker, 1, layout=(1, 1, 1, False), expert_parallel=False)[0]
    res = run_distributed(_moe_worker, layout[2], layout=layout, expert_parallel=ep)
    for r in res:
        for a, b in zip(r, ref):
            assert abs(a - b) < 3e-3 * max(1.0, abs(b)), (cfg, r, ref)


def _zero3_worker(rank, world, zero_stage, n_mbs=1, grad_dtype="fp32"):
    import types

    from realhf_b200.api.config import ModelName
    from realhf_b200.api.data import SequenceSample
    from realhf_b200.api.model import FinetuneSpec, GenerationHyperparameters, Model
    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.engine.engine import TrainBackend
    from realhf_b200.interfaces import basic
    from realhf_b200.models import hf_io
    from realhf_b200.models.real_model import ReaLModel
    cfg = hf_io.family("llama").make_test_config()
    ctx = ParallelContext.build(ProcessTopology(1, world, 1), list(range(world)), rank, backend="gloo")
    m = ReaLModel(cfg, ctx, dtype=torch.float32).instantiate(seed=7)
    n_params = m.flat_numel
    tok = types.SimpleNamespace(eos_token_id=1, pad_token_id=0)
    model = TrainBackend(optimizer=dict(lr=1e-2, weight_decay=0.0, warmup_steps_proportion=0.0, lr_scheduler_type="constant", grad_dtype=grad_dtype),
                         zero_stage=zero_stage).initialize(Model(ModelName("m", 0), m, tok, "cpu"), FinetuneSpec(1, 10, 10))
    resident = model.module.module.instantiated
    mine = _batch(12).split(world)[rank]
    losses = [basic.SFTInterface().train_step(model, mine, n_mbs=n_mbs)["loss"] for _ in range(3)]
    opt = model.module.optim
    z3 = opt.z3
    # generation needs every parameter at once: the optimizer materialises the full buffer, the next call drops it again
    g = GenerationHyperparameters(max_new_tokens=4, min_new_tokens=4, greedy=True)
    prompts = SequenceSample.from_default(seqlens=[4, 6], ids=[0, 1], data=dict(packed_input_ids=(torch.arange(2, 12) % cfg.vocab_size)))
    gen_tokens = torch.cat([o.tokens for o in model.module.generate(prompts, tok, g, num_micro_batches=1)]).tolist()
    logits = model.module.forward(prompts, num_micro_batches=1).float()
    return dict(losses=losses, resident_between_calls=resident and model.module.module.instantiated, gen=gen_tokens,
                logit_sum=float(logits.sum()), per_layer=z3 is not None, n_gathers=0 if z3 is None else z3.n_gathers,
                shard_elems=opt.m.numel(), n_params=n_params, full_grad_buffer=opt.flat_grad is not None)


@pytest.mark.parametrize("cfg", [(2, 1), (3, 2)])
def test_zero3_matches_zero1_and_releases_params(cfg):
    """ZeRO-3 with per-layer gather / release (engine/zero3.py): same losses, generations and logits as ZeRO-1, parameters never
    resident between calls, no full-size gradient buffer, optimizer state ~1/dp of the model; dp=3 exercises uneven slices and
    n_mbs=2 the accumulation of reduce-scattered gradients over micro-batches."""
    from realhf_b200.base.testing import run_distributed
    world, n_mbs = cfg
    z1 = run_distributed(_zero3_worker, world, zero_stage=1, n_mbs=n_mbs)
    z3 = run_distributed(_zero3_worker, world, zero_stage=3, n_mbs=n_mbs)
    assert z1[0]["resident_between_calls"] and not z3[0]["resident_between_calls"]
    assert all(r["per_layer"] and r["n_gathers"] > 20 and not r["full_grad_buffer"] for r in z3)
    assert z3[0]["shard_elems"] <= z3[0]["n_params"] // world + 64 * 16
    for r1, r3 in zip(z1, z3):
        for a, b in zip(r1["losses"], r3["losses"]):
            assert abs(a - b) < 1e-5
        assert r1["gen"] == r3["gen"]
        assert abs(r1["logit_sum"] - r3["logit_sum"]) < 1e-3 * max(1.0, abs(r1["logit_sum"]))


def _dropout_worker(rank, world, ckpt):
    """GPT-2 with dropout 0.2 in training mode, tp=2 without sequence parallelism: activations are replicated over the TP
    group, so the dropout masks (and hence losses / updated weights / sampled tokens) must agree across the group."""
    import types

    from realhf_b200.api.config import ModelName
    from realhf_b200.api.model import FinetuneSpec, GenerationHyperparameters, Model
    from realhf_b200.base import seeding
    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.engine.engine import TrainBackend
    from realhf_b200.interfaces import basic
    from realhf_b200.models import hf_io
    from realhf_b200.models.real_model import ReaLModel
    from realhf_b200.api.data import SequenceSample
    seeding.set_random_seed(3, offset=rank)          # what a model worker does: per-rank global generators
    cfg = hf_io.family("gpt2").make_test_config()
    cfg.n_layers = 3
    cfg.resid_pdrop = cfg.embd_pdrop = 0.2
    cfg.attn_pdrop = 0.0
    ctx = ParallelContext.build(ProcessTopology(1, 1, world), list(range(world)), rank, backend="gloo", sequence_parallel=False,
                                gradient_checkpointing=ckpt)
    m = ReaLModel(cfg, ctx, dtype=torch.float32, device=torch.device("cpu")).instantiate(seed=7)
    tok = types.SimpleNamespace(eos_token_id=1, pad_token_id=0)
    model = TrainBackend(optimizer=dict(lr=1e-2, weight_decay=0.0, warmup_steps_proportion=0.0, lr_scheduler_type="constant",
                                        grad_dtype="fp32", gradient_clipping=1.0)).initialize(Model(ModelName("m", 0), m, tok, "cpu"),
                                                                                               FinetuneSpec(1, 10, 10))
    batch = _batch(8)
    itf = basic.SFTInterface()
    losses = [itf.train_step(model, batch, n_mbs=1)["loss"] for _ in range(3)]
    # a replicated (not TP-sharded) parameter after three dropout-perturbed updates
    ln = m.p["1.attn.ln.weight"].detach().tolist()   # plain lists: tensors in an mp.Queue need the sender alive
    g = GenerationHyperparameters(max_new_tokens=6, min_new_tokens=6, greedy=False, top_k=50, temperature=1.0)
    plens = [4, 6, 5]
    prompts = SequenceSample.from_default(seqlens=plens, ids=list(range(len(plens))),
                                          data=dict(packed_input_ids=(torch.arange(2, 2 + sum(plens)) % cfg.vocab_size)))
    outs = model.module.generate(prompts, tok, g, num_micro_batches=1)
    return dict(losses=losses, ln=ln, gen=torch.cat([o.tokens for o in outs]).tolist())


@pytest.mark.parametrize("ckpt", [False, True])
def test_replicated_dropout_and_sampling_agree_across_tp_ranks(ckpt):
    from realhf_b200.base.testing import run_distributed
    a, b = run_distributed(_dropout_worker, 2, ckpt=ckpt)
    assert a["losses"] == b["losses"], (a["losses"], b["losses"])
    assert a["ln"] == b["ln"]
    assert a["gen"] == b["gen"]
    if ckpt:   # recomputation replays the same masks: identical training trajectory with and without checkpointing
        c, _ = run_distributed(_dropout_worker, 2, ckpt=False)
        assert c["losses"] == pytest.approx(a["losses"], rel=1e-5)


def _zero3_layout_worker(rank, world, layout, zero_stage):
    import types

    from realhf_b200.api.config import ModelName
    from realhf_b200.api.model import FinetuneSpec, Model
    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.engine.engine import TrainBackend
    from realhf_b200.interfaces import basic
    from realhf_b200.models import hf_io
    from realhf_b200.models.real_model import ReaLModel
    pp, dp, tp, sp = layout
    cfg = hf_io.family("llama").make_test_config()
    cfg.n_layers = 4
    ctx = ParallelContext.build(ProcessTopology(pp, dp, tp), list(range(world)), rank, backend="gloo", sequence_parallel=sp)
    m = ReaLModel(cfg, ctx, dtype=torch.float32).instantiate(seed=7)
    tok = types.SimpleNamespace(eos_token_id=1, pad_token_id=0)
    model = TrainBackend(optimizer=dict(lr=1e-2, weight_decay=0.0, warmup_steps_proportion=0.0, lr_scheduler_type="constant", grad_dtype="fp32"),
                         zero_stage=zero_stage).initialize(Model(ModelName("m", 0), m, tok, "cpu"), FinetuneSpec(1, 10, 10))
    full = _batch(8)
    mine = full.split(dp)[ctx.dp_rank] if dp > 1 else full
    losses = [basic.SFTInterface().train_step(model, mine, n_mbs=2 if pp > 1 else 1)["loss"] for _ in range(3)]
    return dict(losses=losses, resident=model.module.module.instantiated)


@pytest.mark.parametrize("layout", [(1, 2, 2, False), (1, 2, 2, True), (2, 2, 1, False)])
def test_zero3_composes_with_tensor_sequence_and_pipeline_parallelism(layout):
    """The parameter shards of ZeRO-3 are cut over the DATA-parallel group of each (pp, tp) coordinate: the same losses as ZeRO-1
    under dp2 x tp2 (with and without sequence parallelism) and dp2 x pp2, and no resident parameters between calls."""
    from realhf_b200.base.testing import run_distributed
    world = layout[0] * layout[1] * layout[2]
    z1 = run_distributed(_zero3_layout_worker, world, layout=layout, zero_stage=1)
    z3 = run_distributed(_zero3_layout_worker, world, layout=layout, zero_stage=3)
    for a, b in zip(z1, z3):
        assert a["resident"] and not b["resident"]
        for x, y in zip(a["losses"], b["losses"]):
            assert abs(x - y) < 1e-4, (layout, a["losses"], b["losses"])
