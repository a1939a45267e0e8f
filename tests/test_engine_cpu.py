"""Single-process CPU end-to-end: engines + interfaces on tiny models (SFT converges; one full PPO iteration runs)."""
import types

import pytest
import torch

from realhf_b200.api.config import ModelName
from realhf_b200.api.data import SequenceSample
from realhf_b200.api.model import FinetuneSpec, Model
from realhf_b200.engine.engine import InferenceBackend, TrainBackend
from realhf_b200.interfaces import basic, ppo  # noqa: F401
from realhf_b200.models import hf_io
from realhf_b200.models.real_model import ReaLModel

TOK = types.SimpleNamespace(eos_token_id=1, pad_token_id=0)


def make_model(role, critic=False, train=True, seed=1, lr=1e-2):
    cfg = hf_io.family("llama").make_test_config()
    cfg.is_critic = critic
    m = ReaLModel(cfg, dtype=torch.float32).instantiate(seed=seed)
    model = Model(ModelName(role, 0), m, TOK, "cpu")
    be = TrainBackend(optimizer=dict(lr=lr, weight_decay=0.0, warmup_steps_proportion=0.0, lr_scheduler_type="constant",
                                     grad_dtype="fp32")) if train else InferenceBackend()
    return be.initialize(model, FinetuneSpec(1, 100, 100))


def sft_batch(bs=8, seed=0):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(6, 20, (bs,), generator=g).tolist()
    ids = torch.randint(2, 128, (sum(lens),), generator=g)
    pm = torch.zeros(sum(lens), dtype=torch.bool)
    off = 0
    for l in lens:
        pm[off:off + 3] = True
        off += l
    return SequenceSample.from_default(seqlens=lens, ids=list(range(bs)), data=dict(packed_input_ids=ids, prompt_mask=pm))


def test_sft_loss_decreases():
    model = make_model("default")
    itf = basic.SFTInterface()
    batch = sft_batch()
    losses = [itf.train_step(model, batch, n_mbs=2)["loss"] for _ in range(8)]
    assert losses[-1] < losses[0] * 0.7, losses
    assert model.version.global_step == 8


def test_train_microbatching_equivalence():
    """n_mbs=1 and n_mbs=4 must produce (nearly) the same update for a mean-per-microbatch loss with equal weights."""
    batch = sft_batch(bs=8, seed=3)
    m1, m4 = make_model("a", lr=1e-3), make_model("b", lr=1e-3)
    basic.SFTInterface().train_step(m1, batch, n_mbs=1)
    basic.SFTInterface().train_step(m4, batch, n_mbs=1)
    for k in m1.module.module.p:
        torch.testing.assert_close(m1.module.module.p[k], m4.module.module.p[k])


def test_full_ppo_iteration_cpu():
    torch.manual_seed(0)
    actor, critic = make_model("actor"), make_model("critic", critic=True)
    ref, rew = make_model("ref", train=False), make_model("reward", critic=True, train=False, seed=7)
    gcfg = dict(max_new_tokens=8, min_new_tokens=2, top_k=50, top_p=0.9, temperature=1.0)
    a_itf = ppo.PPOActorInterface(n_minibatches=2, generation_config=gcfg, value_norm=True)
    c_itf = ppo.PPOCriticInterface(n_minibatches=2, value_norm=True)
    r_itf = basic.PairedRewardInterface()
    bs = 6
    plens = [5, 7, 4, 6, 5, 8]
    prompts = SequenceSample.from_default(seqlens=plens, ids=list(range(bs)),
                                          data=dict(packed_prompts=torch.randint(2, 128, (sum(plens),))))
    gen_out = a_itf.generate(actor, prompts, n_mbs=2)
    assert gen_out.bs == bs
    L = gen_out.flat_seqlens("packed_input_ids")
    assert all(l >= p + 2 for l, p in zip(L, plens))
    assert gen_out.data["packed_logprobs"].shape[0] == sum(L) - bs
    assert gen_out.data["packed_logits_mask"].shape[0] == sum(L) - bs
    # generation log-probs must agree with a recomputation under the same mask / temperature by the same weights
    re = a_itf.inference(actor, gen_out)
    gmask = (~gen_out.data["prompt_mask"])
    lp_gen = gen_out.data["packed_logprobs"]
    torch.testing.assert_close(re.data["packed_ref_logprobs"][lp_gen != 0], lp_gen[lp_gen != 0], atol=1e-4, rtol=1e-3)
    data = gen_out
    data.update_(a_itf.inference(ref, gen_out))
    data.update_(c_itf.inference(critic, gen_out))
    data.update_(r_itf.inference(rew, gen_out))
    assert data.data["values"].shape[0] == sum(L) and data.data["rewards"].shape[0] == bs
    before = actor.module.module.flat_param.data.clone()
    st_a = a_itf.train_step(actor, data, n_mbs=1)
    st_c = c_itf.train_step(critic, data, n_mbs=1)
    assert not torch.equal(before, actor.module.module.flat_param.data)
    for k in ("task_reward", "kl_reward", "actor_loss", "importance_weight", "grad_norm"):
        assert k in st_a and st_a[k] == st_a[k], (k, st_a)
    assert "value_loss" in st_c
    # first minibatch of the first step is on-policy: importance weight ~ 1
    assert abs(st_a["importance_weight"] - 1.0) < 0.2


def test_dpo_and_rw_train_steps():
    torch.manual_seed(0)
    actor, ref = make_model("actor"), make_model("ref", train=False)
    bs = 4
    seqlens = [[9, 11], [7, 8], [10, 6], [12, 9]]
    flat = [x for l in seqlens for x in l]
    ids = torch.randint(2, 128, (sum(flat),))
    pm = torch.zeros(sum(flat), dtype=torch.bool)
    off = 0
    for l in flat:
        pm[off:off + 3] = True
        off += l
    with SequenceSample.disable_validation():
        pass
    batch = SequenceSample(keys=["packed_input_ids", "prompt_mask"], ids=list(range(bs)),
                           seqlens=dict(packed_input_ids=seqlens, prompt_mask=seqlens),
                           trailing_shapes=dict(packed_input_ids=(), prompt_mask=()),
                           dtypes=dict(packed_input_ids=torch.long, prompt_mask=torch.bool),
                           data=dict(packed_input_ids=ids, prompt_mask=pm))
    d = basic.DPOInterface(beta=0.1)
    batch.update_(d.inference(ref, batch))
    st = d.train_step(actor, batch)
    assert abs(st["loss"] - 0.6931) < 0.05  # identical actor/ref at init -> log 2
    rw = make_model("rw", critic=True)
    st = basic.PairedRewardInterface().train_step(rw, batch)
    assert 0.0 <= st["acc"] <= 1.0 and st["loss"] > 0


def test_partial_activation_checkpointing_matches_full():
    """Keeping the activations of the last k blocks (budgeted checkpointing) must not change loss or gradients."""
    from realhf_b200.base.topology import ParallelContext
    cfg = hf_io.family("llama").make_test_config()
    batch = sft_batch(bs=4, seed=5)
    ids = batch.data["packed_input_ids"]
    lens = batch.flat_seqlens("packed_input_ids")
    cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
    cu[1:] = torch.tensor(lens).cumsum(0)
    grads = []
    for keep in (0, 1, cfg.n_layers):
        ctx = ParallelContext.single()
        ctx.gradient_checkpointing = "auto"
        m = ReaLModel(cfg, ctx, dtype=torch.float32).instantiate(seed=2)
        m.train()
        m._n_unckpt_blocks = lambda n_tokens, k=keep: k  # the CUDA path derives k from free HBM
        out = m(input_ids=ids, cu_seqlens=cu, max_seqlen=max(lens))
        loss = out.logits.float().square().mean()
        loss.backward()
        grads.append((loss.item(), torch.cat([p.grad.flatten() for p in m.parameters() if p.grad is not None])))
    for l, g in grads[1:]:
        assert abs(l - grads[0][0]) < 1e-6
        torch.testing.assert_close(g, grads[0][1], atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("variant", [
    dict(adaptive_kl_ctl=True), dict(value_norm=True, value_norm_type="ma"), dict(value_norm=True, adv_norm=False),
    dict(early_stop_imp_ratio=1.0001, early_stop_kl=1e-9), dict(discount=0.99, gae_lambda=0.95, eps_clip=0.1, max_reward_clip=0.5),
    dict(n_minibatches=3), dict(generation_config=dict(max_new_tokens=8, min_new_tokens=0, greedy=True)),
    dict(generation_config=dict(max_new_tokens=8, min_new_tokens=1, top_p=0.5, temperature=0.7, force_no_logits_mask=True))])
def test_ppo_hyperparameter_variants(variant):
    """Two PPO iterations under the option combinations of the reference's PPOHyperparameters (adaptive KL, both value
    normalisers, early stopping, GAE parameters, odd minibatch counts, greedy / unmasked generation)."""
    torch.manual_seed(0)
    actor, critic = make_model("actor"), make_model("critic", critic=True)
    ref, rew = make_model("ref", train=False), make_model("reward", critic=True, train=False, seed=7)
    kw = dict(n_minibatches=2, generation_config=dict(max_new_tokens=8, min_new_tokens=2, top_k=50, top_p=0.9))
    kw.update(variant)
    a = ppo.PPOActorInterface(**kw)
    c = ppo.PPOCriticInterface(**{k: v for k, v in kw.items() if k in ("n_minibatches", "kl_ctl", "discount", "gae_lambda", "value_eps_clip",
                                                                         "max_reward_clip", "adaptive_kl_ctl", "value_norm", "value_norm_type")})
    plens = [5, 7, 4, 6, 5, 8]
    prompts = SequenceSample.from_default(seqlens=plens, ids=list(range(6)), data=dict(packed_prompts=torch.randint(2, 128, (sum(plens),))))
    for _ in range(2):
        d = a.generate(actor, prompts, n_mbs=2)
        d.update_(a.inference(ref, d))
        d.update_(c.inference(critic, d))
        d.update_(basic.PairedRewardInterface().inference(rew, d))
        sa, sc = a.train_step(actor, d, n_mbs=2), c.train_step(critic, d, n_mbs=2)
        assert all(x == x for x in list(sa.values()) + list(sc.values()) if isinstance(x, float)), (sa, sc)


def test_fp8_generation_flag_is_a_noop_where_the_path_is_unsupported():
    """`fp8_weights=True` on a device / layout without the W8A8 kernels (CPU here; TP > 1, MoE) must generate exactly what the
    default path generates and leave no quantised copies behind."""
    import torch
    from realhf_b200.api.model import GenerationHyperparameters
    from realhf_b200.base.topology import ParallelContext
    from realhf_b200.models import generation as gen
    from realhf_b200.models import hf_io
    from realhf_b200.models.real_model import ReaLModel
    from realhf_b200.ops import fp8
    cfg = hf_io.family("llama").make_test_config()
    m = ReaLModel(cfg, ParallelContext.single(), dtype=torch.float32, device=torch.device("cpu")).instantiate(seed=3).eval()
    assert not m.fp8_decode_supported()
    ids = torch.randint(3, cfg.vocab_size, (12,))
    cu = torch.tensor([0, 5, 12], dtype=torch.int32)
    outs = []
    for flag in (False, True):
        g = GenerationHyperparameters(max_new_tokens=6, min_new_tokens=6, greedy=True, fp8_weights=flag)
        out, _ = gen.generate(m, ids, cu, g, eos_id=None, pad_id=0)
        outs.append(out)
        assert m._fp8 is None and not m._fp8_active
    assert torch.equal(outs[0].tokens, outs[1].tokens)
    torch.testing.assert_close(outs[0].logprobs, outs[1].logprobs)
    # the PyTorch quantisation rule itself: per-row scale, saturation, all-zero rows
    x = torch.tensor([[0.0, 0.0, 0.0, 0.0], [1.0, -448.0, 1000.0, 0.5]])
    q, s = fp8.quantize_rows_ref(x)
    assert s[0].item() == 1.0 and (q[0] == 0).all()
    d = fp8.dequantize(q, s)
    assert d[1, 2].item() == 1000.0 and abs(d[1, 1].item() + 448.0) < 448.0 * 0.07


@pytest.mark.parametrize("fam", ["llama", "qwen2", "mistral", "gemma"])
def test_w8a8_decode_wiring_under_cpu_emulation(fam, monkeypatch):
    """`REAL_FP8_EMULATE=1` runs the W8A8 decode path with every fp8 piece in PyTorch (same quantisation rule): which weights
    are quantised, where the qkv bias (qwen2), the norm offset and GeGLU (gemma), grouped KV heads (mistral) and the tied LM head
    (gemma) enter.  The generated log-probs must track the unquantised model's within the quantisation noise, every block
    linear and the head must go through the e4m3 product in every decode step, and nothing may be left behind."""
    import torch
    from realhf_b200.api.model import GenerationHyperparameters
    from realhf_b200.base.topology import ParallelContext
    from realhf_b200.models import generation as gen
    from realhf_b200.models import hf_io
    from realhf_b200.models.real_model import ReaLModel
    from realhf_b200.ops import fp8
    monkeypatch.setenv("REAL_FP8_EMULATE", "1")
    cfg = hf_io.family(fam).make_test_config()
    m = ReaLModel(cfg, ParallelContext.single(), dtype=torch.float32, device=torch.device("cpu")).instantiate(seed=5).eval()
    assert m.fp8_decode_supported()
    calls = [0]
    real = fp8.gemm_fp8

    def counting(*a, **k):
        calls[0] += 1
        return real(*a, **k)
    monkeypatch.setattr(fp8, "gemm_fp8", counting)
    lens = [5, 9, 3]
    ids = torch.randint(3, cfg.vocab_size, (sum(lens),), generator=torch.Generator().manual_seed(1))
    cu = torch.tensor([0, 5, 14, 17], dtype=torch.int32)
    n_new = 6
    g8 = GenerationHyperparameters(max_new_tokens=n_new, min_new_tokens=n_new, greedy=True, fp8_weights=True)
    out8, _ = gen.generate(m, ids, cu, g8, eos_id=None, pad_id=0)
    assert calls[0] == (4 * cfg.n_layers + 1) * (n_new - 1), (calls[0], cfg.n_layers)
    assert m._fp8 is None and not m._fp8_active
    # teacher-forced log-probs of the same tokens under the unquantised model
    diffs = []
    for i, L in enumerate(lens):
        seq = torch.cat([ids[int(cu[i]): int(cu[i + 1])], out8.tokens[i]])
        with torch.no_grad():
            o = m(input_ids=seq, cu_seqlens=torch.tensor([0, seq.numel()], dtype=torch.int32), max_seqlen=int(seq.numel()))
        lp = torch.log_softmax(o.logits.float()[L - 1: L - 1 + n_new], -1)[torch.arange(n_new), out8.tokens[i]]
        diffs.append((out8.logprobs[i] - lp).abs())
    d = torch.stack(diffs)
    assert d[:, 0].max().item() < 1e-4            # the first token comes from the unquantised prefill
    assert 1e-6 < d[:, 1:].mean().item() < 0.3, d  # the rest from the quantised decode: close, and not identical


def test_loss_scale_hysteresis_and_unsupported_post_ln():
    """fp16-style dynamic loss scaling: `hysteresis` overflowing steps are tolerated before the scale halves; the credit refills when the
    scale grows.  `do_layernorm_before=False` is rejected instead of being ignored."""
    import dataclasses as dc

    from realhf_b200.engine.optim import FlatAdamW, OptimizerConfig
    cfg = hf_io.family("llama").make_test_config()
    m = ReaLModel(cfg, dtype=torch.float32).instantiate(seed=1)
    opt = FlatAdamW(m, OptimizerConfig(grad_dtype="fp32", hysteresis=2, loss_scale_window=3, initial_loss_scale=1024.0, min_loss_scale=1.0))
    opt.loss_scale = 1024.0

    def step(overflow):
        opt._skip.fill_(bool(overflow))
        opt._update_loss_scale()
        return opt.loss_scale

    assert step(True) == 1024.0          # first overflow: tolerated
    assert step(True) == 512.0           # second: halve, credit back to 2
    assert step(True) == 512.0
    assert step(False) == 512.0 and step(False) == 512.0 and step(False) == 1024.0   # window of 3 good steps doubles it
    assert step(True) == 1024.0          # the credit was refilled by the growth
    with pytest.raises(NotImplementedError):
        ReaLModel(dc.replace(cfg, do_layernorm_before=False), dtype=torch.float32)
