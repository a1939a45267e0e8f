"""Mixtral-style MoE block on the GPU: grouped tcgen05 GEMMs (device-side expert offsets) vs the fp32 CPU reference."""
import pytest
import torch

from realhf_b200.models import hf_io
from realhf_b200.models.real_model import ReaLModel
from realhf_b200.ops import functional as OF
from realhf_b200.ops import gemm as G

pytestmark = pytest.mark.gpu


def test_moe_model_forward_backward_matches_cpu_reference():
    cfg = hf_io.family("mixtral").make_test_config()
    lens = [33, 64, 17, 70]
    torch.manual_seed(0)
    ids = torch.randint(2, cfg.vocab_size, (sum(lens),))
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    ref = ReaLModel(cfg, dtype=torch.float32).instantiate(seed=4)
    ref.train()
    out_r = ref(input_ids=ids, cu_seqlens=cu, max_seqlen=max(lens))
    loss_r = out_r.logits.float().square().mean()
    loss_r.backward()
    OF.set_gemm_impl(G.linear)
    try:
        m = ReaLModel(cfg, dtype=torch.bfloat16, device=torch.device("cuda")).instantiate(seed=4)
        m.train()
        out = m(input_ids=ids.cuda(), cu_seqlens=cu.cuda(), max_seqlen=max(lens))
        loss = out.logits.float().square().mean()
        loss.backward()
    finally:
        OF.set_gemm_impl(None)
    torch.testing.assert_close(out.logits.float().cpu(), out_r.logits.float(), atol=0.08, rtol=0.08)
    assert abs(loss.item() - loss_r.item()) < 0.05 * max(1.0, abs(loss_r.item()))
    k = "1.mlp.experts.gate_up.weight"
    torch.testing.assert_close(m.p[k].grad.float().cpu(), ref.p[k].grad, atol=2e-2, rtol=0.2)
    from realhf_b200.ops import launches
    assert launches.by_op.get("gemm_grouped", 0) > 0, "the grouped kernel was not used"


def test_frozen_offload_and_async_reload_roundtrip():
    """Dropping a frozen model's device copy (pinned host copy stays valid) and streaming it back on a side stream."""
    cfg = hf_io.family("llama").make_test_config()
    dev = torch.device("cuda")
    m = ReaLModel(cfg, dtype=torch.bfloat16, device=dev).instantiate(seed=6)
    m.eval()
    ids = torch.randint(2, cfg.vocab_size, (50,), device=dev)
    cu = torch.tensor([0, 20, 50], dtype=torch.int32, device=dev)
    with torch.no_grad():
        ref = m(input_ids=ids, cu_seqlens=cu, max_seqlen=30).logits.clone()
    side = torch.cuda.Stream(dev)
    for _ in range(3):
        m.offload(frozen=True)
        assert not m.instantiated
        m.reload(stream=side)
        torch.cuda.current_stream(dev).wait_stream(side)
        with torch.no_grad():
            out = m(input_ids=ids, cu_seqlens=cu, max_seqlen=30).logits
        assert torch.equal(out, ref)
    m.offload(frozen=True)
    with torch.no_grad():  # implicit reload on first use (what the runtime's OffloadHook relies on)
        assert torch.equal(m(input_ids=ids, cu_seqlens=cu, max_seqlen=30).logits, ref)


def test_train_step_bf16_params_fp32_main_grad():
    """bf16 weights with an fp32 flat gradient buffer: GEMM wgrads accumulate into `main_grad` in place, autograd gradients of
    norms / embeddings are folded in by the post-accumulate hook."""
    import types

    from realhf_b200.api.config import ModelName
    from realhf_b200.api.data import SequenceSample
    from realhf_b200.api.model import FinetuneSpec, Model
    from realhf_b200.engine.engine import TrainBackend
    from realhf_b200.interfaces import basic
    cfg = hf_io.family("llama").make_test_config()
    cfg.n_layers = 2
    dev = torch.device("cuda")
    m = ReaLModel(cfg, dtype=torch.bfloat16, device=dev).instantiate(seed=7)
    tok = types.SimpleNamespace(eos_token_id=1, pad_token_id=0)
    model = TrainBackend(optimizer=dict(lr=1e-2, weight_decay=0.0, warmup_steps_proportion=0.0, lr_scheduler_type="constant",
                                        grad_dtype="fp32", gradient_clipping=1.0)).initialize(Model(ModelName("m", 0), m, tok, dev),
                                                                                               FinetuneSpec(1, 10, 10))
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(5, 14, (8,), generator=g).tolist()
    ids = torch.randint(2, cfg.vocab_size, (sum(lens),), generator=g).to(dev)
    batch = SequenceSample.from_default(seqlens=lens, ids=list(range(8)), data=dict(packed_input_ids=ids, prompt_mask=torch.zeros(sum(lens), dtype=torch.bool, device=dev)))
    itf = basic.SFTInterface()
    losses = [itf.train_step(model, batch, n_mbs=n)["loss"] for n in (1, 2, 4, 1)]
    assert all(l == l for l in losses) and losses[-1] < losses[0], losses


def test_moe_decode_is_cuda_graph_capturable():
    """The MoE decode step must not sync the host (expert counts via scatter_add, device-side offsets for the grouped GEMM): greedy
    generation with the CUDA graph equals generation without it."""
    from realhf_b200.api.model import GenerationHyperparameters
    from realhf_b200.models import generation as gen
    from realhf_b200.models import hf_io
    from realhf_b200.models.real_model import ReaLModel
    from realhf_b200.ops import functional as OF
    from realhf_b200.ops import gemm as G
    OF.set_gemm_impl(G.linear)
    cfg = hf_io.family("mixtral").make_test_config()
    cfg.hidden_dim, cfg.intermediate_dim, cfg.n_q_heads, cfg.n_kv_heads, cfg.head_dim, cfg.vocab_size, cfg.n_layers = 512, 1024, 4, 4, 128, 1024, 2
    cfg.moe.num_experts, cfg.moe.top_k = 8, 2
    m = ReaLModel(cfg, dtype=torch.bfloat16, device=torch.device("cuda")).instantiate(seed=3, std=0.05)
    for p in m.parameters():
        p.requires_grad_(False)
    m.eval()
    lens = [7, 19, 4, 11]
    ids = torch.randint(3, 1024, (sum(lens),), device="cuda")
    cu = torch.tensor([0, 7, 26, 30, 41], dtype=torch.int32, device="cuda")
    outs = []
    for graph in (True, False):
        g = GenerationHyperparameters(max_new_tokens=10, min_new_tokens=10, greedy=True, use_cuda_graph=graph, force_cudagraph_recapture=True)
        o, _ = gen.generate(m, ids, cu, g, eos_id=2, pad_id=0)
        outs.append(o.tokens)
    assert torch.equal(outs[0], outs[1])
