"""Per-layer ZeRO-3 on the GPU: parameter shard parked in pinned host memory, layer-by-layer H2D gather with prefetch on a side
stream, recompute-in-backward, gradient shard + AdamW on the shard — against the plain (ZeRO-1, resident) run."""
import os
import sys
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _run(zero_stage, offload_param, steps=3, offload_optimizer=False):
    from realhf_b200.api.config import ModelName
    from realhf_b200.api.data import SequenceSample
    from realhf_b200.api.model import FinetuneSpec, Model, ReaLModelConfig
    from realhf_b200.engine.engine import TrainBackend
    from realhf_b200.interfaces import basic
    from realhf_b200.models.real_model import ReaLModel
    from realhf_b200.ops import functional as OF
    from realhf_b200.ops import gemm as G
    OF.set_gemm_impl(G.linear)
    dev = torch.device("cuda", 0)
    cfg = ReaLModelConfig(n_layers=6, n_kv_heads=4, n_q_heads=8, hidden_dim=1024, intermediate_dim=2816, vocab_size=4096, n_positions=2048,
                          embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, activation_function="silu", scale_attn_by_inverse_layer_idx=False,
                          use_attention_bias=False, use_attn_proj_bias=False, use_mlp_bias=False, layer_norm_type="rms", mlp_type="llama",
                          apply_rotary=True)
    m = ReaLModel(cfg, dtype=torch.bfloat16, device=dev).instantiate(seed=3, std=0.03)
    tok = types.SimpleNamespace(eos_token_id=1, pad_token_id=0)
    model = TrainBackend(optimizer=dict(lr=2e-4, weight_decay=0.0, warmup_steps_proportion=0.0, lr_scheduler_type="constant", grad_dtype="fp32",
                                        gradient_clipping=1.0), zero_stage=zero_stage, offload_param=offload_param, offload_optimizer=offload_optimizer).initialize(
        Model(ModelName("m", 0), m, tok, dev), FinetuneSpec(1, 10, 10))
    g = torch.Generator().manual_seed(0)
    lens = [200, 333, 64, 512, 90, 41]
    ids = torch.randint(2, 4096, (sum(lens),), generator=g).to(dev)
    batch = SequenceSample.from_default(seqlens=lens, ids=list(range(len(lens))),
                                        data=dict(packed_input_ids=ids, prompt_mask=torch.zeros(sum(lens), dtype=torch.bool, device=dev)))
    torch.cuda.reset_peak_memory_stats()
    losses = [float(basic.SFTInterface().train_step(model, batch, n_mbs=2)["loss"]) for _ in range(steps)]
    opt = model.module.optim
    return dict(losses=losses, z3=opt.z3, peak=torch.cuda.max_memory_allocated(), grad_norm=float(opt.last_grad_norm))


@pytest.mark.parametrize("offload_optimizer", [False, True])
def test_per_layer_zero3_with_param_offload_tracks_resident_training(offload_optimizer):
    ref = _run(1, False)
    z = _run(3, True, offload_optimizer=offload_optimizer)
    assert z["z3"] is not None and z["z3"].n_gathers > 0 and not z["z3"].pshard.is_cuda and z["z3"].pshard.is_pinned()
    for a, b in zip(ref["losses"], z["losses"]):
        assert abs(a - b) < 3e-2 * max(1.0, abs(b)), (ref["losses"], z["losses"])
    assert z["losses"][-1] < z["losses"][0]
    assert abs(ref["grad_norm"] - z["grad_norm"]) < 0.1 * max(1.0, ref["grad_norm"])
