"""TP+SP / PP / DP(ZeRO-1) on two real GPUs over NCCL (bf16, tcgen05 GEMMs): losses track the single-GPU run."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("layout", [(1, 1, 2, True), (2, 1, 1, False), (1, 2, 1, False)])
def test_two_gpu_layout_tracks_single_gpu(layout):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import test_parallel_cpu as T
    from realhf_b200.base.testing import run_distributed
    pp, dp, tp, sp = layout
    kw = dict(fam="llama", n_steps=3, device="cuda", dtype=torch.bfloat16, pg_backend="nccl")
    ref = run_distributed(T._worker, 1, backend="nccl", layout=(1, 1, 1, False), n_mbs=dp * (2 * pp if pp > 1 else 1), **kw)[0]
    res = run_distributed(T._worker, 2, backend="nccl", layout=layout, **kw)
    for r in res:
        for a, b in zip(r["losses"], ref["losses"]):
            assert abs(a - b) < 5e-2 * max(1.0, abs(b)), (layout, r["losses"], ref["losses"])
        assert r["losses"][-1] < r["losses"][0]


def _pp_gen_worker(rank, world, use_graph, pp):
    """Greedy generation of a pp-stage pipeline on `world` GPUs, with / without one CUDA graph per (stage, micro-batch)."""
    import types

    from realhf_b200.api.config import ModelName
    from realhf_b200.api.data import SequenceSample
    from realhf_b200.api.model import GenerationHyperparameters, Model, ReaLModelConfig
    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.engine.engine import InferenceBackend
    from realhf_b200.models.real_model import ReaLModel
    from realhf_b200.ops import functional as OF
    from realhf_b200.ops import gemm as G
    from realhf_b200.ops import launches
    OF.set_gemm_impl(G.linear)
    dev = torch.device("cuda", rank)
    cfg = ReaLModelConfig(n_layers=4, n_kv_heads=4, n_q_heads=8, hidden_dim=1024, intermediate_dim=2816, vocab_size=32000, n_positions=2048,
                          embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, activation_function="silu", scale_attn_by_inverse_layer_idx=False,
                          use_attention_bias=False, use_attn_proj_bias=False, use_mlp_bias=False, layer_norm_type="rms", mlp_type="llama",
                          apply_rotary=True)
    ctx = ParallelContext.build(ProcessTopology(pp, 1, 1), list(range(world)), rank, backend="nccl") if world > 1 else ParallelContext.single()
    m = ReaLModel(cfg, ctx, dtype=torch.bfloat16, device=dev).instantiate(seed=5, std=0.05)
    tok = types.SimpleNamespace(eos_token_id=2, pad_token_id=0)
    model = InferenceBackend().initialize(Model(ModelName("m", 0), m, tok, dev), None)
    lens = [9, 33, 17, 64, 5, 12, 40, 21]
    ids = torch.randint(3, 32000, (sum(lens),), generator=torch.Generator().manual_seed(11)).to(dev)
    prompts = SequenceSample.from_default(seqlens=lens, ids=list(range(len(lens))), data=dict(packed_input_ids=ids))
    g = GenerationHyperparameters(max_new_tokens=12, min_new_tokens=12, greedy=True, use_cuda_graph=use_graph, force_cudagraph_recapture=True)
    launches.reset()
    outs = model.module.generate(prompts, tok, g, num_micro_batches=1)
    torch.cuda.synchronize()
    toks = None if outs is None else torch.cat([o.tokens for o in outs]).cpu()
    return dict(tokens=toks, replays=launches.by_op.get("<graph replay>", 0))


def test_pipelined_generation_with_one_cuda_graph_per_microbatch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from realhf_b200.base.testing import run_distributed
    ref = run_distributed(_pp_gen_worker, 1, backend="nccl", use_graph=True, pp=1)[0]
    eager = run_distributed(_pp_gen_worker, 2, backend="nccl", use_graph=False, pp=2)
    graph = run_distributed(_pp_gen_worker, 2, backend="nccl", use_graph=True, pp=2)
    assert all(r["replays"] > 0 for r in graph) and all(r["replays"] == 0 for r in eager)
    t_e = next(r["tokens"] for r in eager if r["tokens"] is not None)
    t_g = next(r["tokens"] for r in graph if r["tokens"] is not None)
    assert torch.equal(t_e, t_g), "graph replay changed the pipeline's tokens"
    # micro-batches are emitted in order; same prompts as the single-GPU run
    assert (t_g == ref["tokens"]).float().mean().item() >= 0.7  # random-init logits are nearly flat: late forks on rounding noise are expected
