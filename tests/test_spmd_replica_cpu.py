"""SPMD executor with a per-MFC layout replica: actor trains on dp2, generation runs on a tp2 replica refreshed by
parameter reallocation; inputs are regrouped over the TP group and outputs sliced back (gloo, world_size 2)."""
import types

import torch


def _worker(rank, world):
    import torch.distributed as dist

    from realhf_b200.api.config import ModelInterfaceAbstraction, ModelInterfaceType, ModelName
    from realhf_b200.api.data import SequenceSample
    from realhf_b200.api.dfg import MFCDef
    from realhf_b200.api.model import FinetuneSpec, Model
    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.engine.engine import TrainBackend
    from realhf_b200.interfaces import basic, ppo  # noqa: F401
    from realhf_b200.models import hf_io
    from realhf_b200.models.real_model import ReaLModel
    from realhf_b200.system.spmd import SPMDExecutor

    tok = types.SimpleNamespace(eos_token_id=1, pad_token_id=0)
    cfg = hf_io.family("llama").make_test_config()
    src_topo, dst_topo = ProcessTopology(1, world, 1), ProcessTopology(1, 1, world)
    ctx_train = ParallelContext.build(src_topo, list(range(world)), rank, backend="gloo")
    ctx_gen = ParallelContext.build(dst_topo, list(range(world)), rank, backend="gloo")
    m = ReaLModel(cfg, ctx_train, dtype=torch.float32).instantiate(seed=5)
    actor = TrainBackend(optimizer=dict(lr=1e-3, weight_decay=0.0, warmup_steps_proportion=0.0, lr_scheduler_type="constant",
                                        grad_dtype="fp32")).initialize(Model(ModelName("actor", 0), m, tok, "cpu"), FinetuneSpec(1, 10, 10))
    gcfg = dict(max_new_tokens=6, min_new_tokens=6, greedy=True)
    itf = ppo.PPOActorInterface(n_minibatches=1, generation_config=gcfg)
    A = lambda t, **a: ModelInterfaceAbstraction(t, a)
    rpcs = [MFCDef("actor_gen", 4, ModelInterfaceType.GENERATE, A("ppo_actor"), "actor", input_keys=("packed_prompts",),
                   output_keys=("seq_no_eos_mask", "packed_input_ids", "packed_logprobs", "prompt_mask", "packed_logits_mask"), n_mbs=1)]
    ex = SPMDExecutor(rpcs, {"actor": actor}, {"actor_gen": itf}, "cpu")
    ex.add_layout_replica("actor_gen", actor, ctx_gen, src_topo, dst_topo, list(range(world)), rank)
    plens = [[5, 7], [4, 6]][rank]
    g = torch.Generator().manual_seed(100 + rank)
    prompts = torch.randint(2, 128, (sum(plens),), generator=g)
    batch = SequenceSample.from_default(seqlens=plens, ids=[f"r{rank}p{j}" for j in range(2)], data=dict(packed_prompts=prompts))
    rec = ex.run_step(batch)
    out = ex.last_pool
    assert out.ids == [f"r{rank}p{j}" for j in range(2)]
    # reference: the un-sharded training replica generates the same sequences for my prompts (greedy)
    ref = itf.generate(actor, SequenceSample.from_default(seqlens=plens, ids=[0, 1], data=dict(packed_prompts=prompts)), n_mbs=1)
    assert out.flat_seqlens("packed_input_ids") == ref.flat_seqlens("packed_input_ids")
    assert torch.equal(out.data["packed_input_ids"], ref.data["packed_input_ids"])
    torch.testing.assert_close(out.data["packed_logprobs"], ref.data["packed_logprobs"], atol=1e-4, rtol=1e-3)
    # weights change -> the hook must refresh the replica before the next generation
    with torch.no_grad():
        m.flat_param.data.mul_(1.5)
    ex.run_step(batch)
    ref2 = itf.generate(actor, SequenceSample.from_default(seqlens=plens, ids=[0, 1], data=dict(packed_prompts=prompts)), n_mbs=1)
    assert torch.equal(ex.last_pool.data["packed_input_ids"], ref2.data["packed_input_ids"])
    dist.barrier()
    return True


def test_spmd_tp_generation_replica_matches_dp_layout():
    from realhf_b200.base.testing import run_distributed
    assert all(run_distributed(_worker, 2, backend="gloo", timeout=300))
