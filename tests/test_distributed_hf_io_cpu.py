"""Sharded HuggingFace checkpoint I/O (parity: tests/model/test_distributed_load_hf.py): load a checkpoint under one
(pp, dp, tp) layout, save it from there, reload under another layout, save again -- the tensors must survive bit-exactly,
`transformers` must load the result, critic heads and init-critic-from-actor included (4 CPU processes, gloo)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.distributed
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, fam, src_dir, out_a, out_b, layout_a, layout_b, is_critic, init_critic_from_actor):
    import torch.distributed as dist

    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    from realhf_b200.models import hf_io
    from realhf_b200.models.real_model import ReaLModel
    for layout, src, dst in ((layout_a, src_dir, out_a), (layout_b, out_a, out_b)):
        ctx = ParallelContext.build(ProcessTopology(*layout), list(range(world)), rank, backend="gloo")
        from_actor = init_critic_from_actor and src == src_dir
        cfg = hf_io.config_from_hf_path(fam, src, is_critic)
        m = ReaLModel(cfg, ctx, dtype=torch.float32)
        hf_io.load_from_hf(m, fam, src, init_critic_from_actor=from_actor)
        hf_io.save_to_hf(m, fam, dst)
        dist.barrier()
    return True


@pytest.mark.parametrize("case", [("llama", (2, 1, 2), (1, 2, 2), False, False), ("gpt2", (4, 1, 1), (1, 1, 4), False, False),
                                  ("llama", (1, 2, 2), (2, 2, 1), True, False), ("qwen2", (2, 2, 1), (1, 1, 4), True, True),
                                  ("mixtral", (1, 2, 2), (2, 1, 2), False, False)])
def test_sharded_save_load_roundtrip(tmp_path, case):
    import fixtures
    from realhf_b200.base.testing import run_distributed
    from realhf_b200.models import hf_io
    fam, la, lb, is_critic, from_actor = case
    src = str(tmp_path / "src")
    fixtures.make_checkpoint(src, fam, is_critic=is_critic and not from_actor, seed=3)
    out_a, out_b = str(tmp_path / "a"), str(tmp_path / "b")
    assert all(run_distributed(_worker, 4, backend="gloo", timeout=600, fam=fam, src_dir=src, out_a=out_a, out_b=out_b, layout_a=la,
                               layout_b=lb, is_critic=is_critic, init_critic_from_actor=from_actor))
    sd_src, sd_a, sd_b = (hf_io.load_hf_state_dict(d) for d in (src, out_a, out_b))
    assert set(sd_a) == set(sd_b)
    for k in sd_a:
        assert torch.equal(sd_a[k], sd_b[k]), k                       # layout A -> layout B is lossless
        if k in sd_src and not (from_actor and "lm_head" in k) and sd_src[k].shape == sd_a[k].shape:
            assert torch.equal(sd_a[k], sd_src[k]), k                 # and equals what was loaded
    nbytes = lambda sd: sum(v.numel() * v.element_size() for v in sd.values())  # (file count and headers depend on pp)
    assert nbytes(sd_a) == nbytes(sd_b)
    if not is_critic:
        import transformers
        transformers.AutoModelForCausalLM.from_pretrained(out_b)
