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
