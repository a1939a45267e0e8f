#!/bin/bash
# First GPU call of a session: run everything that was written without hardware access, each step under its own timeout so a
# hang in one kernel cannot take the box down.  Usage (from the repo root):
#     gpurun --timeout 900 -- 'bash scripts/validate_experimental.sh > gpurun_out/experimental.log 2>&1'
set -u
export REAL_TEST_EXPERIMENTAL=1
export PYTHONPATH="$(pwd):${PYTHONPATH:-}"
mkdir -p gpurun_out
echo "== attention forward (tcgen05) numerics"
timeout 240 python -m pytest tests/test_attention_gpu.py -x -q -k attn_fwd_tcgen05 2>&1 | tail -15
echo "== attention backward (tcgen05) numerics"
timeout 240 python -m pytest tests/test_attention_gpu.py -x -q -k attn_bwd_tcgen05 2>&1 | tail -15
echo "== grouped wgrad (MoE) numerics"
timeout 180 python -m pytest tests/test_gemm_gpu.py -x -q -k grouped_wgrad_single_launch 2>&1 | tail -8
echo "== LayerNorm kernel numerics"
timeout 120 python -m pytest tests/test_ops_gpu.py -x -q -k layernorm_native 2>&1 | tail -8
echo "== attention forward vs flash-attn (CUDA events)"
timeout 120 python scripts/bench_attn.py --seqs 32 --len 640 2>&1 | tail -4
timeout 120 python scripts/bench_attn.py --seqs 8 --len 4096 2>&1 | tail -4
echo "== model-level: one SFT step with both kernels switched on vs the library path"
REAL_ATTN=tcgen05 REAL_ATTN_BWD=tcgen05 timeout 300 python - <<'PY' 2>&1 | tail -6
import os, torch
from realhf_b200.models import hf_io
from realhf_b200.models.real_model import ReaLModel
from realhf_b200.ops import attention as A
cfg = hf_io.family("llama").make_test_config()
cfg.n_layers, cfg.hidden_dim, cfg.n_q_heads, cfg.n_kv_heads, cfg.head_dim, cfg.intermediate_dim = 2, 1024, 8, 8, 128, 2048
torch.manual_seed(0)
lens = [300, 640, 77]
dev = "cuda" if torch.cuda.is_available() else "cpu"
ids = torch.randint(2, cfg.vocab_size, (sum(lens),), device=dev)
cu = torch.tensor([0, 300, 940, 1017], dtype=torch.int32, device=dev)
grads = {}
combos = [("flash", "flash"), ("tcgen05", "flash"), ("flash", "tcgen05"), ("tcgen05", "tcgen05")]   # isolates fwd from bwd problems
for impl in combos:
    os.environ["REAL_ATTN"], os.environ["REAL_ATTN_BWD"] = impl
    m = ReaLModel(cfg, dtype=torch.bfloat16 if dev == "cuda" else torch.float32, device=dev).instantiate(seed=3)
    out = m(input_ids=ids, cu_seqlens=cu, max_seqlen=640)
    loss = out.logits.float().logsumexp(-1).mean()
    params = [p for p in m.p.values() if p.requires_grad]
    gs = torch.autograd.grad(loss, params, allow_unused=True)
    grads[impl] = (loss.item(), torch.cat([g.float().flatten() for g in gs if g is not None]))
l0, g0 = grads[combos[0]]
for c in combos[1:]:
    l1, g1 = grads[c]
    print("fwd/bwd =", c, "loss", l0, l1, "grad rel err", ((g0 - g1).norm() / g0.norm()).item())
PY
echo "== done"
