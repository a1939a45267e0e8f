"""Kernel-time breakdown of one training micro-batch (LLaMA-7B shapes, few layers) with torch.profiler."""
import json, os, sys, types
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realhf_b200.api.config import ModelName
from realhf_b200.api.data import SequenceSample
from realhf_b200.api.model import FinetuneSpec, Model, ReaLModelConfig
from realhf_b200.base.topology import ParallelContext
from realhf_b200.engine.engine import TrainBackend
from realhf_b200.interfaces import basic
from realhf_b200.models.real_model import ReaLModel
from realhf_b200.ops import functional as OF
from realhf_b200.ops import gemm as G

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 4
impl = sys.argv[2] if len(sys.argv) > 2 else "tcgen05"
OF.set_gemm_impl(G.linear if impl == "tcgen05" else False)
dev = torch.device("cuda")
ctx = ParallelContext.single(); ctx.gradient_checkpointing = True
cfg = ReaLModelConfig(n_layers=layers, n_kv_heads=32, n_q_heads=32, hidden_dim=4096, intermediate_dim=11008, vocab_size=32000,
                      n_positions=4096, embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, activation_function="silu",
                      scale_attn_by_inverse_layer_idx=False, use_attention_bias=False, use_attn_proj_bias=False, use_mlp_bias=False,
                      layer_norm_type="rms", mlp_type="llama", apply_rotary=True)
m = ReaLModel(cfg, ctx, dtype=torch.bfloat16, device=dev).init_random_fast()
tok = types.SimpleNamespace(eos_token_id=2, pad_token_id=0)
model = TrainBackend(optimizer=dict(lr=1e-5, state_dtype="bf16", use_master_weights=False)).initialize(Model(ModelName("a", 0), m, tok, dev), FinetuneSpec(1, 10, 10))
lens = [640] * 32
ids = torch.randint(3, 32000, (sum(lens),), device=dev)
batch = SequenceSample.from_default(seqlens=lens, ids=list(range(32)), data=dict(packed_input_ids=ids, prompt_mask=torch.zeros(sum(lens), dtype=torch.bool, device=dev)))
itf = basic.SFTInterface()
for _ in range(2):
    itf.train_step(model, batch, n_mbs=1)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    itf.train_step(model, batch, n_mbs=1)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
    if t > 0 and e.device_type is not None and "cuda" in str(e.device_type).lower():
        rows.append((t, e.count, e.key[:110]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(json.dumps(dict(layers=layers, gemm=impl, total_kernel_ms=round(tot / 1e3, 2))))
for t, c, k in rows[:28]:
    print(f"{t/1e3:9.2f} ms  {100*t/tot:5.1f}%  x{c:<5d} {k}")
