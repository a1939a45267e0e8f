"""End-to-end generation MFC profile (LLaMA-7B, B prompts of 128 tokens, 512 new tokens): wall vs kernel time, top kernels."""
import json, os, sys, time, types
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realhf_b200.api.config import ModelName
from realhf_b200.api.data import SequenceSample
from realhf_b200.api.model import Model, ReaLModelConfig
from realhf_b200.base.topology import ParallelContext
from realhf_b200.engine.engine import InferenceBackend
from realhf_b200.interfaces import ppo
from realhf_b200.models.real_model import ReaLModel
from realhf_b200.ops import functional as OF
from realhf_b200.ops import gemm as G

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
new = int(sys.argv[2]) if len(sys.argv) > 2 else 512
OF.set_gemm_impl(G.linear)
dev = torch.device("cuda")
cfg = ReaLModelConfig(n_layers=32, n_kv_heads=32, n_q_heads=32, hidden_dim=4096, intermediate_dim=11008, vocab_size=32000,
                      n_positions=4096, embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, activation_function="silu",
                      scale_attn_by_inverse_layer_idx=False, use_attention_bias=False, use_attn_proj_bias=False, use_mlp_bias=False,
                      layer_norm_type="rms", mlp_type="llama", apply_rotary=True)
m = ReaLModel(cfg, ParallelContext.single(), dtype=torch.bfloat16, device=dev).init_random_fast()
tok = types.SimpleNamespace(eos_token_id=2, pad_token_id=0)
model = InferenceBackend().initialize(Model(ModelName("actor", 0), m, tok, dev), None)
gcfg = dict(max_new_tokens=new, min_new_tokens=new, greedy=False, top_p=0.9, top_k=1000, temperature=1.0, use_cuda_graph=True,
            force_cudagraph_recapture=True)
itf = ppo.PPOActorInterface(generation_config=gcfg)
batch = SequenceSample.from_default(seqlens=[128] * B, ids=list(range(B)), data=dict(packed_prompts=torch.randint(3, 32000, (128 * B,), device=dev)))
for _ in range(2):
    itf.generate(model, batch, n_mbs=1)
torch.cuda.synchronize()
t0 = time.perf_counter(); itf.generate(model, batch, n_mbs=1); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    itf.generate(model, batch, n_mbs=1)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
    if t > 0:
        rows.append((t, e.count, e.key[:100]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(json.dumps(dict(B=B, new_tokens=new, wall_ms=round(wall, 1), kernel_ms=round(tot / 1e3, 1), ms_per_token_wall=round(wall / new, 3))))
for t, c, k in rows[:22]:
    print(f"{t/1e3:9.2f} ms  {100*t/tot:5.1f}%  x{c:<6d} {k}")
