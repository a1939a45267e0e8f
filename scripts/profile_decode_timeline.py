"""Per-kernel timeline of ONE graph-replayed decode step (LLaMA-7B shapes, few layers): start, duration and the idle gap
before every kernel, from CUPTI timestamps (torch.profiler).  Shows where the step's time above the byte roofline goes.

    python scripts/profile_decode_timeline.py [layers=4] [B=16] [ctx=384]
"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realhf_b200.api.model import GenerationHyperparameters, ReaLModelConfig
from realhf_b200.models import generation as gen
from realhf_b200.models.real_model import ReaLModel
from realhf_b200.ops import functional as OF
from realhf_b200.ops import gemm as G

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ctx_len = int(sys.argv[3]) if len(sys.argv) > 3 else 384
OF.set_gemm_impl(G.linear)
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
ctx = None
if world > 1:  # tensor-parallel decode (launch with torchrun): B is the batch of the whole TP group
    import torch.distributed as dist
    from realhf_b200.base.topology import ParallelContext, ProcessTopology
    torch.cuda.set_device(rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    ctx = ParallelContext.build(ProcessTopology(1, 1, world), list(range(world)), rank, backend="nccl")
    from realhf_b200.parallel.fused_tp import FusedTP
    ctx.symm = FusedTP(ctx, max_tokens=256, max_features=4096, device=torch.device("cuda", rank))
dev = torch.device("cuda", rank)
cfg = ReaLModelConfig(n_layers=layers, n_kv_heads=32, n_q_heads=32, hidden_dim=4096, intermediate_dim=11008, vocab_size=32000,
                      n_positions=4096, embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, activation_function="silu",
                      scale_attn_by_inverse_layer_idx=False, use_attention_bias=False, use_attn_proj_bias=False, use_mlp_bias=False,
                      layer_norm_type="rms", mlp_type="llama", apply_rotary=True)
m = ReaLModel(cfg, ctx, dtype=torch.bfloat16, device=dev).init_random_fast()
for p in m.parameters():
    p.requires_grad_(False)
m.eval()
st = gen.DecodeState(m, B, 640)
st.cache_lens.fill_(ctx_len)
st.input_ids.fill_(5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def step():
    h = m.decode_step(st.input_ids, st.k, st.v, st.cache_lens)
    return gen._final_logits(m, h)


with torch.no_grad():
    step(); step()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = step()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            flush.zero_()  # weights of a 4-layer model would otherwise sit in the 126 MB L2
            graph.replay()
        torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memset" not in e.name.lower()
       and "Memset" not in e.name]
evs.sort(key=lambda e: e.time_range.start)
# keep the last replay: everything after the last fill kernel
last_fill = max(i for i, e in enumerate(evs) if "fill" in e.name.lower() or "vectorized_elementwise" in e.name)
evs = evs[last_fill + 1:]
t0 = evs[0].time_range.start
prev_end = t0
rows = []
for e in evs:
    s, d = e.time_range.start, e.time_range.end - e.time_range.start
    rows.append(dict(name=e.name[:70], start_us=round(s - t0, 1), dur_us=round(d, 1), gap_us=round(s - prev_end, 1)))
    prev_end = max(prev_end, e.time_range.end)
total = prev_end - t0
busy = sum(r["dur_us"] for r in rows)
if rank == 0:
    print(json.dumps(dict(layers=layers, B=B, ctx=ctx_len, tp=world, total_us=round(total, 1), sum_kernel_us=round(busy, 1), n_kernels=len(rows))))
    for r in rows:
        print(json.dumps(r))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
