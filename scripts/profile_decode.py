"""Decode-path breakdown on one GPU: graph-replayed model step vs sampling step, LLaMA-7B shapes."""
import json, os, sys, types
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realhf_b200.api.model import GenerationHyperparameters, ReaLModelConfig
from realhf_b200.models import generation as gen
from realhf_b200.models.real_model import ReaLModel
from realhf_b200.ops import functional as OF
from realhf_b200.ops import gemm as G

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
impl = sys.argv[2] if len(sys.argv) > 2 else "tcgen05"
OF.set_gemm_impl(G.linear if impl == "tcgen05" else False)
dev = torch.device("cuda")
cfg = ReaLModelConfig(n_layers=layers, n_kv_heads=32, n_q_heads=32, hidden_dim=4096, intermediate_dim=11008, vocab_size=32000,
                      n_positions=4096, embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, activation_function="silu",
                      scale_attn_by_inverse_layer_idx=False, use_attention_bias=False, use_attn_proj_bias=False, use_mlp_bias=False,
                      layer_norm_type="rms", mlp_type="llama", apply_rotary=True)
m = ReaLModel(cfg, dtype=torch.bfloat16, device=dev).init_random_fast()
for p in m.parameters():
    p.requires_grad_(False)
m.eval()
g = GenerationHyperparameters(max_new_tokens=512, min_new_tokens=512, top_p=0.9, top_k=1000, use_cuda_graph=True)

def timeit(f, n=20, warm=3):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

for B in ((16, 32) if os.environ.get("SMALL_B") else (64, 128)):
    st = gen.DecodeState(m, B, 640)
    for ctx in (128, 384, 639):
        st.cache_lens.fill_(ctx)
        st.input_ids.fill_(5)
        def step():
            h = m.decode_step(st.input_ids, st.k, st.v, st.cache_lens)
            return gen._final_logits(m, h)
        with torch.no_grad():
            t_eager = timeit(step, n=5, warm=2)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = step()
            t_graph = timeit(graph.replay, n=20)
        logits = out
        unf = torch.ones(B, dtype=torch.bool, device=dev)
        t_samp = timeit(lambda: gen.genstep(logits, g, 5, 2, 0, unf), n=20)
        wbytes = m.flat_numel * 2
        kvbytes = 2 * B * 32 * (ctx + 1) * 128 * 2 * layers
        print(json.dumps(dict(B=B, ctx=ctx, layers=layers, gemm=impl, eager_ms=round(t_eager, 3), graph_ms=round(t_graph, 3),
                              sample_ms=round(t_samp, 3), roofline_ms=round((wbytes + kvbytes) / 6.58e9, 3))), flush=True)
        del graph
    del st
