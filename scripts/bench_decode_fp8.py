"""Decode-step time of LLaMA-7B shapes, bf16 vs the opt-in W8A8 path (graph replay, L2 flushed by the model's own 6.7-13 GB of
weights per step).

    python scripts/bench_decode_fp8.py [layers=32] [ctx=384] [B ...]
"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realhf_b200.api.model import ReaLModelConfig
from realhf_b200.models import generation as gen
from realhf_b200.models.real_model import ReaLModel
from realhf_b200.ops import functional as OF
from realhf_b200.ops import gemm as G

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ctx_len = int(sys.argv[2]) if len(sys.argv) > 2 else 384
Bs = [int(a) for a in sys.argv[3:]] or [16, 64, 128]
OF.set_gemm_impl(G.linear)
dev = torch.device("cuda", 0)
cfg = ReaLModelConfig(n_layers=layers, n_kv_heads=32, n_q_heads=32, hidden_dim=4096, intermediate_dim=11008, vocab_size=32000,
                      n_positions=4096, embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, activation_function="silu",
                      scale_attn_by_inverse_layer_idx=False, use_attention_bias=False, use_attn_proj_bias=False, use_mlp_bias=False,
                      layer_norm_type="rms", mlp_type="llama", apply_rotary=True)
m = ReaLModel(cfg, None, dtype=torch.bfloat16, device=dev).init_random_fast()
for p in m.parameters():
    p.requires_grad_(False)
m.eval()


def time_step(B, fp8):
    st = gen.DecodeState(m, B, 640)
    st.cache_lens.fill_(ctx_len)
    st.input_ids.copy_(torch.arange(B, device=dev) % 1000 + 5)
    gk = torch.Generator(device=dev).manual_seed(11)   # the same synthetic context for both precisions
    for t in list(st.k) + list(st.v):
        t.normal_(0.0, 0.5, generator=gk)
    if fp8:
        m.enable_fp8_decode()

    def step():
        return gen._final_logits(m, m.decode_step(st.input_ids, st.k, st.v, st.cache_lens))
    with torch.no_grad():
        step(); step()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = step()
        for _ in range(5):
            graph.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        n = 50
        for _ in range(n):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    logits = out.float().clone()
    del graph
    if fp8:
        m.disable_fp8_decode()
    return ms, logits


try:
    hbm_gbs = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    hbm_gbs = 6576.4
c = cfg
blk = c.hidden_dim * (c.n_q_heads + 2 * c.n_kv_heads) * c.head_dim + c.n_q_heads * c.head_dim * c.hidden_dim + 3 * c.hidden_dim * c.intermediate_dim
w_elems = layers * blk + c.vocab_size * c.hidden_dim          # every linear weight is read once per step
scale_bytes = 4 * (layers * ((c.n_q_heads + 2 * c.n_kv_heads) * c.head_dim + 2 * c.hidden_dim + 2 * c.intermediate_dim) + c.vocab_size)

for B in Bs:
    kv_bytes = layers * B * ctx_len * 2 * c.n_kv_heads * c.head_dim * 2      # bf16 K and V of the context
    bytes16, bytes8 = 2 * w_elems + kv_bytes, w_elems + scale_bytes + kv_bytes
    t16, l16 = time_step(B, False)
    t8, l8 = time_step(B, True)
    p16, p8 = torch.log_softmax(l16, -1), torch.log_softmax(l8, -1)
    tok = l16.argmax(-1)
    drift = (p16.gather(1, tok[:, None]) - p8.gather(1, tok[:, None])).abs().mean().item()
    print(json.dumps(dict(layers=layers, B=B, ctx=ctx_len, bf16_ms_per_step=round(t16, 4), fp8_ms_per_step=round(t8, 4),
                          speedup=round(t16 / t8, 3), us_per_layer_bf16=round(1e3 * t16 / layers, 1), us_per_layer_fp8=round(1e3 * t8 / layers, 1),
                          bytes_per_step_bf16=bytes16, bytes_per_step_fp8=bytes8, bytes_ratio=round(bytes16 / bytes8, 3),
                          roofline_frac_bf16=round(bytes16 / (t16 * 1e-3) / (hbm_gbs * 1e9), 3),
                          roofline_frac_fp8=round(bytes8 / (t8 * 1e-3) / (hbm_gbs * 1e9), 3), hbm_gbs=hbm_gbs,
                          argmax_agree=round((l8.argmax(-1) == tok).float().mean().item(), 3), mean_abs_logprob_drift=round(drift, 4))))
