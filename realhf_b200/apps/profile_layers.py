"""`python -m realhf_b200.apps.profile_layers --family llama --size 7 --bs 1 8 --seqlen 128 1024` (parity: apps/profile_layers.py)."""

import argparse
import json

import torch


def main(argv=None):
    from realhf_b200.api.quickstart import ModelTrainEvalConfig
    from realhf_b200.api.config import ModelFamily
    from realhf_b200.api.model import ReaLModelConfig
    from realhf_b200.search.engine import model_shape
    from realhf_b200.search.layers import dump_profile, profile_decode, profile_head, profile_layers, profile_optimizer
    ap = argparse.ArgumentParser()
    ap.add_argument("--family", default="llama")
    ap.add_argument("--size", type=int, default=7)
    ap.add_argument("--bs", type=int, nargs="+", default=[1, 8])
    ap.add_argument("--seqlen", type=int, nargs="+", default=[256, 1024])
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    ap.add_argument("--decode-bs", type=int, nargs="*", default=[16, 64, 128], help="sequences per decode step (empty: skip decode rows)")
    ap.add_argument("--decode-ctx", type=int, nargs="*", default=[384])
    ap.add_argument("--no-head", action="store_true")
    ap.add_argument("--no-optimizer", action="store_true")
    a = ap.parse_args(argv)
    sh = model_shape(ModelTrainEvalConfig(type=ModelFamily(a.family, a.size, False)))
    cfg = ReaLModelConfig(n_layers=2, n_kv_heads=int(sh["h"]) // 128, n_q_heads=int(sh["h"]) // 128, hidden_dim=int(sh["h"]),
                          intermediate_dim=int(sh["f"]), vocab_size=int(sh["v"]), n_positions=4096, embd_pdrop=0.0, resid_pdrop=0.0,
                          attn_pdrop=0.0, activation_function="silu", scale_attn_by_inverse_layer_idx=False, use_attention_bias=False,
                          use_attn_proj_bias=False, use_mlp_bias=False, layer_norm_type="rms", mlp_type="llama", apply_rotary=True)
    dt = torch.bfloat16 if a.device == "cuda" else torch.float32
    rows = profile_layers(cfg, a.bs, a.seqlen, device=a.device, dtype=dt)
    if not a.no_head:
        rows += profile_head(cfg, a.bs, a.seqlen, device=a.device, dtype=dt)
    if a.decode_bs and a.decode_ctx:
        rows += profile_decode(cfg, a.decode_bs, a.decode_ctx, device=a.device, dtype=dt)
    if not a.no_optimizer:
        rows += profile_optimizer(cfg, device=a.device, dtype=dt)
    print(json.dumps(rows, indent=1))
    print("written to", dump_profile(rows, f"{a.family}-{a.size}"))


if __name__ == "__main__":
    main()
