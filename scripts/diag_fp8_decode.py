"""Per-call comparison of the W8A8 decode path with its PyTorch emulation on a tiny LLaMA (one decode step): for every quantiser
call the fraction of identical bytes / scales, for every GEMM call the relative error of the kernel against fp32 math on the
SAME quantised operands.  Prints one JSON line per call and a summary."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_sampling_gpu import _tiny_llama
from realhf_b200.models import generation as gen
from realhf_b200.ops import fp8
from realhf_b200.ops import functional as OF

m = _tiny_llama()
B = 4
st = gen.DecodeState(m, B, 64)
g = torch.Generator(device="cuda").manual_seed(5)
for t in list(st.k) + list(st.v):
    t.normal_(0.0, 0.5, generator=g)
ids = torch.tensor([5, 17, 300, 31999], device="cuda")


def step():
    st.cache_lens.fill_(20)
    with torch.no_grad():
        return gen._final_logits(m, m.decode_step(ids, st.k, st.v, st.cache_lens)).float()


bf16 = step()
m.enable_fp8_decode()
real = dict(gemm=fp8.gemm_fp8, addnorm=fp8.add_rmsnorm_quant, gated=fp8.gated_act_quant, qrows=fp8.quantize_rows)
log = []


def gemm(qx, sx, qw, sw, bias=None, out_dtype=torch.bfloat16, out=None, bn=0, split=0):
    y = real["gemm"](qx, sx, qw, sw, bias, out_dtype, out, bn, split)
    ref = fp8.dequantize(qx, sx[: qx.shape[0]]) @ fp8.dequantize(qw, sw).t()
    log.append(dict(op="gemm", M=qx.shape[0], N=qw.shape[0], K=qx.shape[1], rel=round(((y.float() - ref).norm() / ref.norm()).item(), 5),
                    max_abs_over_max=round(((y.float() - ref).abs().max() / ref.abs().max()).item(), 5)))
    return y


def cmp_q(name, q, s, h):
    qr, sr = fp8.quantize_rows_ref(h)
    log.append(dict(op=name, shape=list(q.shape), byte_eq=round((q == qr).float().mean().item(), 5), scale_eq=bool(torch.equal(s, sr)),
                    scale_rel=float(((s - sr).abs() / sr).max())))


def addnorm(d, x, w, eps, w_offset=0.0):
    r = real["addnorm"](d, x, w, eps, w_offset)
    h = OF.rmsnorm(x, w, eps, w_offset) if d is None else OF.add_rmsnorm(d, x, w, eps, w_offset)[0]
    cmp_q("add_rmsnorm_quant", r[0], r[1], h)
    return r


def gated(gu, kind):
    r = real["gated"](gu, kind)
    cmp_q("gated_act_quant", r[0], r[1], OF.gated_act(gu, kind))
    return r


def qrows(x, q_out=None, scale_out=None):
    r = real["qrows"](x, q_out, scale_out)
    cmp_q("quantize_rows", r[0], r[1], x)
    return r


fp8.gemm_fp8, fp8.add_rmsnorm_quant, fp8.gated_act_quant, fp8.quantize_rows = gemm, addnorm, gated, qrows
kern = step()
for l in log:
    print(json.dumps(l))


def gemm_ref(qx, sx, qw, sw, bias=None, out_dtype=torch.bfloat16, out=None, bn=0, split=0):
    return (fp8.dequantize(qx, sx[: qx.shape[0]]) @ fp8.dequantize(qw, sw).t()).to(out_dtype)


def addnorm_ref(d, x, w, eps, w_offset=0.0):
    h, r = (OF.rmsnorm(x, w, eps, w_offset), x) if d is None else OF.add_rmsnorm(d, x, w, eps, w_offset)
    q, s = fp8.quantize_rows_ref(h)
    return q, s, r


fp8.gemm_fp8, fp8.add_rmsnorm_quant = gemm_ref, addnorm_ref
fp8.gated_act_quant = lambda gu, kind: fp8.quantize_rows_ref(OF.gated_act(gu, kind))
fp8.quantize_rows = lambda x, q_out=None, scale_out=None: fp8.quantize_rows_ref(x)
emu = step()
emu2 = step()
fp8.gemm_fp8, fp8.add_rmsnorm_quant, fp8.gated_act_quant, fp8.quantize_rows = real["gemm"], real["addnorm"], real["gated"], real["qrows"]
kern2 = step()
print(json.dumps(dict(summary=True, rel_kernel_vs_emulation=((kern - emu).norm() / emu.norm()).item(),
                      rel_kernel_vs_kernel_again=((kern - kern2).norm() / kern.norm()).item(),
                      rel_emulation_vs_emulation_again=((emu - emu2).norm() / emu.norm()).item(),
                      rel_kernel_vs_bf16=((kern - bf16).norm() / bf16.norm()).item(),
                      rel_emulation_vs_bf16=((emu - bf16).norm() / bf16.norm()).item())))
