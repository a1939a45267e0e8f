cd /root/repo
REAL_GEMM_DEBUG=1 timeout -s KILL 300 python - <<'PY' 2>&1 | sort | uniq -c | head -20
import torch
from realhf_b200.ops import gemm as G
for M in (64, 128):
    for (N, K) in [(12288, 4096), (4096, 4096), (22016, 4096), (4096, 11008), (32000, 4096)]:
        G.gemm(torch.randn(M, K, device="cuda", dtype=torch.bfloat16), torch.randn(N, K, device="cuda", dtype=torch.bfloat16))
torch.cuda.synchronize()
PY
timeout -s KILL 600 python scripts/bench_decode_gemm.py > gpurun_out/decode_gemm7.log 2>&1; cut -c1-150 gpurun_out/decode_gemm7.log
timeout -s KILL 900 python scripts/profile_decode.py > gpurun_out/profile_decode_tc3.log 2>&1; tail -8 gpurun_out/profile_decode_tc3.log
