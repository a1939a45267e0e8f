cd /root/repo
SMALL_B=1 REAL_GEMM_DEBUG=1 timeout -s KILL 600 python scripts/bench_decode_gemm.py > gpurun_out/decode_gemm_smallB.log 2>gpurun_out/decode_gemm_smallB.err; cut -c1-330 gpurun_out/decode_gemm_smallB.log; sort gpurun_out/decode_gemm_smallB.err | uniq -c | grep smallm | head -12
SMALL_B=1 timeout -s KILL 600 python scripts/profile_decode.py > gpurun_out/profile_decode_smallB.log 2>&1; tail -6 gpurun_out/profile_decode_smallB.log
SMALL_B=1 timeout -s KILL 600 python scripts/profile_decode.py cublas > gpurun_out/profile_decode_smallB_cublas.log 2>&1; tail -6 gpurun_out/profile_decode_smallB_cublas.log
