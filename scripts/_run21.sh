cd /root/repo
timeout -s KILL 600 python -m pytest tests/test_gemm_gpu.py tests/test_attention_gpu.py tests/test_ops_gpu.py -x -q > gpurun_out/gpu_tests3.log 2>&1; tail -3 gpurun_out/gpu_tests3.log
SMALL_B=1 timeout -s KILL 600 python scripts/bench_decode_gemm.py > gpurun_out/decode_gemm_smallB2.log 2>&1; cut -c1-150 gpurun_out/decode_gemm_smallB2.log
timeout -s KILL 600 python scripts/bench_decode_gemm.py > gpurun_out/decode_gemm8.log 2>&1; cut -c1-150 gpurun_out/decode_gemm8.log
timeout -s KILL 500 python scripts/profile_gen.py 16 > gpurun_out/profile_gen_b16_v2.log 2>&1; grep -A8 "^{" gpurun_out/profile_gen_b16_v2.log | cut -c1-150
