cd /root/repo
REAL_PDL=1 timeout -s KILL 600 python -m pytest tests/test_gemm_gpu.py tests/test_attention_gpu.py tests/test_ops_gpu.py -x -q > gpurun_out/gpu_tests5.log 2>&1; tail -2 gpurun_out/gpu_tests5.log
timeout -s KILL 500 python scripts/profile_gen.py 16 > gpurun_out/profile_gen_b16_v4.log 2>&1; grep "^{" gpurun_out/profile_gen_b16_v4.log | cut -c1-150
timeout -s KILL 500 python scripts/profile_gen.py 32 > gpurun_out/profile_gen_b32_v4.log 2>&1; grep "^{" gpurun_out/profile_gen_b32_v4.log | cut -c1-150
REAL_PDL=0 timeout -s KILL 500 python scripts/profile_gen.py 32 > gpurun_out/profile_gen_b32_nopdl.log 2>&1; grep "^{" gpurun_out/profile_gen_b32_nopdl.log | cut -c1-150
REAL_PDL=1 timeout -s KILL 500 python scripts/profile_gen.py 64 > gpurun_out/profile_gen_b64_pdl.log 2>&1; grep "^{" gpurun_out/profile_gen_b64_pdl.log | cut -c1-150
REAL_PDL=0 timeout -s KILL 500 python scripts/profile_gen.py 64 > gpurun_out/profile_gen_b64_nopdl.log 2>&1; grep "^{" gpurun_out/profile_gen_b64_nopdl.log | cut -c1-150
