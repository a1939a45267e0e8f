cd /root/repo
REAL_PDL=1 timeout -s KILL 600 python -m pytest tests/test_attention_gpu.py tests/test_ops_gpu.py tests/test_gemm_gpu.py -x -q 2>&1 | tail -2
timeout -s KILL 500 python scripts/profile_gen.py 16 2>&1 | grep "^{" | cut -c1-150
REAL_PDL=1 timeout -s KILL 500 python scripts/profile_gen.py 128 2>&1 | grep "^{" | cut -c1-150
REAL_PDL=0 timeout -s KILL 500 python scripts/profile_gen.py 128 2>&1 | grep "^{" | cut -c1-150
