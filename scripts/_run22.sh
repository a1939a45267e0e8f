cd /root/repo
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests4.log 2>&1; tail -3 gpurun_out/gpu_tests4.log
timeout -s KILL 500 python scripts/profile_gen.py 16 > gpurun_out/profile_gen_b16_v3.log 2>&1; grep -A6 "^{" gpurun_out/profile_gen_b16_v3.log | cut -c1-150
REAL_PDL=0 timeout -s KILL 500 python scripts/profile_gen.py 16 > gpurun_out/profile_gen_b16_nopdl.log 2>&1; grep "^{" gpurun_out/profile_gen_b16_nopdl.log | cut -c1-150
timeout -s KILL 900 python scripts/profile_decode.py > gpurun_out/profile_decode_tc4.log 2>&1; tail -6 gpurun_out/profile_decode_tc4.log
