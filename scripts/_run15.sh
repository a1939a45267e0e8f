cd /root/repo
timeout -s KILL 240 python -m pytest tests/test_gemm_gpu.py -x -q -k "cta_pair" > gpurun_out/pair1.log 2>&1; tail -5 gpurun_out/pair1.log
timeout -s KILL 300 python scripts/bench_gemm.py > gpurun_out/gemm_bench3.log 2>&1; cat gpurun_out/gemm_bench3.log | cut -c1-400
