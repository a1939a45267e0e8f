cd /root/repo
timeout -s KILL 300 python -m pytest tests/test_gemm_gpu.py -x -q -k streamk > gpurun_out/sk1.log 2>&1; tail -5 gpurun_out/sk1.log
timeout -s KILL 300 python -m pytest tests/test_sampling_gpu.py -x -q > gpurun_out/samp3.log 2>&1; tail -3 gpurun_out/samp3.log
timeout -s KILL 600 python scripts/bench_decode_gemm.py > gpurun_out/decode_gemm2.log 2>&1; tail -12 gpurun_out/decode_gemm2.log
