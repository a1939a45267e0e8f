cd /root/repo
timeout 300 python -m pytest tests/test_sampling_gpu.py -x -q > gpurun_out/samp2.log 2>&1; tail -3 gpurun_out/samp2.log
timeout 900 python scripts/bench_decode_gemm.py > gpurun_out/decode_gemm1.log 2>&1; tail -12 gpurun_out/decode_gemm1.log
