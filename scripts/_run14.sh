cd /root/repo
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_all.log 2>&1; tail -4 gpurun_out/gpu_tests_all.log
timeout -s KILL 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_full4.log 2>&1; tail -2 gpurun_out/bench_full4.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 3 -c 1 -o gpurun_out/prof_gemm_big -f python scripts/ncu_gemm.py big > gpurun_out/ncu_big.log 2>&1; tail -2 gpurun_out/ncu_big.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gemm_streamk_kernel -s 3 -c 1 -o gpurun_out/prof_gemm_smallm -f python scripts/ncu_gemm.py small > gpurun_out/ncu_small.log 2>&1; tail -2 gpurun_out/ncu_small.log
ls -la gpurun_out/*.ncu-rep
