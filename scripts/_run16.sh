cd /root/repo
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_all2.log 2>&1; tail -4 gpurun_out/gpu_tests_all2.log
timeout -s KILL 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_full5.log 2>&1; tail -1 gpurun_out/bench_full5.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gemm_2cta_kernel -s 3 -c 1 -o gpurun_out/prof_gemm_pair -f python scripts/ncu_gemm.py big > gpurun_out/ncu_pair.log 2>&1; tail -2 gpurun_out/ncu_pair.log
