set -u
mkdir -p gpurun_out
echo "== zero3 gpu tests"; timeout 600 python -m pytest tests/test_zero3_gpu.py -q --tb=short 2>&1 | tail -6
echo "== sanitizer racecheck"; timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_kernels.py gemm norm misc > gpurun_out/sanitizer_racecheck.log 2>&1; tail -6 gpurun_out/sanitizer_racecheck.log | cut -c1-200
echo "== sanitizer memcheck"; timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_kernels.py > gpurun_out/sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck.log | cut -c1-200
echo "== bench n1 master runtime"; timeout 1500 python bench.py --runtime master --gpus 1 --steps 2 --warmup 2 2> gpurun_out/bench_master.err | grep "^{" | tee gpurun_out/bench_n1_master_runtime.json | cut -c1-1200; grep -v "^W09" gpurun_out/bench_master.err | tail -12 | cut -c1-250
