set -u
mkdir -p gpurun_out
echo "== fp8 + moe graph tests"; timeout 600 python -m pytest tests/test_fp8_gpu.py tests/test_moe_gpu.py -q --tb=short -x 2>&1 | grep -v "W921\|NCCL version" | tail -15
echo "== dpo zero3 (2 GPUs, 8 layers)"; timeout 600 python scripts/bench_configs.py dpo-zero3 --gpus 2 --layers 8 --steps 2 --warmup 1 2> gpurun_out/cfg_dpo.err | grep "^{" | tee gpurun_out/cfg_dpo_zero3.json | cut -c1-900; tail -3 gpurun_out/cfg_dpo.err | cut -c1-300
echo "== mixtral ep (2 GPUs, 4 layers, graph decode)"; timeout 600 python scripts/bench_configs.py mixtral-ep --gpus 2 --layers 4 --steps 2 --warmup 1 2> gpurun_out/cfg_moe.err | grep "^{" | tee gpurun_out/cfg_mixtral_ep.json | cut -c1-900; tail -3 gpurun_out/cfg_moe.err | cut -c1-300
if ! grep -q '"value"' gpurun_out/cfg_mixtral_ep.json; then
echo "== mixtral ep (eager decode fallback)"; timeout 600 python scripts/bench_configs.py mixtral-ep --gpus 2 --layers 4 --steps 2 --warmup 1 --gen-graph false 2> gpurun_out/cfg_moe2.err | grep "^{" | tee gpurun_out/cfg_mixtral_ep.json | cut -c1-900; tail -3 gpurun_out/cfg_moe2.err | cut -c1-300
fi
