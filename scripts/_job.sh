set -u
mkdir -p gpurun_out
echo "== mixtral ep (2 GPUs, 4 layers, graph decode)"; timeout 300 python scripts/bench_configs.py mixtral-ep --gpus 2 --layers 4 --steps 2 --warmup 1 2> gpurun_out/cfg_moe.err | grep "^{" | tee gpurun_out/cfg_mixtral_ep.json | cut -c1-1000
if ! grep -q '"value"' gpurun_out/cfg_mixtral_ep.json; then
for f in /tmp/realhf_b200_cfgbench_*/fileroot/logs/*/t0/model_worker-*; do echo "--- $f"; grep -v "^frame #" $f | grep -n "Assert\|Error\|error\|File \"" | head -30 | cut -c1-300; done
fi
echo "== EP training without SP (2 GPUs)"; timeout 240 python -m pytest tests/test_nvls_gpu.py -q -x -k "without_sequence_parallel" --tb=short 2>&1 | grep -v "W921\|NCCL version" | tail -15 | cut -c1-400
