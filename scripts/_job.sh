set -u
mkdir -p gpurun_out
echo "== mixtral ep (2 GPUs, 4 layers, eager decode)"; timeout 400 python scripts/bench_configs.py mixtral-ep --gpus 2 --layers 4 --steps 2 --warmup 1 --gen-graph false 2> gpurun_out/cfg_moe2.err | grep "^{" | tee gpurun_out/cfg_mixtral_ep.json | cut -c1-900; grep -n "tail of model_worker/1" -A45 gpurun_out/cfg_moe2.err | cut -c1-300 | tail -60
