#!/bin/bash
# Two-GPU job: multi-GPU tests that fit on 2 devices and the secondary BASELINE configurations at reduced depth through the
# master / worker runtime.
#     gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_jobs/two_gpu_configs.sh'
set -u
mkdir -p gpurun_out
echo "== multi-GPU tests (world 2)"; timeout 900 python -m pytest tests/test_nvls_gpu.py tests/test_parallel_gpu.py tests/test_comm_gpu.py -q --tb=short 2>&1 | grep -v "W921\|NCCL version" | tail -12 | cut -c1-400
echo "== dpo zero3 (8 layers)"; timeout 600 python scripts/bench_configs.py dpo-zero3 --gpus 2 --layers 8 --steps 2 --warmup 1 2> gpurun_out/cfg_dpo.err | grep "^{" | tee gpurun_out/config_dpo_zero3_2gpu.json | cut -c1-900
echo "== mixtral ep (4 layers, graph decode)"; timeout 600 python scripts/bench_configs.py mixtral-ep --gpus 2 --layers 4 --steps 2 --warmup 1 2> gpurun_out/cfg_moe.err | grep "^{" | tee gpurun_out/config_mixtral_ep_2gpu.json | cut -c1-900
for f in gpurun_out/cfg_dpo.err gpurun_out/cfg_moe.err; do grep -n "tail of\|Error" "$f" | head -5; done
