set -u
mkdir -p gpurun_out
echo "== multi-GPU tests (2 GPUs): tp decode, nvls zero, pp graph"; timeout 900 python -m pytest tests/test_nvls_gpu.py tests/test_parallel_gpu.py -q -k "tp_decode or zero1 or pipelined_generation" --tb=short 2>&1 | grep -v "W921\|NCCL version" | tail -8
echo "== bench n2 default"; REAL_GEN_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --verbose 2> gpurun_out/bench_n2.err | tee gpurun_out/bench_n2_r2b.json | cut -c1-200; grep "warmup" gpurun_out/bench_n2.err | tail -2 | cut -c1-400
echo "== dpo zero3 (2 GPUs, 8 layers)"; timeout 900 python scripts/bench_configs.py dpo-zero3 --gpus 2 --layers 8 --steps 2 --warmup 1 2> gpurun_out/cfg_dpo.err | grep "^{" | tee gpurun_out/cfg_dpo_zero3.json | cut -c1-700; tail -3 gpurun_out/cfg_dpo.err | cut -c1-300
echo "== mixtral ep (2 GPUs, 4 layers)"; timeout 900 python scripts/bench_configs.py mixtral-ep --gpus 2 --layers 4 --steps 2 --warmup 1 2> gpurun_out/cfg_moe.err | grep "^{" | tee gpurun_out/cfg_mixtral_ep.json | cut -c1-700; tail -3 gpurun_out/cfg_moe.err | cut -c1-300
