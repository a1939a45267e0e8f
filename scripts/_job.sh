set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "== nvls tests"; timeout 1500 python -m pytest tests/test_nvls_gpu.py -q -s 2>&1 | tail -40
echo "== comm tests (2 GPUs)"; timeout 900 python -m pytest tests/test_comm_gpu.py tests/test_parallel_gpu.py -q 2>&1 | tail -12
echo "== bench n2 (default: search-chosen gen layout)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --verbose 2> gpurun_out/bench_n2.err | tee gpurun_out/bench_n2_r2a.json; grep -v "^W09\|^\*\*\*\|^$" gpurun_out/bench_n2.err | tail -8
echo "== bench n2 gen-tp 2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 2 --gen-tp 2 --verbose 2> gpurun_out/bench_n2_tp2.err | tee gpurun_out/bench_n2_gentp2.json; grep -v "^W09\|^\*\*\*\|^$" gpurun_out/bench_n2_tp2.err | tail -8
