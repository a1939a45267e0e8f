set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "== sampling tests"; timeout 600 python -m pytest tests/test_sampling_gpu.py -x -q 2>&1 | tail -8
echo "== nvls tests"; timeout 900 python -m pytest tests/test_nvls_gpu.py -x -q -s 2>&1 | tail -25
echo "== comm tests (2 GPUs)"; timeout 900 python -m pytest tests/test_comm_gpu.py -x -q 2>&1 | tail -8
echo "== bench n2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --verbose 2> gpurun_out/bench_n2.err | tee gpurun_out/bench_n2_r2a.json; grep -v "^W09\|^\*\*\*\|^$" gpurun_out/bench_n2.err | tail -12
