set -u
mkdir -p gpurun_out
nvidia-smi -L | head -8
NP=${NP:-8}
echo "== nvls/comm tests at world=$NP"; timeout 1500 python -m pytest tests/test_nvls_gpu.py -q -k "[$NP]" --tb=short 2>&1 | grep -v "W921\|NCCL version" | tail -25
echo "== decode timeline tp=$NP B=128"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29521 scripts/profile_decode_timeline.py 4 128 384 > gpurun_out/decode_timeline_tp${NP}_b128.jsonl 2> gpurun_out/tl.err; head -1 gpurun_out/decode_timeline_tp${NP}_b128.jsonl; tail -3 gpurun_out/tl.err | cut -c1-300
echo "== bench n=$NP default"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NP --steps 2 --warmup 3 --verbose 2> gpurun_out/bench_n${NP}.err | tee gpurun_out/bench_n${NP}_default.json | cut -c1-200; grep "warmup" gpurun_out/bench_n${NP}.err | tail -3
echo "== bench n=$NP gen-tp $NP"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NP --steps 2 --warmup 3 --gen-tp $NP --verbose 2> gpurun_out/bench_n${NP}_tp.err | tee gpurun_out/bench_n${NP}_gentp${NP}.json | cut -c1-200; grep "warmup" gpurun_out/bench_n${NP}_tp.err | tail -3
