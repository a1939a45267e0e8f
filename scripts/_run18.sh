cd /root/repo
timeout -s KILL 600 python -m pytest tests/test_comm_gpu.py -x -q -s -k allreduce > gpurun_out/comm4.log 2>&1; tail -6 gpurun_out/comm4.log | cut -c1-600
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --gen-tp 2 > gpurun_out/bench_n2_tp2b.log 2>&1; tail -1 gpurun_out/bench_n2_tp2b.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['mfc_ms'])"
