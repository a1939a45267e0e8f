cd /root/repo
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/bench_n8_b.log 2>&1; tail -1 gpurun_out/bench_n8_b.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['mfc_ms'], d['clocks'], d['memory'])"
