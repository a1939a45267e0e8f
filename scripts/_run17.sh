cd /root/repo
timeout -s KILL 600 python -m pytest tests/test_comm_gpu.py -x -q -s > gpurun_out/comm3.log 2>&1; tail -6 gpurun_out/comm3.log | cut -c1-600
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_n2_a.log 2>&1; tail -2 gpurun_out/bench_n2_a.log | cut -c1-1500
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --gen-tp 2 > gpurun_out/bench_n2_tp2.log 2>&1; tail -2 gpurun_out/bench_n2_tp2.log | cut -c1-1500
