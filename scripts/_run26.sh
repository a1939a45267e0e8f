cd /root/repo
timeout -s KILL 600 python -m pytest tests/test_comm_gpu.py -x -q -s -k fused > gpurun_out/comm6.log 2>&1; tail -6 gpurun_out/comm6.log | cut -c1-700
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/bench_fused_tp.py > gpurun_out/fused_tp_bench2.log 2>&1; tail -2 gpurun_out/fused_tp_bench2.log | cut -c1-900
