cd /root/repo
timeout -s KILL 600 python -m pytest tests/test_comm_gpu.py -x -q -s > gpurun_out/comm5.log 2>&1; tail -8 gpurun_out/comm5.log | cut -c1-700
