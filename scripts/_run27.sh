cd /root/repo
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests6.log 2>&1; tail -3 gpurun_out/gpu_tests6.log
timeout -s KILL 400 python scripts/profile_train.py 8 > gpurun_out/profile_train2.log 2>&1; grep -A12 "^{" gpurun_out/profile_train2.log | cut -c1-160
timeout -s KILL 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_full6.log 2>&1; tail -1 gpurun_out/bench_full6.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['mfc_ms'], d['clocks'])"
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
