#!/bin/bash
# The master/worker runtime on N GPUs under three allocations of the headline PPO config: every MFC data-parallel over all GPUs (what
# the SPMD arm runs), the heuristic, and the allocation search (per-MFC device meshes + parameter reallocation; with
# exp_ctrl.max_inflight_steps=2 the generation of step s+1 overlaps critic_train of step s when they live on different GPUs).
#     gpurun --gpus 8 --timeout 2400 -- 'NP=8 bash scripts/gpu_jobs/master_runtime_allocations.sh'
# NOT yet run on hardware (written after the round's GPU minutes were spent): the walk across steps and the two-phase dispatch are
# covered on CPU by tests/test_master_schedule_cpu.py and the gloo system tests.  Results to keep: gpurun_out/master_n*_*.json.
set -u
mkdir -p gpurun_out
NP=${NP:-8}
for alloc in dp search heuristic; do
  echo "== master runtime, allocation=$alloc, look-ahead 2"
  timeout 700 python bench.py --runtime master --gpus $NP --steps 3 --warmup 2 --allocation $alloc 2> gpurun_out/master_n${NP}_${alloc}.err \
    | grep "^{" | tee gpurun_out/master_n${NP}_${alloc}.json | cut -c1-400
  tail -3 gpurun_out/master_n${NP}_${alloc}.err | cut -c1-300
done
echo "== master runtime, allocation=search, barrier after every step"
REAL_MASTER_INFLIGHT_STEPS=1 timeout 700 python bench.py --runtime master --gpus $NP --steps 3 --warmup 2 --allocation search 2> gpurun_out/master_n${NP}_search_barrier.err \
  | grep "^{" | tee gpurun_out/master_n${NP}_search_barrier.json | cut -c1-400
