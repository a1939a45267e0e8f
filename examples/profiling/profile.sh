#!/bin/bash
# Time model function calls on mock data over a grid of batch shapes (no dataset, no reward model needed): every handle is
# run `repeats` times per grid point after one warm-up call; results are printed as a table and written as JSONL, the format
# the allocation search's cost model reads.
#
# REAL_DUMP_TRACE=1 additionally writes one PyTorch-profiler chrome trace per call, REAL_DUMP_MEMORY=1 an allocator snapshot.
MODEL_FAMILY=llama
MODEL_PATH=${MODEL_PATH:-}          # empty: random weights of the llama-7b shape below
python3 -m realhf_b200.apps.quickstart profile \
    experiment_name=profile-example trial_name=test \
    model.type._class=$MODEL_FAMILY model.path=$MODEL_PATH \
    interface=ppo_actor 'handles=[generate,inference,train_step]' \
    'batch_sizes=[32,128]' 'seqlens=[640]' 'n_mbs=[1,2,4]' repeats=3 \
    gen.max_new_tokens=512 gen.min_new_tokens=512 gen.use_cuda_graph=True \
    output_file=profile_results.jsonl
