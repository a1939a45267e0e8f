#!/bin/bash
# Batch generation from a checkpoint; results go to <fileroot>/logs/.../output.jsonl unless output_file is given.
MODEL_FAMILY=llama
MODEL_PATH=${MODEL_PATH:?path to the checkpoint}
python3 -m realhf_b200.apps.quickstart gen \
    mode=local experiment_name=quickstart-gen trial_name=$MODEL_FAMILY-local \
    model.type._class=$MODEL_FAMILY model.path=$MODEL_PATH \
    dataset.path=.data/ppo_prompt.jsonl dataset.max_prompt_len=1024 dataset.train_bs_n_seqs=128 \
    gen.max_new_tokens=1024 gen.min_new_tokens=1 gen.top_p=0.9 gen.top_k=1000 gen.use_cuda_graph=True \
    allocation_mode=manual allocation.parallel.data_parallel_size=8
