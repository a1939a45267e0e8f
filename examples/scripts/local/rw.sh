#!/bin/bash
# Reward modelling from an SFT checkpoint on one node (paired comparisons; same options as the reference's local/rw.sh).
MODEL_FAMILY=llama
SFT_MODEL_PATH=${SFT_MODEL_PATH:?path to the SFT checkpoint}
python3 -m realhf_b200.apps.quickstart rw \
    mode=local experiment_name=quickstart-rw trial_name=$MODEL_FAMILY-local-manual \
    exp_ctrl.total_train_epochs=1 exp_ctrl.save_freq_steps=5 exp_ctrl.eval_freq_epochs=1 \
    model.type._class=$MODEL_FAMILY model.type.is_critic=True model.path=$SFT_MODEL_PATH model.init_critic_from_actor=True \
    model.optimizer.lr=1e-5 model.optimizer.lr_scheduler_type=cosine \
    dataset.train_path=.data/rm_paired-train.jsonl dataset.valid_path=.data/rm_paired-valid.jsonl \
    dataset.max_pairs_per_prompt=2 dataset.max_seqlen=1024 dataset.train_bs_n_seqs=512 dataset.valid_bs_n_seqs=512 \
    allocation_mode=manual allocation.parallel.data_parallel_size=8 allocation.parallel.model_parallel_size=1 \
    allocation.parallel.pipeline_parallel_size=1
