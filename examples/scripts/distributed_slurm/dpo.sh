#!/bin/bash
export CLUSTER_SPEC_PATH=${CLUSTER_SPEC_PATH:?path to the cluster spec json}
MODEL_FAMILY=llama
SFT_MODEL_PATH=${SFT_MODEL_PATH:?path to the SFT checkpoint}
python3 -m realhf_b200.apps.quickstart dpo \
    mode=slurm experiment_name=quickstart-dpo trial_name=$MODEL_FAMILY-slurm n_nodes=2 \
    exp_ctrl.total_train_epochs=2 exp_ctrl.save_freq_steps=5 \
    actor.type._class=$MODEL_FAMILY actor.path=$SFT_MODEL_PATH actor.optimizer.lr=2e-6 \
    ref.type._class=$MODEL_FAMILY ref.path=$SFT_MODEL_PATH \
    dataset.train_path=.data/rm_paired-train.jsonl dataset.max_pairs_per_prompt=2 dataset.max_seqlen=1024 dataset.train_bs_n_seqs=512 \
    beta=0.1 allocation_mode=heuristic
