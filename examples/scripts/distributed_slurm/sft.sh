#!/bin/bash
# SFT on a Slurm + pyxis cluster.  The cluster is described by a JSON spec (see examples/cluster_config.json): shared
# fileroot, partition, container images and mounts, node naming.  The launcher submits one job array per worker type.
export CLUSTER_SPEC_PATH=${CLUSTER_SPEC_PATH:?path to the cluster spec json}
MODEL_FAMILY=llama
MODEL_PATH=${MODEL_PATH:?path to the pretrained checkpoint (on the shared filesystem)}
python3 -m realhf_b200.apps.quickstart sft \
    mode=slurm experiment_name=quickstart-sft trial_name=$MODEL_FAMILY-slurm n_nodes=2 \
    exp_ctrl.total_train_epochs=8 exp_ctrl.save_freq_steps=50 exp_ctrl.eval_freq_epochs=1 \
    model.type._class=$MODEL_FAMILY model.path=$MODEL_PATH model.optimizer.lr=2e-5 model.optimizer.lr_scheduler_type=cosine \
    dataset.train_path=.data/sft_pos-train.jsonl dataset.valid_path=.data/sft_pos-valid.jsonl \
    dataset.max_seqlen=1024 dataset.train_bs_n_seqs=512 dataset.valid_bs_n_seqs=512 \
    allocation_mode=d16m1p1
