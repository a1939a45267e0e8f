#!/bin/bash
# DPO on one node: the frozen reference runs inference with tp=2 x dp=4, the actor trains with ZeRO over dp=8; the two
# layouts share the node and the reference's weights are offloaded between its calls.
MODEL_FAMILY=llama
SFT_MODEL_PATH=${SFT_MODEL_PATH:?path to the SFT checkpoint}
python3 -m realhf_b200.apps.quickstart dpo \
    mode=local experiment_name=quickstart-dpo trial_name=$MODEL_FAMILY-local-manual \
    exp_ctrl.total_train_epochs=2 exp_ctrl.save_freq_steps=5 \
    actor.type._class=$MODEL_FAMILY actor.path=$SFT_MODEL_PATH actor.optimizer.lr=2e-6 \
    ref.type._class=$MODEL_FAMILY ref.path=$SFT_MODEL_PATH \
    dataset.train_path=.data/rm_paired-train.jsonl dataset.max_pairs_per_prompt=2 dataset.max_seqlen=1024 dataset.train_bs_n_seqs=512 \
    beta=0.1 allocation_mode=manual \
    actor_train.parallel.data_parallel_size=8 \
    ref_inf.parallel.data_parallel_size=4 ref_inf.parallel.model_parallel_size=2
