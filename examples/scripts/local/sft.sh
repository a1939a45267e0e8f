#!/bin/bash
python3 -m realhf_b200.apps.quickstart sft mode=local experiment_name=quickstart-sft trial_name=llama-local \
    exp_ctrl.total_train_epochs=8 exp_ctrl.save_freq_steps=50 exp_ctrl.eval_freq_epochs=1 \
    model.type._class=llama model.path=${MODEL_PATH:?} dataset.train_path=.data/sft_pos-train.jsonl dataset.valid_path=.data/sft_pos-valid.jsonl \
    dataset.max_seqlen=1024 dataset.train_bs_n_seqs=512 allocation_mode=d8m1p1
