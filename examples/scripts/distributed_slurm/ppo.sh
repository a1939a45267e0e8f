#!/bin/bash
# PPO over 4 nodes (32 GPUs).  `heuristic` derives one layout per model function call from the model sizes and the 180 GB of
# each GPU; `search` runs the MCMC allocation search instead; `ppo_manual.sh` shows a hand-written allocation.
# recover_mode=auto restarts the whole run from the states saved at a failure (weights, optimizer, LR step, data position,
# KL controller, value normaliser), up to recover_retries times.
export CLUSTER_SPEC_PATH=${CLUSTER_SPEC_PATH:?path to the cluster spec json}
MODEL_FAMILY=llama
SFT_MODEL_PATH=${SFT_MODEL_PATH:?path to the SFT checkpoint}
RW_MODEL_PATH=${RW_MODEL_PATH:?path to the reward-model checkpoint}
python3 -m realhf_b200.apps.quickstart ppo \
    mode=slurm experiment_name=quickstart-ppo trial_name=$MODEL_FAMILY-slurm-heuristic n_nodes=4 \
    exp_ctrl.total_train_epochs=1 exp_ctrl.save_freq_steps=null allocation_mode=heuristic recover_mode=auto recover_retries=2 \
    actor.type._class=$MODEL_FAMILY actor.path=$SFT_MODEL_PATH \
    critic.type._class=$MODEL_FAMILY critic.type.is_critic=True critic.path=$RW_MODEL_PATH \
    ref.type._class=$MODEL_FAMILY ref.path=$SFT_MODEL_PATH \
    rew.type._class=$MODEL_FAMILY rew.type.is_critic=True rew.path=$RW_MODEL_PATH \
    dataset.path=.data/ppo_prompt.jsonl dataset.max_prompt_len=128 dataset.train_bs_n_seqs=512 \
    ppo.gen.max_new_tokens=512 ppo.gen.min_new_tokens=512 ppo.gen.use_cuda_graph=True ppo.gen.top_p=0.9 ppo.gen.top_k=1000 \
    ppo.ppo_n_minibatches=4 ppo.kl_ctl=0.1 ppo.value_eps_clip=0.2 ppo.reward_output_scaling=1.0 ppo.adv_norm=True ppo.value_norm=True
