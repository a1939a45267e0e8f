#!/bin/bash
# PPO with ONE layout for every model function call (the "symmetric" baseline of the reference's benchmarks, and the layout
# bench.py times): d8m1p1 = data parallel over the node, no parameter reallocation, frozen models offloaded between calls.
MODEL_FAMILY=llama
SFT_MODEL_PATH=${SFT_MODEL_PATH:?path to the SFT checkpoint}
RW_MODEL_PATH=${RW_MODEL_PATH:?path to the reward-model checkpoint}
python3 -m realhf_b200.apps.quickstart ppo \
    mode=local experiment_name=quickstart-ppo trial_name=$MODEL_FAMILY-local-symm \
    exp_ctrl.total_train_epochs=1 exp_ctrl.save_freq_steps=null n_nodes=1 allocation_mode=d8m1p1 \
    actor.type._class=$MODEL_FAMILY actor.path=$SFT_MODEL_PATH \
    critic.type._class=$MODEL_FAMILY critic.type.is_critic=True critic.path=$RW_MODEL_PATH \
    ref.type._class=$MODEL_FAMILY ref.path=$SFT_MODEL_PATH \
    rew.type._class=$MODEL_FAMILY rew.type.is_critic=True rew.path=$RW_MODEL_PATH \
    dataset.path=.data/ppo_prompt.jsonl dataset.max_prompt_len=128 dataset.train_bs_n_seqs=128 \
    ppo.gen.max_new_tokens=512 ppo.gen.min_new_tokens=512 ppo.gen.use_cuda_graph=True ppo.gen.top_p=0.9 ppo.gen.top_k=1000 \
    ppo.ppo_n_minibatches=4 ppo.kl_ctl=0.1 ppo.value_eps_clip=0.2 ppo.reward_output_scaling=1.0 ppo.adv_norm=True ppo.value_norm=True
