#!/bin/bash
# PPO for batches that do not fit at once: every MFC splits its input into n_mbs micro-batches (token-balanced), which bounds
# activations / KV cache; training accumulates gradients over them inside each of the ppo_n_minibatches optimizer steps.
MODEL_FAMILY=llama
SFT_MODEL_PATH=${SFT_MODEL_PATH:?path to the SFT checkpoint}
RW_MODEL_PATH=${RW_MODEL_PATH:?path to the reward-model checkpoint}
python3 -m realhf_b200.apps.quickstart ppo \
    mode=local experiment_name=quickstart-ppo trial_name=$MODEL_FAMILY-local-minibatched \
    exp_ctrl.total_train_epochs=1 exp_ctrl.save_freq_steps=null n_nodes=1 allocation_mode=heuristic \
    actor.type._class=$MODEL_FAMILY actor.path=$SFT_MODEL_PATH \
    critic.type._class=$MODEL_FAMILY critic.type.is_critic=True critic.path=$RW_MODEL_PATH \
    ref.type._class=$MODEL_FAMILY ref.path=$SFT_MODEL_PATH \
    rew.type._class=$MODEL_FAMILY rew.type.is_critic=True rew.path=$RW_MODEL_PATH \
    dataset.path=.data/ppo_prompt.jsonl dataset.max_prompt_len=1024 dataset.train_bs_n_seqs=1024 \
    ppo.gen.max_new_tokens=1024 ppo.gen.min_new_tokens=1024 ppo.gen.use_cuda_graph=True ppo.gen.top_p=0.9 ppo.gen.top_k=1000 \
    ppo.ppo_n_minibatches=4 ppo.kl_ctl=0.1 ppo.value_eps_clip=0.2 \
    actor_gen.n_mbs=4 actor_train.n_mbs=8 critic_train.n_mbs=8 critic_inf.n_mbs=4 rew_inf.n_mbs=4 ref_inf.n_mbs=4
