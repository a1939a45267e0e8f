#!/bin/bash
# PPO with the allocation SEARCH: the C++ MCMC picks a device mesh and a (dp, tp, pp) layout per model function call, costed by
# the layer-profile table of the model (python -m realhf_b200.apps.profile_layers --family llama --size 7 writes a measured one;
# a B200 calibration for LLaMA-7B ships with the package) and by reallocation times taken from the real planner.  With the actor
# and the critic on different GPUs the generation of step s+1 overlaps critic_train of step s (exp_ctrl.max_inflight_steps=2).
# allocation_use_cache=true stores the result under the profiler cache, so the next launch of the same problem skips the search.
MODEL_FAMILY=llama
SFT_MODEL_PATH=${SFT_MODEL_PATH:?path to the SFT checkpoint}
RW_MODEL_PATH=${RW_MODEL_PATH:?path to the reward-model checkpoint}
python3 -m realhf_b200.apps.quickstart ppo \
    mode=local experiment_name=quickstart-ppo trial_name=$MODEL_FAMILY-local-search \
    exp_ctrl.total_train_epochs=1 exp_ctrl.save_freq_steps=null exp_ctrl.max_inflight_steps=2 n_nodes=1 \
    allocation_mode=search allocation_use_cache=true \
    actor.type._class=$MODEL_FAMILY actor.type.size=7 actor.path=$SFT_MODEL_PATH \
    critic.type._class=$MODEL_FAMILY critic.type.size=7 critic.type.is_critic=True critic.path=$RW_MODEL_PATH \
    ref.type._class=$MODEL_FAMILY ref.type.size=7 ref.path=$SFT_MODEL_PATH \
    rew.type._class=$MODEL_FAMILY rew.type.size=7 rew.type.is_critic=True rew.path=$RW_MODEL_PATH \
    dataset.path=.data/ppo_prompt.jsonl dataset.max_prompt_len=128 dataset.train_bs_n_seqs=128 \
    ppo.gen.max_new_tokens=512 ppo.gen.min_new_tokens=512 ppo.gen.use_cuda_graph=True ppo.gen.top_p=0.9 ppo.gen.top_k=1000 \
    ppo.ppo_n_minibatches=4 ppo.kl_ctl=0.1 ppo.value_eps_clip=0.2 ppo.reward_output_scaling=1.0 ppo.adv_norm=True ppo.value_norm=True
