#!/bin/bash
# PPO with a hand-written allocation: generation and the frozen models' inference use layouts different from training, so
# the actor's and critic's weights are re-laid-out (parameter reallocation over NVLink peer stores) around every call.
# Device meshes name GPUs as "<node>:<gpu list>" inside the cluster; unspecified MFCs use the whole node.
MODEL_FAMILY=llama
SFT_MODEL_PATH=${SFT_MODEL_PATH:?path to the SFT checkpoint}
RW_MODEL_PATH=${RW_MODEL_PATH:?path to the reward-model checkpoint}
python3 -m realhf_b200.apps.quickstart ppo \
    mode=local experiment_name=quickstart-ppo trial_name=$MODEL_FAMILY-local-manual \
    exp_ctrl.total_train_epochs=1 exp_ctrl.save_freq_steps=null n_nodes=1 allocation_mode=manual \
    actor.type._class=$MODEL_FAMILY actor.path=$SFT_MODEL_PATH \
    critic.type._class=$MODEL_FAMILY critic.type.is_critic=True critic.path=$RW_MODEL_PATH \
    ref.type._class=$MODEL_FAMILY ref.path=$SFT_MODEL_PATH \
    rew.type._class=$MODEL_FAMILY rew.type.is_critic=True rew.path=$RW_MODEL_PATH \
    dataset.path=.data/ppo_prompt.jsonl dataset.max_prompt_len=128 dataset.train_bs_n_seqs=128 \
    ppo.gen.max_new_tokens=512 ppo.gen.min_new_tokens=512 ppo.gen.use_cuda_graph=True ppo.gen.top_p=0.9 ppo.gen.top_k=1000 \
    ppo.ppo_n_minibatches=4 ppo.kl_ctl=0.1 ppo.value_eps_clip=0.2 ppo.reward_output_scaling=1.0 ppo.adv_norm=True ppo.value_norm=True \
    actor_gen.parallel.data_parallel_size=8 \
    actor_train.parallel.data_parallel_size=4 actor_train.parallel.model_parallel_size=2 actor_train.parallel.use_sequence_parallel=True \
    critic_train.parallel.data_parallel_size=4 critic_train.parallel.model_parallel_size=2 critic_train.parallel.use_sequence_parallel=True \
    critic_inf.parallel.data_parallel_size=4 critic_inf.device_mesh=NODE01:0,1,2,3 \
    rew_inf.parallel.data_parallel_size=2 rew_inf.device_mesh=NODE01:4,5 \
    ref_inf.parallel.data_parallel_size=2 ref_inf.device_mesh=NODE01:6,7
