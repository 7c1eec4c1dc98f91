"""ncu target: PPO optimiser steps (launch by launch) on a 4096-env rollout buffer; prints nothing but runs 1 train() call."""
import sys
import torch
sys.path.insert(0, ".")
from factory_marl_b200.vec_env import FactoryVecEnv
from factory_marl_b200.ppo import PPO
env = FactoryVecEnv("AllFullRLProgressRewardEnv", num_envs=4096, num_arms=2, seed=0, gripper_to_closest_cube_reward_factor=0.2,
                    closest_cube_to_bucket_reward_factor=0.4, base_reward=0.4)
algo = PPO(env, n_steps=2, batch_size=8192, n_epochs=2, seed=0, use_graph=False)
algo.collect_rollouts(); algo.train()
torch.cuda.synchronize()
print("done")
