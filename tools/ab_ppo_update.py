"""Same-box A/B of the PPO update: forward of the update on the fused tcgen05 kernel (mode 3) or on the mma.sync GEMM chain (mode 1).
usage: python tools/ab_ppo_update.py [n_envs] [n_steps]"""
import sys
import torch
sys.path.insert(0, ".")
from factory_marl_b200.ppo import PPO
from factory_marl_b200.vec_env import FactoryVecEnv

n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 32
kw = dict(gripper_to_closest_cube_reward_factor=0.2, closest_cube_to_bucket_reward_factor=0.4, base_reward=0.4)
env = FactoryVecEnv("AllFullRLProgressRewardEnv", num_envs=n_envs, num_arms=2, seeds=list(range(42, 42 + n_envs)), **kw)
algo = PPO(env, n_steps=n_steps, batch_size=8192, n_epochs=10)
algo.collect_rollouts()
for mode in (3, 1, 3, 1):
    algo.pol.lib.fm_ppo_use_tcgen05(mode)
    algo._graphs = None
    algo.train()                      # launch by launch, then captures the graphs for this mode
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record(); i1 = algo.train(); ev[1].record(); algo.train(); ev[2].record(); torch.cuda.synchronize()
    print("mode %d: update %.1f / %.1f ms for %d optimiser steps (graph %s), approx_kl of the first timed pass %.5f" %
          (mode, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), i1["n_updates"], i1["cuda_graph"], i1["approx_kl"]))
algo.pol.lib.fm_ppo_use_tcgen05(3)
