"""Short run for ncu: warm steps, then a few env steps of the bench workload at one full wave of CTAs."""
import sys
import torch
sys.path.insert(0, ".")
from factory_marl_b200.vec_env import FactoryVecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
E = int(sys.argv[3]) if len(sys.argv) > 3 else 0
env = FactoryVecEnv("AllFullRLProgressRewardEnv", num_envs=n, num_arms=2, seeds=list(range(42, 42 + n)), envs_per_block=E,
                    gripper_to_closest_cube_reward_factor=0.2, closest_cube_to_bucket_reward_factor=0.4, base_reward=0.4)
env.reset()
gen = torch.Generator(device="cuda").manual_seed(0)
for t in range(steps):
    env.step(torch.rand((n, 16), device="cuda", generator=gen) * 2 - 1)
torch.cuda.synchronize()
print("done", env.gpu_launches)
