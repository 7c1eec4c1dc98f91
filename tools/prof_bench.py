"""ncu target: the bench workload in its steady state (N(0,1)-clipped actions, like the fixed random policy), one launch."""
import sys
import torch
sys.path.insert(0, ".")
from factory_marl_b200.vec_env import FactoryVecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
E = int(sys.argv[3]) if len(sys.argv) > 3 else 0
env = FactoryVecEnv("AllFullRLProgressRewardEnv", num_envs=n, num_arms=2, seeds=list(range(42, 42 + n)), envs_per_block=E,
                    gripper_to_closest_cube_reward_factor=0.2, closest_cube_to_bucket_reward_factor=0.4, base_reward=0.4)
env.reset()
gen = torch.Generator(device="cuda").manual_seed(0)
nterm = 0
for t in range(steps):
    _, _, term, _ = env.step(torch.randn((n, 16), device="cuda", generator=gen).clamp(-1, 1))
    nterm += int(term.sum())
torch.cuda.synchronize()
c = env.core
print("done", env.gpu_launches, "terminations", nterm, "of", n * steps, "in-scene mean", float(c.i[c.off("I_NIN")].float().mean()),
      "ep_len mean", float(c.i[c.off("I_EP_LEN")].float().mean()), "maxcon mean", float(c.i[c.off("I_MAXCON")].float().mean()))
