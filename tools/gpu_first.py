"""First GPU contact: smoke parity vs oracle, then raw timing of the step kernel at a few sizes / block shapes."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import __graft_entry__ as g
g.smoke()
from factory_marl_b200.vec_env import FactoryVecEnv
res = []
for (n, E, outer, inner) in [(4096, 4, 0, 0), (4096, 1, 0, 0), (16384, 1, 0, 0), (16384, 4, 0, 0)]:
    env = FactoryVecEnv("AllFullRLProgressRewardEnv", num_envs=n, num_arms=2, seeds=list(range(42, 42 + n)), envs_per_block=E,
                        pgs_outer=outer, pgs_inner=inner, gripper_to_closest_cube_reward_factor=0.2,
                        closest_cube_to_bucket_reward_factor=0.4, base_reward=0.4)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(0)
    acts = [torch.rand((n, 16), device="cuda", generator=gen) * 2 - 1 for _ in range(8)]
    for t in range(30):
        env.step(acts[t % 8])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 20
    s.record()
    for t in range(K):
        env.step(acts[t % 8])
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / K
    fl = int(env.core.i[env.core.off("I_FLAGS")].max()); mc = int(env.core.i[env.core.off("I_MAXCON")].max())
    r = dict(n=n, E=E, outer=outer, inner=inner, ms_per_step=ms, env_steps_per_s=n / ms * 1e3, flags=fl, maxcon=mc,
             episodes=int(env.core.i[env.core.off("I_EPISODES")].sum()))
    print(json.dumps(r), flush=True)
    res.append(r)
    env.close()
json.dump(res, open("gpurun_out/first.json", "w"), indent=1)
