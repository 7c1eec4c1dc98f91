"""Which block sets the launch time?  Needs the profiling variant (python tools/ab_variant.py cyc -DFM_BLOCK_CYCLES, FM_B200_LIB=...):
the kernel leaves each block's cycle count in ep_info[env][0] of its non-terminated envs.  Bench workload, de-synchronised."""
import sys
import torch
sys.path.insert(0, ".")
from factory_marl_b200.ppo import PolicyValueNet
from factory_marl_b200.vec_env import FactoryVecEnv

n = 4096
kw = dict(gripper_to_closest_cube_reward_factor=0.2, closest_cube_to_bucket_reward_factor=0.4, base_reward=0.4)
env = FactoryVecEnv("AllFullRLProgressRewardEnv", num_envs=n, num_arms=2, seeds=list(range(42, 42 + n)), sort_envs=False, **kw)
pol = PolicyValueNet(env.core.obs_dim, env.core.act_dim, discrete=False, seed=0)
obs = env.reset()
c = env.core
# de-synchronise like bench.py: staggered resets over the first 200 steps
for t in range(230):
    obs, _, _, _ = env.step(pol.act(obs), with_infos=False)
    if t < 200:
        mask = torch.zeros(n, dtype=torch.uint8, device="cuda"); mask[t::200] = 1
        c.reset(mask)
rows = {k: c.off(k) for k in ("I_ARMCON", "I_PREV_N", "I_NIN", "I_FLAGS")}
for t in range(40):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pre = {k: c.i[r].clone() for k, r in rows.items()}
    a = pol.act(obs)
    e0.record(); obs, rew, term, _ = env.step(a, with_infos=False); e1.record(); torch.cuda.synchronize()
    cyc = c.ep_info[:, 0].clone(); cyc[c.terminated.bool()] = 0
    blk = cyc.view(-1)[: (n // 14) * 14].view(-1, 14).max(dim=1).values          # identity order: block b = envs 14 b .. 14 b + 13
    top = torch.topk(blk, 3)
    med = float(blk.median())
    b = int(top.indices[0]); envs = list(range(14 * b, 14 * b + 14))
    print("step %2d  %.2f ms  block cycles median %.2fM  top %s  | slowest block %d: armcon before %s after %s, contacts before %s after %s, cubes %s, terminated %s" % (
        t, e0.elapsed_time(e1), med / 1e6, [round(float(x) / 1e6, 2) for x in top.values], b,
        pre["I_ARMCON"][envs].tolist(), c.i[rows["I_ARMCON"]][envs].tolist(), pre["I_PREV_N"][envs].tolist(), c.i[rows["I_PREV_N"]][envs].tolist(),
        c.i[rows["I_NIN"]][envs].tolist(), c.terminated[envs].tolist()))
