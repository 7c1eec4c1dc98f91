import sys, torch
sys.path.insert(0, ".")
from factory_marl_b200.vec_env import FactoryVecEnv
n = 2050
kw = dict(gripper_to_closest_cube_reward_factor=0.2, closest_cube_to_bucket_reward_factor=0.4, base_reward=0.4)
gen = torch.Generator(device="cuda").manual_seed(11)
acts = [torch.randn((n, 16), device="cuda", generator=gen).mul(1.5).clamp(-1, 1) for _ in range(25)]
def run(sort, pgs_outer=0):
    env = FactoryVecEnv("AllFullRLProgressRewardEnv", num_envs=n, num_arms=2, seeds=list(range(42, 42 + n)), sort_envs=sort, pgs_outer=pgs_outer, **kw)
    env.reset()
    hist = []
    for a in acts:
        obs, rew, term, _ = env.step(a, with_infos=False)
        hist.append((obs.clone(), rew.clone()))
    return hist, env.core.i.clone()
for po in (0, 1):
    A, ia = run(False, po); B, ib = run(False, po); Csorted, ic = run(True, po)
    def first_diff(X, Y):
        for t, ((o1, r1), (o2, r2)) in enumerate(zip(X, Y)):
            if not (torch.equal(o1, o2) and torch.equal(r1, r2)):
                bad = torch.nonzero((o1 != o2).any(dim=1) | (r1 != r2)).flatten()
                return t, bad[:8].tolist(), float((o1 - o2).abs().max())
        return None
    print("pgs_outer", po, "unsorted twice:", first_diff(A, B), "| unsorted vs sorted:", first_diff(A, Csorted))
# detail of the first differing env: is it one that overflowed a contact capacity (cumulative I_FLAGS row)?
from factory_marl_b200.vec_env import FactoryVecEnv as _E
for rep in range(3):
    A, ia = run(False, 0); B, ib = run(False, 0)
    for t, ((o1, r1), (o2, r2)) in enumerate(zip(A, B)):
        bad = torch.nonzero((r1 != r2) | (o1 != o2).any(dim=1)).flatten()
        if len(bad):
            e = int(bad[0])
            env = _E("AllFullRLProgressRewardEnv", num_envs=8, num_arms=2)
            fl = env.core.off("I_FLAGS")
            print("rep", rep, "step", t, "env", e, "reward", float(r1[e]), float(r2[e]), "I_FLAGS", int(ia[fl][e]), int(ib[fl][e]), "n differing envs", len(bad))
            break
    else:
        print("rep", rep, "identical")
