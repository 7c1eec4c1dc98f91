"""IK-base-policy statistics on the GPU for several solver settings (compare with tests/golden/ik_policy_stats.json)."""
import json, sys
import torch
sys.path.insert(0, ".")
from factory_marl_b200.vec_env import FactoryVecEnv
st = json.load(open("tests/golden/ik_policy_stats.json"))
N, H = 1024, st["horizon"]
print("oracle", st["mean_total"], st["ci95_total"], st["mean_length"], st["reasons"])
for kw in [dict(pgs_inner=4), dict(pgs_inner=8), dict(pgs_inner=32)]:
    env = FactoryVecEnv("FactoryManipulationEnv", num_envs=N, num_arms=2, seeds=list(range(42, 42 + N)), auto_reset=True, **kw)
    env.reset()
    done = torch.zeros(N, dtype=torch.bool, device="cuda"); total = torch.zeros(N, device="cuda"); length = torch.full((N,), float(H), device="cuda")
    nforce = torch.zeros(N, device="cuda")
    dummy = torch.zeros((N, 1), device="cuda")
    for t in range(H):
        obs, rew, term, _ = env.step(dummy)
        first = term & ~done
        ep = env.core.ep_info
        total = torch.where(first, ep[:, 2] + ep[:, 3], total)
        length = torch.where(first, torch.full_like(length, t + 1.0), length)
        nforce += (first & ((env.core.flags & 2) > 0)).float()
        done |= term
    total = torch.where(done, total, env.core.scores.float().sum(dim=1))
    print(kw, "mean total %.3f mean len %.1f force-term frac %.3f horizon frac %.3f flags %d" % (float(total.mean()), float(length.mean()), float(nforce.mean()), float((~done).float().mean()), int(env.core.i[env.core.off("I_FLAGS")].max())), "maxcon max/mean", int(env.core.i[env.core.off("I_MAXCON")].max()), float(env.core.i[env.core.off("I_MAXCON")].float().mean()), "envs with overflow flags", [int(((env.core.i[env.core.off("I_FLAGS")] & b) > 0).sum()) for b in (1, 2, 4)], flush=True)
