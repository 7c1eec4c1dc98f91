"""Throughput of the other BASELINE.json configurations (not the bench line): IK-action mode with IK toggling, 2 and 4 arms,
4096 envs, spawning on, random toggle actions; env.step includes the IK policy kernel.  Prints env-steps/s per config."""
import json, sys
import torch
sys.path.insert(0, ".")
from factory_marl_b200.vec_env import FactoryVecEnv
res = []
for cls, arms, E in [("PauseIKToggleEnv", 2, 0), ("PauseIKToggleEnv", 4, 0), ("PauseIKToggleEnv", 4, 4), ("FactoryManipulationEnv", 2, 0), ("AllDeltaProgressRewardEnv", 2, 0)]:
    n = 4096
    env = FactoryVecEnv(cls, num_envs=n, num_arms=arms, seeds=list(range(42, 42 + n)), envs_per_block=E)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(0)
    def act():
        if cls == "PauseIKToggleEnv":
            return torch.randint(0, 2, (n, arms), device="cuda", generator=gen, dtype=torch.int32)
        if cls == "FactoryManipulationEnv":
            return torch.zeros((n, 1), device="cuda")
        return torch.randn((n, 8 * arms), device="cuda", generator=gen).clamp(-1, 1)
    for t in range(40):
        env.step(act())
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 15
    s.record()
    for t in range(K):
        env.step(act())
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / K
    fl = int(env.core.i[env.core.off("I_FLAGS")].max())
    r = dict(env_class=cls, n_arms=arms, envs_per_block=E, n=n, ms_per_step=ms, env_steps_per_s=n / ms * 1e3, flags=fl,
             mean_score=float(env.core.scores.float().sum(1).mean()))
    print(json.dumps(r), flush=True)
    res.append(r)
    env.close()
json.dump(res, open("gpurun_out/configs.json", "w"), indent=1)
