"""Short IK-action-mode run for ncu: PauseIKToggleEnv, always-follow-IK actions, one full wave of CTAs; by control step ~100 the arms
are grasping (arm + contact islands in most blocks)."""
import sys
import torch
sys.path.insert(0, ".")
from factory_marl_b200.vec_env import FactoryVecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2072
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 112
arms = int(sys.argv[3]) if len(sys.argv) > 3 else 2
sort = {"0": False, "1": True}.get(sys.argv[4]) if len(sys.argv) > 4 else None
env = FactoryVecEnv("PauseIKToggleEnv", num_envs=n, num_arms=arms, seeds=list(range(42, 42 + n)), sort_envs=sort)
env.reset()
a = torch.ones((n, arms), dtype=torch.int32, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
ev[0].record()
for t in range(steps):
    env.step(a, with_infos=False)
    ev[t + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
print("done launches", env.gpu_launches, "ms/step first 10: %.1f, steps 50-60: %.1f, last 10: %.1f" % (sum(ms[:10]) / 10, sum(ms[50:60]) / 10, sum(ms[-10:]) / 10))
print("cap flag envs", float(((env.core.i[env.core.off("I_FLAGS")] & 15) != 0).float().mean()), "episodes", int(env.monitor.ring_head.item()))
