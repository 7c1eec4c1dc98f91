import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from factory_marl_b200.vec_env import FactoryVecEnv
n = 4096
dev = torch.device("cuda", 0)
kw = dict(gripper_to_closest_cube_reward_factor=0.2, closest_cube_to_bucket_reward_factor=0.4, base_reward=0.4)
env = FactoryVecEnv("AllFullRLProgressRewardEnv", num_envs=n, num_arms=2, seeds=list(range(42, 42 + n)), device=dev, **kw)
env.reset()
rng = np.random.default_rng(0)
Ke = 12
host_acts = [torch.from_numpy(np.clip(rng.normal(0, 1, (n, 16)), -1, 1).astype(np.float32)).pin_memory() for _ in range(Ke)]
for k in range(230):                     # de-synchronise like bench.py: staggered resets over the first 200 steps
    env.step(host_acts[k % Ke].to(dev), with_infos=False)
    if k < 200:
        mask = torch.zeros(n, dtype=torch.uint8, device=dev); mask[k::200] = 1
        env.core.reset(mask)
torch.cuda.synchronize()
c = env.core
D = c.obs_dim
host_obs = [torch.empty((n, D), dtype=torch.float32).pin_memory() for _ in range(2)]
host_rew = [torch.empty((n,), dtype=torch.float32).pin_memory() for _ in range(2)]
host_done = [torch.empty((n,), dtype=torch.uint8).pin_memory() for _ in range(2)]
outs = [dict(obs=c.obs, reward=c.reward, terminated=c.terminated, scores=c.scores, flags=c.flags, terminal_obs=c.terminal_obs, ep_info=c.ep_info)]
outs.append({k: torch.zeros_like(v) for k, v in outs[0].items()})
copy_stream = torch.cuda.Stream(device=dev)
main = torch.cuda.current_stream(dev)

def loop(mode):
    torch.cuda.synchronize()
    evk = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Ke)]
    done_ev = [None, None]; copies = []
    t0 = time.perf_counter()
    for k in range(Ke):
        b = k % 2
        for name, t in outs[b if mode != "sync" else 0].items():
            setattr(c, name, t)
        if mode != "sync" and done_ev[b] is not None:
            main.wait_event(done_ev[b])
        a_dev = host_acts[k].to(dev, non_blocking=True)
        evk[k][0].record(main)
        o2, rew, term, _ = env.step(a_dev, with_infos=False)
        evk[k][1].record(main)
        if mode == "sync":
            host_obs[b].copy_(o2, non_blocking=True); host_rew[b].copy_(rew, non_blocking=True); host_done[b].copy_(c.terminated, non_blocking=True)
            main.synchronize()
        else:
            ready = torch.cuda.Event(enable_timing=True); ready.record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ready)
                host_obs[b].copy_(o2, non_blocking=True); host_rew[b].copy_(rew, non_blocking=True); host_done[b].copy_(c.terminated, non_blocking=True)
                done_ev[b] = torch.cuda.Event(enable_timing=True); done_ev[b].record(copy_stream); copies.append((k, ready, done_ev[b]))
            if mode == "pipe" and k >= 1:
                done_ev[1 - b].synchronize()
            if mode == "pipe2" and k >= 2:
                done_ev[b].synchronize() if False else done_ev[1 - b].synchronize()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for name, t in outs[0].items():
        setattr(c, name, t)
    ks = [a.elapsed_time(b) for a, b in evk]
    if copies: print("   copy k finishes this long after kernel k ended (ms):", " ".join("%.2f" % r.elapsed_time(d) for k, r, d in copies))
    print("%-6s wall %.2f ms/step; kernel-span per step: %s" % (mode, (t1 - t0) * 1e3 / Ke, " ".join("%.1f" % x for x in ks)))

for mode in ("sync", "pipe", "pipe", "sync", "sync", "pipe", "pipe", "sync"):
    loop(mode)
