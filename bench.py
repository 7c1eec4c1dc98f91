#!/usr/bin/env python
"""bench.py — headline benchmark of the batched TaskEnv hot path (BASELINE.json: env-steps/s; PPO samples/s).

  python bench.py --gpus N --steps K --warmup W            ours: CUDA step kernel (+ on-GPU policy), one rank per GPU
  python bench.py --impl reference --gpus N ...            reference arm: the fp64 CPU restatement (oracle port) on host cores

Headline (`value`, `e2e`, `roofline`): BASELINE.json configs[1] — 4096 batched TaskEnv per GPU (AllFullRLProgressRewardEnv action
map), two arms, fixed random policy; a "step" is one control step (100 physics sub-steps + task logic + obs/reward) of every env;
weak scaling over ranks (env shards independent, no data-path collective).  The other BASELINE configurations are measured in the
same process and reported as extra keys of the same JSON line: `cfg3` (4096 envs, IK-action mode, 2 and 4 arms), `cfg4` (PPO,
16 384 envs, declared default n_steps = 32), `cfg5` (65 536 envs sharded over the ranks: strong scaling, NCCL gradient all-reduce).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENV_CLASS = "AllFullRLProgressRewardEnv"
ENV_KW = dict(gripper_to_closest_cube_reward_factor=0.2, closest_cube_to_bucket_reward_factor=0.4, base_reward=0.4)  # runs/rk5rxnav.json:25-41
N_ENVS_PER_GPU = 4096
METRIC = "batched TaskEnv env-steps/sec"
UNIT = "env-steps/s"
# algorithmic work per env-step: SURVEY.md 8d hand count of the path as MuJoCo runs it (25 solver iterations over
# ~170 rows) and the fp32 flop this kernel really executes per env-step in the bench workload (ncu fadd + fmul + 2 ffma, profiles/)
FLOP_PER_ENV_STEP = 2.1e7
FLOP_PER_ENV_STEP_EXECUTED = 4.76e6
EPISODE_STEPS = 230            # pre-roll length: one episode of the zero/random-action workload (ends near control step 218)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured", d.get("sm_max_mhz", 1965.0)
    return 6650.0, "fallback", 1965.0


def ncu_dram_bytes_per_env():
    """dram__bytes_read.sum + dram__bytes_write.sum of fm_step_kernel per env-step from the committed `ncu --set full` capture of
    this round (profiles/r4k_step_kernel_dram.json: the end-state kernel; r2_step_kernel_dram.json: mid-round), else None"""
    p = os.path.join(ROOT, "profiles", "r4k_step_kernel_dram.json")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r2_step_kernel_dram.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["dram_bytes_per_launch"] / d["envs_per_launch"], d.get("source")
    return None, None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx, self.samples, self.reasons, self.stop_flag, self.max_mhz = gpu_index, [], set(), False, None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], out[2:6]):
                    if v.strip().lower() == "active":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ reference arm / cpu baseline
def _cpu_worker(args):
    seed, n_steps, warm = args
    from factory_marl_b200 import model as fm
    from oracle.oracle_py import Oracle
    o = Oracle(fm.build_blob(2))
    o.env_create(ENV_CLASS, seed=42 + seed, **ENV_KW)
    o.env_reset()
    rng = np.random.default_rng(seed)
    # same action statistics as the GPU arm's fixed random policy: N(~0, 1) clipped to [-1, 1]
    acts = np.clip(rng.normal(0, 1, (warm + n_steps, 16)), -1, 1).astype(np.float32)
    for t in range(warm):
        _, _, info = o.env_step(acts[t])
        if info["terminated"]:
            o.env_reset()
    t0 = time.perf_counter()
    for t in range(n_steps):
        _, _, info = o.env_step(acts[warm + t])
        if info["terminated"]:
            o.env_reset()
    return time.perf_counter() - t0


def cpu_reference_rate(steps_per_env: int, warm: int = 3):
    """The fp64 CPU restatement of the reference path (oracle/, built -O3 -march=native), one env per host core on all cores, wall
    clock.  It is a PORT: the reference's own stack (Python + dm_control + MuJoCo 3.2.0 + SubprocVecEnv pipes) is not installable
    in this image, so how the two compare is not known; SURVEY.md 8a7 estimates 30-80 us per mj_step for MuJoCo."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        times = pool.map(_cpu_worker, [(i, steps_per_env, warm) for i in range(cores)])
    rate = cores * steps_per_env / max(times)
    us_per_mj_step = 1e6 * max(times) / (steps_per_env * 100)
    sample = (f"{cores} envs x {steps_per_env} control steps (100 sub-steps each) of the bench workload, one env per core, clipped N(0,1) "
              f"actions; {us_per_mj_step:.0f} us per mj_step-equivalent per core (exact Newton to convergence, fp64)")
    return rate, cores, sample, us_per_mj_step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = 4           # control steps per env per "step" of this arm (bounded sample)
    K, W = args.steps, args.warmup
    rate, cores, sample, us = cpu_reference_rate(per_step * K, warm=max(1, W))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * cores * per_step / rate, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{ENV_CLASS}, 2 arms, clipped N(0,1) actions, fp64 CPU restatement of the MuJoCo TaskEnv (the reference's MuJoCo/dm_control/SB3 stack is not installable here)",
                       "envs": cores, "steps_per_env_per_step": per_step, "us_per_mj_step_equivalent": us},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def desync_preroll(env, policy_fn, n_steps, obs, stagger=True):
    """Untimed pre-roll that de-synchronises the envs: during `n_steps` (one episode length) control steps, env e is reset at step
    e mod n_steps, so that afterwards the batch samples every episode phase uniformly (SURVEY.md 8d config 2 asks for
    de-synchronised envs; throughput drifts by ~10 % along an episode)."""
    import torch
    n = env.num_envs
    ids = torch.arange(n, device=env.device)
    for t in range(n_steps):
        if stagger:
            mask = ((ids % n_steps) == t).to(torch.uint8)
            env.core.reset(mask)
            obs = env.core.obs
        act = policy_fn(obs)
        obs, _, _, _ = env.step(act, with_infos=False)
    return obs


def time_env_steps(env, policy_fn, obs, K, flush, world, dist):
    """K timed control steps, CUDA events on the launching stream, L2 flushed between iterations; returns (total_ms, act_ms, kernel_ms)"""
    import torch
    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(K)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for k in range(K):
        flush.fill_(float(k))               # L2 flush between timed iterations (outside the per-step event pairs)
        ev[k][0].record()
        act = policy_fn(obs)
        ev[k][1].record()
        obs, _, _, _ = env.step(act, with_infos=False)
        ev[k][2].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total = sum(a.elapsed_time(c) for a, b, c in ev)
    time_env_steps.last_kernel_ms = [b.elapsed_time(c) for a, b, c in ev]       # per-launch spread (one launch per step)
    return total, sum(a.elapsed_time(b) for a, b, c in ev) / K, sum(b.elapsed_time(c) for a, b, c in ev) / K, obs


def run_ours(args):
    import torch
    import torch.distributed as dist
    from factory_marl_b200 import dist as fdist, ppo as fppo
    from factory_marl_b200.vec_env import FactoryVecEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator is created: keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1); os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        warm = torch.zeros(1, device=dev); dist.all_reduce(warm); torch.cuda.synchronize()
        sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    n = args.envs
    K, W = args.steps, max(args.warmup, 3)
    seeds = fdist.shard_seeds(42, world * n, rank, world)          # env shard of this rank (weak scaling: n envs per rank)

    def make_env(cls=ENV_CLASS, envs=n, arms=2, kw=ENV_KW, sd=None):
        return FactoryVecEnv(cls, num_envs=envs, num_arms=arms, seeds=sd if sd is not None else seeds[:envs], device=dev,
                             envs_per_block=args.envs_per_block, pgs_outer=int(os.environ.get("FM_BENCH_PGS_OUTER", "0")),
                             sort_envs={"0": False, "1": True}.get(os.environ.get("FM_BENCH_SORT_ENVS", ""), None), **kw)

    env = make_env()
    _STATE_ROWS.update(nf=env.core.f.shape[0], nd=env.core.d.shape[0], ni=env.core.i.shape[0])
    # fixed random policy: freshly initialised 178->128->128->16 tanh MLP, orthogonal init, torch seed 0, log_std = 0
    pol = fppo.PolicyValueNet(env.core.obs_dim, env.core.act_dim, discrete=False, device=dev, seed=0)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2
    obs = env.reset()
    obs = desync_preroll(env, pol.act, args.preroll, obs)
    for _ in range(W):
        obs, _, _, _ = env.step(pol.act(obs), with_infos=False)
    torch.cuda.synchronize()
    l0 = env.gpu_launches + pol.launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    step_ms, act_ms, kern_ms, obs = time_env_steps(env, pol.act, obs, K, flush, world, dist)
    kern_each = list(time_env_steps.last_kernel_ms)
    sampler.stop_flag = True
    launches = env.gpu_launches + pol.launches - l0
    total_ms = fdist.max_over_ranks(step_ms, dev)
    value = world * n * K / (total_ms * 1e-3)
    per_rank = None
    if world > 1:          # the scaling loss gets a name: per-rank policy-forward and step-kernel times, gathered to rank 0
        mine = torch.tensor([step_ms / K, act_ms, kern_ms], device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"rank": r, "step_ms": float(x[0]), "act_ms": float(x[1]), "kernel_ms": float(x[2])} for r, x in enumerate(allr)]

    # ---- the round-1 window for comparison across rounds: all envs in the SAME early episode phase (60 synchronised pre-roll steps,
    # control steps 65-85: two cubes in the scene).  The headline above samples every episode phase and is ~15 % lower.
    sync_window = None
    if not args.no_sync_window:
        es = make_env()
        o_s = es.reset()
        for _ in range(60 + W):
            o_s, _, _, _ = es.step(pol.act(o_s), with_infos=False)
        ms_s, _, k_s, o_s = time_env_steps(es, pol.act, o_s, K, flush, world, dist)
        sync_window = {"value": world * n * K / (fdist.max_over_ranks(ms_s, dev) * 1e-3), "unit": UNIT, "kernel_ms": k_s,
                       "note": "round-1 protocol: 60 synchronised pre-roll steps, all envs at control steps 65-85 of their first episode"}
        es.close(); del es
        torch.cuda.empty_cache()

    # ---- e2e: through the public VecEnv API with HOST buffers (pinned): H2D actions and D2H obs/reward/done every step.  The
    # copies of step k run on a second stream under the kernel of step k + 1 (double-buffered pinned buffers, one host sync per step
    # on the copy stream): the host reads every step's result, one step late.  `e2e_sync` is the fully serialised variant.
    Ke = max(3, min(K, 10))
    rng_e = np.random.default_rng(rank)
    host_acts = [torch.from_numpy(np.clip(rng_e.normal(0, 1, (n, 16)), -1, 1).astype(np.float32)).pin_memory() for _ in range(Ke)]
    D = env.core.obs_dim
    host_obs = [torch.empty((n, D), dtype=torch.float32).pin_memory() for _ in range(2)]
    host_rew = [torch.empty((n,), dtype=torch.float32).pin_memory() for _ in range(2)]
    host_done = [torch.empty((n,), dtype=torch.uint8).pin_memory() for _ in range(2)]
    c = env.core
    outs = [dict(obs=c.obs, reward=c.reward, terminated=c.terminated, scores=c.scores, flags=c.flags, terminal_obs=c.terminal_obs, ep_info=c.ep_info)]
    outs.append({k: torch.zeros_like(v) for k, v in outs[0].items()})     # second set of caller-owned output arrays (the C-ABI takes pointers)
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)

    def e2e_loop(pipelined):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        done_ev = [None, None]
        e0.record()
        for k in range(Ke):
            b = k % 2
            for name, t in outs[b if pipelined else 0].items():        # step k writes output set b: set 1 - b is still being read out
                setattr(c, name, t)
            if pipelined and done_ev[b] is not None:
                main.wait_event(done_ev[b])                             # its D2H of step k - 2 has finished
            a_dev = host_acts[k].to(dev, non_blocking=True)
            o2, rew, term, _ = env.step(a_dev, with_infos=False)
            if pipelined:
                ready = torch.cuda.Event(); ready.record(main)
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(ready)
                    host_obs[b].copy_(o2, non_blocking=True); host_rew[b].copy_(rew, non_blocking=True)
                    host_done[b].copy_(c.terminated, non_blocking=True)
                    done_ev[b] = torch.cuda.Event(); done_ev[b].record(copy_stream)
                if k >= 1:
                    done_ev[1 - b].synchronize()                    # the host consumes step k - 1 while step k runs
            else:
                host_obs[b].copy_(o2, non_blocking=True); host_rew[b].copy_(rew, non_blocking=True)
                host_done[b].copy_(c.terminated, non_blocking=True)
                main.synchronize()
        if pipelined:
            done_ev[(Ke - 1) % 2].synchronize()
        e1.record(); torch.cuda.synchronize()
        for name, t in outs[0].items():
            setattr(c, name, t)
        return fdist.max_over_ranks(e0.elapsed_time(e1), dev)

    # the two variants run on consecutive windows of the same episode stream, whose cost drifts by several per cent from window
    # to window: A B B A order, each variant's time summed over its two windows
    t_sync, t_pipe = e2e_loop(False), e2e_loop(True)
    t_pipe += e2e_loop(True); t_sync += e2e_loop(False)
    e2e_sync, e2e_pipe = world * n * 2 * Ke / (t_sync * 1e-3), world * n * 2 * Ke / (t_pipe * 1e-3)
    e2e_value = max(e2e_sync, e2e_pipe)
    h2d = host_acts[0].numel() * 4
    d2h = host_obs[0].numel() * 4 + host_rew[0].numel() * 4 + host_done[0].numel()

    # ---- PPO samples/s at the headline size (4096 envs per GPU; kept for continuity with round 1)
    def ppo_measure(e, n_steps, label):
        algo = fppo.PPO(e, n_steps=n_steps, batch_size=min(8192, n_steps * e.num_envs), n_epochs=10, seed=0)
        algo.collect_rollouts(); algo.train()                       # warm-up iteration (also captures the CUDA graphs)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        p0, p1, p2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        p0.record(); algo.collect_rollouts(); p1.record(); info = algo.train(); p2.record(); torch.cuda.synchronize()
        t_all, t_roll = fdist.max_over_ranks(p0.elapsed_time(p2), dev), fdist.max_over_ranks(p0.elapsed_time(p1), dev)
        samples = world * e.num_envs * n_steps
        chk = float(algo.pol.params.double().sum())
        same = True
        if world > 1:          # driver-visible proof that the gradient all-reduce keeps the replicas identical
            c = torch.tensor([chk], device=dev, dtype=torch.float64); lst = [torch.zeros_like(c) for _ in range(world)]
            dist.all_gather(lst, c); same = all(float(x) == float(lst[0]) for x in lst)
        out = {"workload": label, "samples_per_s": samples / (t_all * 1e-3), "env_steps_per_s_rollout": samples / (t_roll * 1e-3),
               "rollout_ms": t_roll, "update_ms": t_all - t_roll, "samples": samples, "envs_per_gpu": e.num_envs, "n_steps": n_steps,
               "batch_size": algo.batch_size, "n_epochs": 10, "optimizer_steps": info["n_updates"], "cuda_graph": info.get("cuda_graph"),
               "approx_kl": info["approx_kl"], "clip_fraction": info["clip_fraction"], "grad_allreduce": "nccl" if world > 1 else "none",
               "params_identical_on_all_ranks": same, **{k: v for k, v in algo.episode_stats().items() if k.startswith("rollout/")}}
        del algo
        return out

    extra = {}
    if not args.no_ppo:
        extra["ppo_train"] = ppo_measure(env, args.ppo_steps, f"{n} envs per GPU (headline size), n_steps {args.ppo_steps}")
    env.close()
    del env
    torch.cuda.empty_cache()

    # ---- cfg3: 4096 envs in IK-action mode (PauseIKToggleEnv, the learning.py default class), 2 and 4 arms, spawning on
    if not args.no_cfg3:
        cfg3 = {}
        for arms in (2, 4):
            e3 = make_env("PauseIKToggleEnv", n, arms, {})
            p3 = fppo.PolicyValueNet(e3.core.obs_dim, e3.core.act_dim, discrete=True, device=dev, seed=0, capacity=n)
            src = "fresh orthogonal init (seed 0): ~50 % toggles"
            wpath = os.path.join(ROOT, "tests", "golden", "ref_policy_y6lp1j7k.npz")
            if arms == 4 and os.path.exists(wpath):
                p3.set_sb3_state_dict({k: torch.from_numpy(v) for k, v in np.load(wpath).items()})
                src = "the reference's trained policy runs/y6lp1j7k.zip:policy.pth (tests/golden/ref_policy_y6lp1j7k.npz)"
            o3 = e3.reset()
            # episodes under the IK policies end after 30 .. 400 steps (force terminations): a plain pre-roll mixes the phases
            o3 = desync_preroll(e3, p3.act, args.preroll3, o3, stagger=False)
            for _ in range(3):
                o3, _, _, _ = e3.step(p3.act(o3), with_infos=False)
            K3 = max(3, min(K, 10))
            ms, a_ms, k_ms, o3 = time_env_steps(e3, p3.act, o3, K3, flush, world, dist)
            cfg3[f"{arms}_arms"] = {"env_steps_per_s": world * n * K3 / (fdist.max_over_ranks(ms, dev) * 1e-3), "ms_per_step": ms / K3, "policy": src,
                                    "envs_per_gpu": n, "steps": K3, "capacity_flag_envs": float(((e3.core.i[e3.core.off("I_FLAGS")] & 15) != 0).float().mean())}
            e3.close(); del e3, p3
            torch.cuda.empty_cache()
        extra["cfg3"] = {"workload": "4096 batched TaskEnv per GPU in IK-action mode (PauseIKToggleEnv: IK base policy + toggle), conveyor spawning on, de-synchronised", **cfg3}

    # ---- cfg4: PPO with the learning.py hyper-parameters under the declared mapping, 16 384 envs on one GPU
    if not args.no_ppo and not args.no_cfg4:
        e4 = make_env(envs=16384, sd=fdist.shard_seeds(42, world * 16384, rank, world))
        extra["cfg4"] = ppo_measure(e4, 32, "PPO, 16384 envs per GPU, n_steps 32, batch 8192, 10 epochs (SB3 defaults otherwise)")
        e4.close(); del e4
        torch.cuda.empty_cache()

    # ---- cfg5: 65 536 envs sharded over the ranks (strong scaling): env shards per GPU + NCCL gradient all-reduce
    if not args.no_ppo and not args.no_cfg5:
        lo, hi = fdist.env_shard(65536, rank, world)
        e5 = make_env(envs=hi - lo, sd=fdist.shard_seeds(42, 65536, rank, world))
        extra["cfg5"] = {**ppo_measure(e5, 32, f"PPO, 65536 envs over {world} GPU(s) = {hi - lo} per rank, n_steps 32, batch 8192"), "scaling": "strong"}
        e5.close(); del e5
        torch.cuda.empty_cache()

    if rank == 0:
        hbm_peak, peak_kind, sm_max = peaks()
        nf, nd, ni = _STATE_ROWS["nf"], _STATE_ROWS["nd"], _STATE_ROWS["ni"]
        bytes_per_env = 2 * (4 * nf + 8 * nd + 4 * ni) + 4 * 16 + 4 * D + 4 + 1 + 8 + 1 + 16
        ach_hbm = bytes_per_env * n / (kern_ms * 1e-3) / 1e9
        fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
        ach = FLOP_PER_ENV_STEP * n / (kern_ms * 1e-3) / 1e12
        dram_env, dram_src = ncu_dram_bytes_per_env()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{n} batched TaskEnv per GPU ({ENV_CLASS} action map), two arms, fixed random policy (178-128-128-16 tanh MLP, seed 0), 100 sub-steps per step",
                           "envs_per_gpu": n, "l2": "flushed between timed iterations (256 MB write)",
                           "preroll": f"{args.preroll} untimed steps with staggered resets: the envs sample every episode phase",
                           "solver": "warm start from persisted forces; converged islands skipped; one cube vs static/belt: exact active-set Newton (6x6); other islands: PGS <= 4 sweeps (32 with an arm), tol 1e-3 m/s^2; resting hidden cubes asleep; broad-phase queue cached with a 2 cm margin",
                           "block_shape": "14 envs x 12 warps per block, one block per SM"},
                "roofline": {"bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach / fp32_peak,
                             "traffic": dram_env * n if dram_env else None, "traffic_source": dram_src,
                             "kernel": "fm_step_kernel", "kernel_ms": kern_ms, "kernel_ms_min": min(kern_each), "kernel_ms_max": max(kern_each),
                             "kernel_ms_spread": (max(kern_each) - min(kern_each)) / (2.0 * kern_ms), "kernel_ms_each": [round(x, 2) for x in kern_each], "act_ms": act_ms,
                             "flop_per_env_step": FLOP_PER_ENV_STEP, "flop_per_env_step_executed": FLOP_PER_ENV_STEP_EXECUTED,
                             "achieved_executed": FLOP_PER_ENV_STEP_EXECUTED * n / (kern_ms * 1e-3) / 1e12,
                             "frac_executed": FLOP_PER_ENV_STEP_EXECUTED * n / (kern_ms * 1e-3) / 1e12 / fp32_peak,
                             "peak_source": "148 SM x 128 lanes x 2 x max SM clock (non-tensor fp32)",
                             "note": "achieved = SURVEY 8d's algorithmic hand count / kernel time; *_executed = what the kernel really issues (ncu): converged islands skipped, sleeping cubes, 6x6 Newton. The kernel is latency / lock-step bound (DESIGN.md 4)"},
                "roofline_hbm": {"bound": "hbm", "achieved": ach_hbm, "peak": hbm_peak, "unit": "GB/s", "frac": ach_hbm / hbm_peak,
                                 "bytes_per_env_step": bytes_per_env, "peak_source": peak_kind},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "mode": "pipelined" if e2e_pipe >= e2e_sync else "synchronous",
                        "pipelined": e2e_pipe, "synchronous": e2e_sync,
                        "note": "pipelined: D2H of step k runs on a copy stream under the kernel of step k+1 (two sets of output arrays and pinned buffers, one host sync per step); synchronous: the host synchronises the stream after every step; `value` is the faster of the two"},
                "gpu_launches": int(launches), "clocks": sampler.result(), "synchronised_window": sync_window, **extra}
        if per_rank is not None:
            line["per_rank"] = per_rank
        if world == 1 and not args.no_cpu:
            rate, cores, sample, us = cpu_reference_rate(args.cpu_steps)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                    "us_per_mj_step_equivalent": us}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


_STATE_ROWS = {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=N_ENVS_PER_GPU)
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ppo", action="store_true")
    ap.add_argument("--no-cfg3", action="store_true")
    ap.add_argument("--no-cfg4", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true")
    ap.add_argument("--no-sync-window", action="store_true")
    ap.add_argument("--ppo-steps", type=int, default=8)
    ap.add_argument("--preroll", type=int, default=EPISODE_STEPS)
    ap.add_argument("--preroll3", type=int, default=120)
    ap.add_argument("--envs-per-block", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
