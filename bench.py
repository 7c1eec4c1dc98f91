#!/usr/bin/env python
"""bench.py — headline benchmark of the batched TaskEnv hot path (BASELINE.json: env-steps/s; PPO samples/s).

  python bench.py --gpus N --steps K --warmup W            ours: CUDA step kernel (+ on-GPU policy), one rank per GPU
  python bench.py --impl reference --gpus N ...            reference arm: the fp64 CPU restatement (oracle port) on host cores

A "step" is one control step (100 physics sub-steps + task logic + obs/reward) of every env of the workload
`4096 batched TaskEnv (AllFullRLProgressRewardEnv map), two arms, fixed random policy` (BASELINE.json configs[1])
per GPU (weak scaling: 4096 envs per rank, env shards independent, no data-path collective).
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
# algorithmic work per env-step (DESIGN.md "Measurement"): state in+out + action + obs + scalars, fp32 flop hand count
FLOP_PER_ENV_STEP = 2.1e7            # SURVEY.md 8d hand count of the path as MuJoCo runs it (25 solver iterations over ~170 rows)
FLOP_PER_ENV_STEP_EXECUTED = 4.76e6   # fp32 flop this kernel executes per env-step in the bench workload: ncu fadd + fmul + 2 ffma (profiles/r1m, r1n)
# dram__bytes_read.sum + dram__bytes_write.sum of fm_step_kernel per env, from the ncu --set full capture
# profiles/r1o_E14_step_kernel_ncu_raw.csv (8.34 MB read + 0.41 MB written per launch of 2072 envs = one wave of the bench's
# block shape); the 4096-env launch is two such waves
DRAM_BYTES_PER_ENV_NCU = (8.341e6 + 0.413e6) / 2072


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured", d.get("sm_max_mhz", 1965.0)
    return 6650.0, "fallback", 1965.0


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
            time.sleep(0.15)

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
    acts = rng.uniform(-1, 1, (warm + n_steps, 16)).astype(np.float32)
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
    """The fp64 CPU restatement of the reference path (oracle/), one env per host core, all cores, wall clock.
    The real reference (Python + dm_control + MuJoCo + SubprocVecEnv pipes) cannot run in this image;
    this port has none of that overhead."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        times = pool.map(_cpu_worker, [(i, steps_per_env, warm) for i in range(cores)])
    wall = time.perf_counter() - t0
    rate = cores * steps_per_env / max(times)
    return rate, cores, f"{cores} envs x {steps_per_env} control steps (100 sub-steps each) of the bench workload, one env per core, seeded random actions", wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = 4           # control steps per env per "step" of this arm (bounded sample)
    K, W = args.steps, args.warmup
    rate, cores, sample, wall = cpu_reference_rate(per_step * K, warm=max(1, W))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * cores * per_step / rate, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{ENV_CLASS}, 2 arms, seeded random actions, fp64 CPU restatement of MuJoCo TaskEnv (the reference's MuJoCo/dm_control/SB3 stack is not installable here)",
                       "envs": cores, "steps_per_env_per_step": per_step},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from factory_marl_b200.vec_env import FactoryVecEnv
    from factory_marl_b200 import ppo as fppo

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
    env = FactoryVecEnv(ENV_CLASS, num_envs=n, num_arms=2, seeds=list(range(42 + rank * n, 42 + (rank + 1) * n)),
                        device=dev, envs_per_block=args.envs_per_block, **ENV_KW)
    # fixed random policy: freshly initialised 178->128->128->16 tanh MLP, orthogonal init, torch seed 0, log_std = 0
    pol = fppo.PolicyValueNet(env.core.obs_dim, env.core.act_dim, discrete=False, device=dev, seed=0)
    obs = env.reset()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def one_step(o):
        act = pol.act(o)                    # on-GPU MLP forward + Gaussian sample, clipped to [-1, 1] as SB3 does
        o2, rew, term, _ = env.step(act)
        return o2

    # untimed pre-roll into the steady state of an episode (first cube riding the belt, hidden cubes settled), then W warm-up steps
    for _ in range(args.preroll + W):
        obs = one_step(obs)
    torch.cuda.synchronize()
    l0 = env.gpu_launches + pol.launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for k in range(K):
        flush.fill_(float(k))               # L2 flush between timed iterations (outside the per-step event pairs)
        ev[k][0].record()
        act = pol.act(obs)
        ev[k][1].record()
        obs, rew, term, _ = env.step(act)
        ev[k][2].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    step_ms = sum(a.elapsed_time(c) for a, b, c in ev)
    kern_ms = sum(b.elapsed_time(c) for a, b, c in ev) / K      # the dominant kernel: fm_step_kernel
    launches = env.gpu_launches + pol.launches - l0
    t = torch.tensor([step_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * n * K / (total_ms * 1e-3)

    # ---- e2e: through the public VecEnv API with HOST buffers (pinned), H2D actions + D2H obs/reward/done every step.
    # Measured on a second, identically seeded env brought to the same episode phase by the same untimed pre-roll, so that
    # `value` and `e2e` see the same contact load (throughput drifts by ~10 % along an episode).
    env.close()
    env = FactoryVecEnv(ENV_CLASS, num_envs=n, num_arms=2, seeds=list(range(42 + rank * n, 42 + (rank + 1) * n)),
                        device=dev, envs_per_block=args.envs_per_block, **ENV_KW)
    obs = env.reset()
    for _ in range(args.preroll + W):
        obs = one_step(obs)
    torch.cuda.synchronize()
    Ke = max(3, min(K, 10))
    # same action statistics as the fixed random policy of the device-resident arm: a fresh N(~0, 1) draw clipped to [-1, 1]
    # every control step (the host-side user computes its actions before the call; the draw itself is not timed)
    rng_e = np.random.default_rng(rank)
    host_acts = [torch.from_numpy(np.clip(rng_e.normal(0, 1, (n, 16)), -1, 1).astype(np.float32)).pin_memory() for _ in range(Ke)]
    host_act = host_acts[0]
    host_obs = torch.empty((n, env.core.obs_dim), dtype=torch.float32).pin_memory()
    host_rew = torch.empty((n,), dtype=torch.float32).pin_memory()
    host_done = torch.empty((n,), dtype=torch.bool).pin_memory()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(Ke):
        a_dev = host_acts[k].to(dev, non_blocking=True)
        o2, rew, term, _ = env.step(a_dev)
        host_obs.copy_(o2, non_blocking=True); host_rew.copy_(rew, non_blocking=True); host_done.copy_(term, non_blocking=True)
        torch.cuda.current_stream().synchronize()       # the user reads the result on the host every step
    e1.record(); torch.cuda.synchronize()
    te = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n * Ke / (float(te.item()) * 1e-3)
    h2d = host_act.numel() * 4
    d2h = host_obs.numel() * 4 + host_rew.numel() * 4 + host_done.numel()

    # ---- secondary metric (BASELINE.json: "PPO train samples/sec"): one PPO iteration = rollout of n_steps + 10 epochs of updates,
    # SB3 hyper-parameters under the declared large-batch mapping; gradients all-reduced over NCCL when world > 1
    ppo_info = None
    if not args.no_ppo:
        algo = fppo.PPO(env, n_steps=args.ppo_steps, batch_size=min(8192, args.ppo_steps * n), n_epochs=10, seed=0)
        algo.collect_rollouts(); algo.train()                       # warm-up iteration
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        p0, p1, p2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        p0.record(); algo.collect_rollouts(); p1.record(); info = algo.train(); p2.record(); torch.cuda.synchronize()
        tp = torch.tensor([p0.elapsed_time(p2), p0.elapsed_time(p1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        samples = world * n * args.ppo_steps
        ppo_info = {"samples_per_s": samples / (float(tp[0]) * 1e-3), "rollout_ms": float(tp[1]), "update_ms": float(tp[0] - tp[1]),
                    "samples": samples, "n_steps": args.ppo_steps, "batch_size": algo.batch_size, "n_epochs": 10,
                    "optimizer_steps": info["n_updates"], "cuda_graph": info.get("cuda_graph"), "approx_kl": info["approx_kl"], "clip_fraction": info["clip_fraction"],
                    "grad_allreduce": "nccl" if world > 1 else "none"}
    if rank == 0:
        hbm_peak, peak_kind, sm_max = peaks()
        nf, nd, ni = env.core.f.shape[0], env.core.d.shape[0], env.core.i.shape[0]
        bytes_per_env = 2 * (4 * nf + 8 * nd + 4 * ni) + 4 * 16 + 4 * env.core.obs_dim + 4 + 1 + 8 + 1 + 16
        ach = bytes_per_env * n / (kern_ms * 1e-3) / 1e9
        fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{n} batched TaskEnv per GPU ({ENV_CLASS} action map), two arms, fixed random policy (178-128-128-16 tanh MLP, seed 0), 100 sub-steps per step",
                           "envs_per_gpu": n, "l2": "flushed between timed iterations (256 MB write)", "preroll_steps": args.preroll,
                           "solver": "warm start from persisted forces; islands already converged are skipped; one cube vs static/belt: exact active-set Newton (6x6); other islands: PGS <= 4 sweeps (32 with an arm), tol 1e-3 m/s^2; resting hidden cubes asleep", "block_shape": "14 envs x 12 warps per block, one block per SM"},
                "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": DRAM_BYTES_PER_ENV_NCU * n,
                             "peak_source": peak_kind, "kernel": "fm_step_kernel", "kernel_ms": kern_ms, "bytes_per_env_step": bytes_per_env,
                             "note": "the step kernel is FP32-issue/latency bound, not HBM bound (SURVEY.md 8d); see roofline_fp32"},
                "roofline_fp32": {"achieved": FLOP_PER_ENV_STEP * n / (kern_ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                                  "frac": FLOP_PER_ENV_STEP * n / (kern_ms * 1e-3) / 1e12 / fp32_peak, "flop_per_env_step": FLOP_PER_ENV_STEP,
                                  "flop_per_env_step_executed": FLOP_PER_ENV_STEP_EXECUTED,
                                  "achieved_executed": FLOP_PER_ENV_STEP_EXECUTED * n / (kern_ms * 1e-3) / 1e12,
                                  "frac_executed": FLOP_PER_ENV_STEP_EXECUTED * n / (kern_ms * 1e-3) / 1e12 / fp32_peak,
                                  "note": "achieved = SURVEY's algorithmic hand count / kernel time; *_executed = flop the kernel really issues (ncu): warm-started islands that are already converged, sleeping cubes and the 6x6 Newton solve do 4.4x less arithmetic than the hand count assumes",
                                  "peak_source": "148 SM x 128 lanes x 2 x max SM clock"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches), "clocks": sampler.result()}
        if ppo_info is not None:
            line["ppo_train"] = ppo_info
        if world == 1 and not args.no_cpu:
            rate, cores, sample, wall = cpu_reference_rate(args.cpu_steps)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--ppo-steps", type=int, default=8)
    ap.add_argument("--preroll", type=int, default=60)
    ap.add_argument("--envs-per-block", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
