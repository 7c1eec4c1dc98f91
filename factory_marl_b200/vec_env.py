"""Batched TaskEnv as an SB3-VecEnv-shaped object on one B200.

Drop-in for the vector env that src/learning.py:98-101 builds (make_vec_env(SubprocVecEnv) over Monitor(EnvClass)):
`num_envs`, `observation_space`/`action_space` descriptions, `reset() -> obs[N,D]`, `step_async/step_wait`, `step`,
`get_attr("ep_score_history")`, `close()`.  Observations, rewards and dones are CUDA tensors (no host round trip).  `step_wait` returns an SB3-style `infos` sequence
that is LAZY: its per-env dicts (scores, out_of_reach, force_terminate, terminal_observation, episode) are only built — one device
to host copy of the step's outputs — when somebody indexes or iterates it, because building Python dicts per env every step is
exactly the pipe traffic the reference pays (learning.py:61-72).  Episode statistics (Monitor, ep_score_history) are kept on
the device by one small kernel per step (`EpisodeMonitor`).
"""
from __future__ import annotations

from typing import List, Optional

import ctypes as C

import numpy as np
import torch

from . import _capi, lib as _lib, model as _model


class TorchCudaBackend:
    def __init__(self, device):
        self.device = torch.device(device)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=_TORCH_DT[np.dtype(dtype)], device=self.device)

    def from_numpy(self, a):
        a = np.ascontiguousarray(a)
        if a.dtype == np.uint64:
            return torch.from_numpy(a.view(np.int64)).to(self.device)
        return torch.from_numpy(a).to(self.device)

    def to_numpy(self, a):
        return a.detach().cpu().numpy()

    def ptr(self, a):
        if a is None:
            return None
        assert a.is_cuda and a.is_contiguous()
        return a.data_ptr()

    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream


_TORCH_DT = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32,
             np.dtype(np.uint8): torch.uint8, np.dtype(np.uint64): torch.int64}


class EpisodeMonitor:
    """stable-baselines3 `Monitor` + the env's `ep_score_history` + what `AdditionalMetricsCallback` computes from it
    (learning.py:53-72,98; environments.py:53,200-201), kept on the device: one launch per env step (fm_monitor_update)."""
    WINDOW = 100

    def __init__(self, num_envs, device):
        self.n, self.dev = num_envs, device
        self.lib = _lib.load()
        self.lib.fm_monitor_update.argtypes = [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 6
        self.hist = torch.zeros((num_envs, self.WINDOW, 2), dtype=torch.int16, device=device)
        self.first = torch.zeros((num_envs, 2), dtype=torch.int16, device=device)
        self.count = torch.zeros(num_envs, dtype=torch.int32, device=device)
        self.ring = torch.zeros((self.WINDOW, 2), dtype=torch.float32, device=device)
        self.ring_head = torch.zeros(1, dtype=torch.int32, device=device)

    def update(self, terminated, ep_info):
        rc = self.lib.fm_monitor_update(terminated.data_ptr(), ep_info.data_ptr(), self.n, self.hist.data_ptr(), self.first.data_ptr(),
                                        self.count.data_ptr(), self.ring.data_ptr(), self.ring_head.data_ptr(),
                                        torch.cuda.current_stream(self.dev).cuda_stream)
        if rc != 0:
            raise RuntimeError("fm_monitor_update failed")

    def ep_score_history(self):
        """per env: the scores [bucket0, bucket1] of its finished episodes, oldest first (the last 100 are kept)"""
        cnt = self.count.cpu().numpy(); hist = self.hist.cpu().numpy()
        out = []
        for e in range(self.n):
            c = int(cnt[e]); k = min(c, self.WINDOW)
            out.append([[int(x) for x in hist[e, (c - k + j) % self.WINDOW]] for j in range(k)])
        return out

    def metrics(self):
        """`rollout/ep_rew_mean`, `rollout/ep_len_mean` (SB3 logger over its ep_info_buffer of the last 100 episodes) and
        `rollout/ep_score_mean_bucket{k}` exactly as learning.py:61-72 computes it: only once every env has finished an episode;
        per env the entries `history[-t] for t in range(min(100, len))` — t = 0 indexes the FIRST episode, so the window is the
        first episode plus the last n - 1 (reference quirk, reproduced)."""
        out = {}
        n_ep = int(self.ring_head.item())
        if n_ep > 0:
            r = self.ring[:min(n_ep, self.WINDOW)]
            out["rollout/ep_rew_mean"] = float(r[:, 0].mean()); out["rollout/ep_len_mean"] = float(r[:, 1].mean())
        out["episodes"] = n_ep
        cnt = self.count
        if bool((cnt > 0).all()):
            W = self.WINDOW
            n = torch.clamp(cnt, max=W).to(torch.int64)                                # entries per env
            j = torch.arange(W, device=self.dev)[None, :]                              # t = 1 .. n-1  ->  history[-t]
            pos = (cnt.to(torch.int64)[:, None] - j) % W                               # ring slot of history[-t], t = j
            valid = (j >= 1) & (j < n[:, None])
            h = torch.gather(self.hist.to(torch.float32), 1, pos[:, :, None].expand(-1, -1, 2))
            tot = (h * valid[:, :, None]).sum(dim=(0, 1)) + self.first.to(torch.float32).sum(dim=0)   # + history[-0] = history[0]
            den = float(valid.sum() + self.n)
            for b in range(2):
                out[f"rollout/ep_score_mean_bucket{b}"] = float(tot[b]) / den
        return out


class LazyInfos:
    """SB3 `infos` (one dict per env) materialised on first access from the step's device outputs."""

    def __init__(self, env):
        self._env, self._step_id, self._list = env, env._step_id, None

    def _build(self):
        if self._list is None:
            env = self._env
            if env._step_id != self._step_id:
                raise RuntimeError("these infos belong to an earlier step: read them before stepping again (they are built lazily "
                                   "from the step's device outputs)")
            c = env.core
            term = c.terminated.cpu().numpy().astype(bool); flags = c.flags.cpu().numpy(); scores = c.scores.cpu().numpy()
            ep = c.ep_info.cpu().numpy(); tobs = c.terminal_obs.clone()
            d = c.d.cpu().numpy()
            conv, play = d[c.off("D_CONV")], d[c.off("D_PLAY_TIME")]
            out = []
            for e in range(len(term)):
                d = {"scores": [int(scores[e, 0]), int(scores[e, 1])], "play_time": float(play[e]), "conveyor_speed": np.array([conv[e]]),
                     "out_of_reach": bool(flags[e] & 1), "force_terminate": bool(flags[e] & 2)}      # base_env.py:274-280
                if term[e]:
                    d["terminal_observation"] = tobs[e]
                    d["episode"] = {"r": float(ep[e, 0]), "l": int(ep[e, 1])}                       # Monitor
                    d["TimeLimit.truncated"] = False
                out.append(d)
            self._list = out
        return self._list

    def __len__(self):
        return self._env.num_envs

    def __getitem__(self, i):
        return self._build()[i]

    def __iter__(self):
        return iter(self._build())


class Box:
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


class MultiDiscrete:
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec)
        self.shape = (len(nvec),)


class FactoryVecEnv:
    """env_class: one of the 7 class names of src/environments.py; env_kwargs = BaseEnv.__init__ kwargs
    (base_env.py:15-35) plus the ProgressRewardEnv weights (environments.py:252-259)."""

    def __init__(self, env_class: str = "PauseIKToggleEnv", num_envs: int = 8, num_arms: int = 2, seed=42, seeds=None,
                 device="cuda:0", auto_reset: bool = True, max_num_objects: int = 10, render_mode=None, sort_envs=None, **env_kwargs):
        if not torch.cuda.is_available():
            raise RuntimeError("FactoryVecEnv needs a CUDA device (B200); there is no CPU path")
        if max_num_objects != 10:
            raise ValueError("max_num_objects is fixed to 10 in this build (base_env.py:30 default)")
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.blob = _model.build_blob(num_arms=num_arms, seed=seed)
        self.core = _capi.FmCore(_lib.load(), TorchCudaBackend(self.device), self.blob, num_envs, num_arms, env_class,
                                 seeds=seeds if seeds is not None else seed, auto_reset=auto_reset, **env_kwargs)
        self.num_envs, self.num_arms, self.dof = num_envs, num_arms, 8
        self.env_class = env_class
        D = self.core.obs_dim
        self.observation_space = Box(-np.inf, np.inf, (D,))
        if self.core.act_is_int:
            self.action_space = MultiDiscrete([2] * num_arms)
        else:
            self.action_space = Box(-1.0, 1.0, (self.core.act_dim,))
        self.monitor = EpisodeMonitor(num_envs, self.device)
        self._actions = None
        self._step_id = 0
        # order the envs by expected cost before every launch (fm_set_env_order)
        # Grouping pays when a launch is at least ~two waves of blocks (the long blocks start first and the short ones fill in behind
        # them: + 9 % at 4096 two-arm envs); with a single wave the block of the most expensive envs IS the launch time (- 12 %
        # at 2072 envs), so the default switches on from 1.5 waves
        per_wave = torch.cuda.get_device_properties(self.device).multi_processor_count * {2: 14, 4: 7, 6: 6}[num_arms]
        self.sort_envs = (num_envs >= 1.5 * per_wave) if sort_envs is None else bool(sort_envs)
        self._armcon_row, self._ncon_row, self._nin_row = self.core.off("I_ARMCON"), self.core.off("I_PREV_N"), self.core.off("I_NIN")
        self._since_row = self.core.off("I_SINCE")

    # ---- VecEnv API
    def reset(self):
        return self.core.reset()

    def step_async(self, actions):
        a = torch.as_tensor(actions, device=self.device)
        a = a.to(torch.int32) if self.core.act_is_int else a.to(torch.float32)
        want = (self.num_envs, self.core.act_dim) if self.core.act_dim > 0 else None
        if want is not None and tuple(a.shape) != want:
            raise ValueError(f"actions must have shape {want}, got {tuple(a.shape)}")      # base_env.py:252 asserts the ctrl dimension
        self._actions = a.contiguous()

    def step_wait(self, with_infos: bool = True):
        """-> (obs[N,D], reward[N], dones[N], infos).  `infos` is lazy (see LazyInfos); with_infos=False returns None instead."""
        c = self.core
        c.env_step(self._actions)
        self._step_id += 1
        self.monitor.update(c.terminated, c.ep_info)
        if self.sort_envs:
            # a block lasts as long as its most expensive env and the launch as long as the slowest SM's two blocks: order the envs by
            # expected cost (arm in contact, cubes in the scene, contacts), most expensive first, so that like envs share a block and
            # the long blocks start first (measured + 9 % env-steps/s at 4096 envs; results do not depend on the order)
            key = c.i[self._armcon_row] * 4096 + c.i[self._nin_row] * 64 + c.i[self._ncon_row] + (c.i[self._since_row] < 8).int() * 48   # + a cube in flight
            c.set_env_order(torch.argsort(key, descending=True, stable=True).to(torch.int32))
        return c.obs, c.reward, c.terminated.bool(), (LazyInfos(self) if with_infos else None)

    def step(self, actions, with_infos: bool = True):
        self.step_async(actions)
        return self.step_wait(with_infos)

    @property
    def ep_score_history(self):
        return self.monitor.ep_score_history()

    def get_attr(self, name: str):
        if name == "ep_score_history":
            return self.monitor.ep_score_history()
        return [getattr(self, name)] * self.num_envs

    def close(self):
        self.core.close()

    @property
    def gpu_launches(self) -> int:
        return int(_lib.load().fm_launch_count(self.core.h))


def dump_trajectory(env: FactoryVecEnv, policy, n_steps: int, path: str, env_index: int = 0):
    """Export one env's rollout for OFFLINE rendering: `qpos` [T+1, nq] and `qvel` [T+1, nv] per control step in the reference's
    MuJoCo layout (belt slide, 10 x free joint (xyz, wxyz), arms 7 hinges + 2 plates), `ctrl` [T+1, nu], rewards,
    dones and scores.  On a machine that has MuJoCo, `physics.data.qpos[:] = qpos[t]; physics.forward()` followed by the
    reference's `rendering.py` replays it (challenge_env/rendering.py is out of scope here: no GPU renderer).
    `policy(obs) -> actions` (e.g. PolicyValueNet.act); None = zero / pause actions."""
    c = env.core
    obs = env.reset()
    qs, vs, cs, rs, ds, ss = [], [], [], [], [], []
    q, v, ct = c.get_state(env_index); qs.append(q); vs.append(v); cs.append(ct)
    for t in range(n_steps):
        if policy is None:
            a = torch.zeros((env.num_envs, max(c.act_dim, 1)), dtype=torch.int32 if c.act_is_int else torch.float32, device=env.device)
        else:
            a = policy(obs)
        obs, rew, done, _ = env.step(a, with_infos=False)
        q, v, ct = c.get_state(env_index); qs.append(q); vs.append(v); cs.append(ct)
        rs.append(float(rew[env_index])); ds.append(bool(done[env_index])); ss.append(c.scores[env_index].cpu().numpy())
    np.savez_compressed(path, qpos=np.array(qs), qvel=np.array(vs), ctrl=np.array(cs), reward=np.array(rs), done=np.array(ds),
                        scores=np.array(ss), env_class=env.env_class, num_arms=env.num_arms, dt=0.1,
                        layout="qpos: [belt, 10 x (x y z qw qx qy qz), arms x (7 hinges, left plate, right plate)]")
    return path


class FactoryEnv:
    """One env with the reference's per-env gymnasium contract (src/environments.py:151-248): `reset(seed=None, options=None) ->
    (obs, {})`, `step(action) -> (obs, reward, terminated, False, info)`; numpy in / out.  A convenience view of a one-env
    FactoryVecEnv without auto-reset — the batched class is the product."""

    def __init__(self, env_class: str = "PauseIKToggleEnv", num_arms: int = 2, seed=42, **env_kwargs):
        self.vec = FactoryVecEnv(env_class, num_envs=1, num_arms=num_arms, seed=seed, auto_reset=False, **env_kwargs)
        self.observation_space, self.action_space = self.vec.observation_space, self.vec.action_space
        self.num_arms, self.dof = num_arms, 8

    @property
    def ep_score_history(self):
        return self.vec.monitor.ep_score_history()[0]

    def reset(self, *, seed=None, options=None):
        return self.vec.reset()[0].cpu().numpy(), {}          # gym seeding does not touch the TaskManager RNG

    def step(self, action):
        a = np.asarray(action).reshape(1, -1)
        obs, rew, done, infos = self.vec.step(torch.from_numpy(a))
        info = dict(infos[0])
        info.pop("terminal_observation", None)
        return obs[0].cpu().numpy(), float(rew[0]), bool(done[0]), False, info

    def close(self):
        self.vec.close()
