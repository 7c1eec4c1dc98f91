"""Batched TaskEnv as an SB3-VecEnv-shaped object on one B200.

Drop-in for the vector env that src/learning.py:98-101 builds (make_vec_env(SubprocVecEnv) over Monitor(EnvClass)):
`num_envs`, `observation_space`/`action_space` descriptions, `reset() -> obs[N,D]`, `step_async/step_wait`, `step`,
`get_attr("ep_score_history")`, `close()`.  Observations, rewards and dones are CUDA tensors (no host round trip);
`step_wait` returns SB3-style `infos` only when asked (`with_infos=True`), because building Python dicts per env
is exactly the per-step pipe traffic the reference pays (learning.py:61-72).
"""
from __future__ import annotations

from collections import deque
from typing import List, Optional

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
                 device="cuda:0", auto_reset: bool = True, max_num_objects: int = 10, render_mode=None, **env_kwargs):
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
        self.ep_score_history: List[deque] = [deque(maxlen=100) for _ in range(num_envs)]
        self._actions = None

    # ---- VecEnv API
    def reset(self):
        return self.core.reset()

    def step_async(self, actions):
        a = torch.as_tensor(actions, device=self.device)
        a = a.to(torch.int32) if self.core.act_is_int else a.to(torch.float32)
        self._actions = a.contiguous()

    def step_wait(self, with_infos: bool = False):
        c = self.core
        c.env_step(self._actions)
        infos = None
        if with_infos:
            infos = self._infos()
        return c.obs, c.reward, c.terminated.bool(), infos

    def step(self, actions, with_infos: bool = False):
        self.step_async(actions)
        return self.step_wait(with_infos)

    def _infos(self):
        c = self.core
        term = c.terminated.cpu().numpy().astype(bool)
        flags = c.flags.cpu().numpy()
        scores = c.scores.cpu().numpy()
        ep = c.ep_info.cpu().numpy()
        tobs = c.terminal_obs
        infos = []
        for e in range(self.num_envs):
            d = {"scores": [int(scores[e, 0]), int(scores[e, 1])], "out_of_reach": bool(flags[e] & 1),
                 "force_terminate": bool(flags[e] & 2)}
            if term[e]:
                d["terminal_observation"] = tobs[e]
                d["episode"] = {"r": float(ep[e, 0]), "l": int(ep[e, 1])}
                self.ep_score_history[e].append([int(ep[e, 2]), int(ep[e, 3])])
            infos.append(d)
        return infos

    def get_attr(self, name: str):
        if name == "ep_score_history":
            return [list(h) for h in self.ep_score_history]
        return [getattr(self, name)] * self.num_envs

    def close(self):
        self.core.close()

    @property
    def gpu_launches(self) -> int:
        return int(_lib.load().fm_launch_count(self.core.h))
