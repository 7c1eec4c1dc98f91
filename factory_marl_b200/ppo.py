"""On-GPU PPO around the batched TaskEnv: host-side mirror of what src/learning.py:106-118 asks of stable-baselines3.

`PolicyValueNet`  — SB3 ActorCriticPolicy("MlpPolicy", net_arch=[128,128]) parameters in one flat fp32 buffer (+ bf16 copy for the
                    tensor-core GEMMs), orthogonal init as SB3 does, loader/saver for the reference's runs/*.zip `policy.pth`.
`PPO`             — collect_rollouts / GAE / train loop with SB3 2.3.2 default hyper-parameters (runs/r666unuv.json), the declared
                    large-batch mapping (n_steps, batch_size configurable), and the NCCL gradient all-reduce for env shards.
All arithmetic runs in csrc/fm_ppo.cu through the C-ABI; torch provides device memory, streams and torch.distributed only.
"""
from __future__ import annotations

import ctypes as C
import io
import math
import zipfile
from typing import Optional

import numpy as np
import torch

from . import dist as _dist, lib as _lib


class fm_ppo_net(C.Structure):
    _fields_ = [("obs_dim", C.c_int32), ("obs_pad", C.c_int32), ("act_dim", C.c_int32), ("out_pad", C.c_int32), ("discrete", C.c_int32),
                ("n_params", C.c_int32), ("r0", C.c_int32), ("r1", C.c_int32),
                ("w1", C.c_int32 * 2), ("b1", C.c_int32 * 2), ("w2", C.c_int32 * 2), ("b2", C.c_int32 * 2), ("w3", C.c_int32 * 2),
                ("b3", C.c_int32 * 2), ("log_std", C.c_int32), ("r2", C.c_int32),
                ("params", C.c_void_p), ("params_bf16", C.c_void_p), ("grads", C.c_void_p), ("adam_m", C.c_void_p), ("adam_v", C.c_void_p),
                ("params_pk", C.c_void_p)]


class fm_ppo_ws(C.Structure):
    _fields_ = [("cap", C.c_int32), ("r", C.c_int32), ("x", C.c_void_p), ("h1", C.c_void_p * 2), ("h2", C.c_void_p * 2),
                ("out", C.c_void_p * 2), ("dout", C.c_void_p * 2), ("dh2", C.c_void_p * 2), ("dh1", C.c_void_p * 2), ("scal", C.c_void_p)]


def _bind(lib):
    if getattr(lib, "_ppo_bound", False):
        return lib
    vp = C.c_void_p
    lib.fm_ppo_last_error.restype = C.c_char_p
    lib.fm_policy_forward.argtypes = [C.POINTER(fm_ppo_net), C.POINTER(fm_ppo_ws), vp, C.c_int, C.c_uint64, C.c_int, vp, vp, vp, vp, vp]
    lib.fm_gae.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, vp, vp, vp]
    lib.fm_ppo_update.argtypes = [C.POINTER(fm_ppo_net), C.POINTER(fm_ppo_ws), vp, vp, vp, vp, vp, vp, C.c_int, C.c_float, C.c_float,
                                  C.c_float, C.c_int, vp, vp]
    lib.fm_adam_step.argtypes = [C.POINTER(fm_ppo_net), C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, vp, vp]
    lib.fm_ppo_sync_bf16.argtypes = [C.POINTER(fm_ppo_net), vp]
    lib._ppo_bound = True
    return lib


def _ortho(rng: torch.Generator, rows: int, cols: int, gain: float) -> torch.Tensor:
    """torch.nn.init.orthogonal_ restated for a weight of shape [rows(out), cols(in)] (SB3 ActorCriticPolicy.init_weights [3P])."""
    flat = torch.randn((rows, cols), generator=rng)
    if rows < cols:
        flat = flat.t()
    q, r = torch.linalg.qr(flat)
    q = q * torch.sign(torch.diagonal(r))
    if rows < cols:
        q = q.t()
    return gain * q


class PolicyValueNet:
    HID = 128

    def __init__(self, obs_dim: int, act_dim: int, discrete: bool, device="cuda:0", seed: int = 0, capacity: int = 4096):
        self.lib = _bind(_lib.load())
        self.dev = torch.device(device)
        self.obs_dim, self.act_dim, self.discrete = obs_dim, act_dim, bool(discrete)
        self.obs_pad = (obs_dim + 63) // 64 * 64
        n_out = 2 * act_dim if discrete else act_dim
        self.n_out = n_out
        self.out_pad = (n_out + 15) // 16 * 16
        H = self.HID
        off, o = {}, 0
        for which, op in ((0, self.out_pad), (1, 16)):
            for name, sz in (("w1", self.obs_pad * H), ("b1", H), ("w2", H * H), ("b2", H), ("w3", H * op), ("b3", op)):
                off[(name, which)] = o
                o += sz
        off["log_std"] = o
        o += max(32, (act_dim + 15) // 16 * 16)      # 32 for two and four arms (the layout of round-1 checkpoints), 48 for six
        self.off, self.n_params = off, o
        self.params = torch.zeros(o, dtype=torch.float32, device=self.dev)
        self.params_bf16 = torch.zeros(o, dtype=torch.bfloat16, device=self.dev)
        self.params_pk = torch.zeros(o, dtype=torch.bfloat16, device=self.dev)   # weights in the tcgen05 B-operand image (fm_ppo.cu pack_weights_kernel)
        self.grads = torch.zeros(o, dtype=torch.float32, device=self.dev)
        self.adam_m = torch.zeros(o, dtype=torch.float32, device=self.dev)
        self.adam_v = torch.zeros(o, dtype=torch.float32, device=self.dev)
        self.net = fm_ppo_net(obs_dim=obs_dim, obs_pad=self.obs_pad, act_dim=act_dim, out_pad=self.out_pad, discrete=int(discrete), n_params=o,
                              w1=(C.c_int32 * 2)(off[("w1", 0)], off[("w1", 1)]), b1=(C.c_int32 * 2)(off[("b1", 0)], off[("b1", 1)]),
                              w2=(C.c_int32 * 2)(off[("w2", 0)], off[("w2", 1)]), b2=(C.c_int32 * 2)(off[("b2", 0)], off[("b2", 1)]),
                              w3=(C.c_int32 * 2)(off[("w3", 0)], off[("w3", 1)]), b3=(C.c_int32 * 2)(off[("b3", 0)], off[("b3", 1)]),
                              log_std=off["log_std"], params=self.params.data_ptr(), params_bf16=self.params_bf16.data_ptr(),
                              grads=self.grads.data_ptr(), adam_m=self.adam_m.data_ptr(), adam_v=self.adam_v.data_ptr(),
                              params_pk=self.params_pk.data_ptr())
        self._ws = None
        self._alloc_ws(capacity)
        self.launches = 0
        self._call = 0
        self.seed = seed
        self.init_weights(seed)
        # outputs of act()
        self.adam_step = 0

    # ---- parameters
    def _alloc_ws(self, cap):
        cap = (cap + 63) // 64 * 64
        self.ws_generation = getattr(self, "ws_generation", 0) + 1   # CUDA graphs captured over the old buffers are stale (PPO checks)
        d, H = self.dev, self.HID
        z = lambda *s, dt=torch.bfloat16: torch.zeros(s, dtype=dt, device=d)
        self._bufs = dict(x=z(cap, self.obs_pad), h1=[z(cap, H), z(cap, H)], h2=[z(cap, H), z(cap, H)],
                          out=[z(cap, self.out_pad, dt=torch.float32), z(cap, 16, dt=torch.float32)],
                          dout=[z(cap, self.out_pad), z(cap, 16)], dh2=[z(cap, H), z(cap, H)], dh1=[z(cap, H), z(cap, H)],
                          scal=z(16, dt=torch.float32))
        b = self._bufs
        p2 = lambda l: (C.c_void_p * 2)(l[0].data_ptr(), l[1].data_ptr())
        self._ws = fm_ppo_ws(cap=cap, x=b["x"].data_ptr(), h1=p2(b["h1"]), h2=p2(b["h2"]), out=p2(b["out"]), dout=p2(b["dout"]),
                             dh2=p2(b["dh2"]), dh1=p2(b["dh1"]), scal=b["scal"].data_ptr())
        self.capacity = cap
        self.act_env = torch.zeros((cap, self.act_dim), dtype=torch.int32 if self.discrete else torch.float32, device=d)
        self.act_raw = torch.zeros((cap, self.act_dim), dtype=torch.float32, device=d)
        self.logp = torch.zeros(cap, dtype=torch.float32, device=d)
        self.value = torch.zeros(cap, dtype=torch.float32, device=d)

    def _view(self, name, which, rows, cols):
        o = self.off[(name, which)]
        return self.params[o:o + rows * cols].view(rows, cols)

    def set_sb3_state_dict(self, sd, into=None):
        """Load an SB3 `policy.pth` state-dict (runs/*.zip): Linear weights are [out, in]; ours are [in(padded), out(padded)]."""
        H, D = self.HID, self.obs_dim
        buf = self.params if into is None else into          # `into`: the Adam moment buffers share the parameter layout
        saved, self.params = self.params, buf
        try:
            self._set_sb3(sd, H, D)
        finally:
            self.params = saved
        if into is None:
            self.sync_bf16()

    def _set_sb3(self, sd, H, D):
        self.params.zero_()
        for which, pre in ((0, "mlp_extractor.policy_net"), (1, "mlp_extractor.value_net")):
            self._view("w1", which, self.obs_pad, H)[:D].copy_(sd[f"{pre}.0.weight"].t())
            self._view("b1", which, 1, H)[0].copy_(sd[f"{pre}.0.bias"])
            self._view("w2", which, H, H).copy_(sd[f"{pre}.2.weight"].t())
            self._view("b2", which, 1, H)[0].copy_(sd[f"{pre}.2.bias"])
        self._view("w3", 0, H, self.out_pad)[:, :self.n_out].copy_(sd["action_net.weight"].t())
        self._view("b3", 0, 1, self.out_pad)[0, :self.n_out].copy_(sd["action_net.bias"])
        self._view("w3", 1, H, 16)[:, :1].copy_(sd["value_net.weight"].t())
        self._view("b3", 1, 1, 16)[0, :1].copy_(sd["value_net.bias"])
        if not self.discrete:
            self.params[self.off["log_std"]:self.off["log_std"] + self.act_dim].copy_(sd["log_std"])

    def sb3_state_dict(self, src=None):
        H, D = self.HID, self.obs_dim
        if src is not None:                                   # same layout, another buffer (Adam moments)
            saved, self.params = self.params, src
            try:
                return self.sb3_state_dict()
            finally:
                self.params = saved
        sd = {}
        if not self.discrete:
            sd["log_std"] = self.params[self.off["log_std"]:self.off["log_std"] + self.act_dim].clone().cpu()
        for which, pre in ((0, "mlp_extractor.policy_net"), (1, "mlp_extractor.value_net")):
            sd[f"{pre}.0.weight"] = self._view("w1", which, self.obs_pad, H)[:D].t().contiguous().cpu()
            sd[f"{pre}.0.bias"] = self._view("b1", which, 1, H)[0].clone().cpu()
            sd[f"{pre}.2.weight"] = self._view("w2", which, H, H).t().contiguous().cpu()
            sd[f"{pre}.2.bias"] = self._view("b2", which, 1, H)[0].clone().cpu()
        sd["action_net.weight"] = self._view("w3", 0, H, self.out_pad)[:, :self.n_out].t().contiguous().cpu()
        sd["action_net.bias"] = self._view("b3", 0, 1, self.out_pad)[0, :self.n_out].clone().cpu()
        sd["value_net.weight"] = self._view("w3", 1, H, 16)[:, :1].t().contiguous().cpu()
        sd["value_net.bias"] = self._view("b3", 1, 1, 16)[0, :1].clone().cpu()
        return sd

    def load_sb3_zip(self, path):
        """visualisation.py:53 equivalent for the weights: read `policy.pth` out of an SB3 model zip (runs/*.zip)."""
        with zipfile.ZipFile(path) as z:
            sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
        self.set_sb3_state_dict(sd)

    def init_weights(self, seed=0):
        """SB3 ortho_init: gain sqrt(2) for the extractors, 0.01 for action_net, 1 for value_net, zero biases, log_std = 0 [3P]."""
        g = torch.Generator().manual_seed(seed)
        H, D = self.HID, self.obs_dim
        sd = {}
        for pre in ("mlp_extractor.policy_net", "mlp_extractor.value_net"):
            sd[f"{pre}.0.weight"] = _ortho(g, H, D, math.sqrt(2)); sd[f"{pre}.0.bias"] = torch.zeros(H)
            sd[f"{pre}.2.weight"] = _ortho(g, H, H, math.sqrt(2)); sd[f"{pre}.2.bias"] = torch.zeros(H)
        sd["action_net.weight"] = _ortho(g, self.n_out, H, 0.01); sd["action_net.bias"] = torch.zeros(self.n_out)
        sd["value_net.weight"] = _ortho(g, 1, H, 1.0); sd["value_net.bias"] = torch.zeros(1)
        sd["log_std"] = torch.zeros(self.act_dim)
        self.set_sb3_state_dict({k: v.to(self.dev) for k, v in sd.items()})

    def sync_bf16(self):
        self._chk(self.lib.fm_ppo_sync_bf16(C.byref(self.net), self._stream()))

    # ---- calls
    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("libfm ppo: " + self.lib.fm_ppo_last_error().decode())

    def forward(self, obs: torch.Tensor, deterministic=False, with_value=True):
        n = obs.shape[0]
        if n > self.capacity:
            self._alloc_ws(n)
        assert obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous() and n % 64 == 0
        self._call += 1
        seed = (self.seed * 0x9E3779B1 + self._call * 0x85EBCA77) & 0xFFFFFFFFFFFFFFFF
        self._chk(self.lib.fm_policy_forward(C.byref(self.net), C.byref(self._ws), obs.data_ptr(), n, seed, int(deterministic),
                                             self.act_env.data_ptr(), self.act_raw.data_ptr(), self.logp.data_ptr(),
                                             self.value.data_ptr() if with_value else None, self._stream()))
        self.launches += 9 if with_value else 5
        return self.act_env[:n], self.act_raw[:n], self.logp[:n], self.value[:n]

    def check_status(self):
        """the tcgen05 rollout forward reports an mbarrier time-out in ws->scal[8] (fm_mlp_tc5.cu); read it once per rollout and
        fail loudly instead of training on garbage TMEM results"""
        st = float(self._bufs["scal"][8])
        if st != 0.0:
            self._bufs["scal"][8] = 0.0
            raise RuntimeError(f"tcgen05 MLP forward reported status {st} (mbarrier time-out): results of this rollout are invalid; "
                               "fm_ppo_use_tcgen05(0) selects the mma.sync path")

    def act(self, obs: torch.Tensor, deterministic=False):
        return self.forward(obs, deterministic, with_value=False)[0]


class PPO:
    """PPO with the reference's hyper-parameters (SB3 2.3.2 defaults recorded in runs/r666unuv.json:11-73, gamma learning.py:47).
    Declared mapping for large batches (SURVEY.md hard part 6): n_steps and batch_size are free parameters; everything else
    (gamma .99, lambda .95, clip .2, 10 epochs, lr 3e-4, Adam eps 1e-5, vf .5, ent 0, max-grad-norm .5, per-minibatch
    advantage normalisation, separate 128-128 tanh nets) is kept.  With env shards on several GPUs the flat gradient is
    summed with one NCCL all-reduce per optimiser step and divided by the world size (advantage normalisation stays rank-local)."""

    def __init__(self, env, n_steps=32, batch_size=8192, n_epochs=10, gamma=0.99, gae_lambda=0.95, clip_range=0.2, vf_coef=0.5,
                 ent_coef=0.0, learning_rate=3e-4, max_grad_norm=0.5, adam_eps=1e-5, normalize_advantage=True, seed=0, use_graph=True):
        self.env, self.dev = env, env.device
        c = env.core
        N, D, A = c.num_envs, c.obs_dim, c.act_dim
        self.pol = PolicyValueNet(D, A, c.act_is_int, device=self.dev, seed=seed, capacity=max(N, batch_size))
        self.lib = self.pol.lib
        self.n_steps, self.batch_size, self.n_epochs = n_steps, min(batch_size, n_steps * N), n_epochs
        assert self.batch_size % 64 == 0 and N % 64 == 0
        self.gamma, self.lam, self.clip, self.vf, self.ent = gamma, gae_lambda, clip_range, vf_coef, ent_coef
        self.lr, self.max_norm, self.eps, self.norm_adv = learning_rate, max_grad_norm, adam_eps, normalize_advantage
        T = n_steps
        f = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.dev)
        self.b_obs, self.b_act, self.b_logp, self.b_val, self.b_rew = f(T, N, D), f(T, N, A), f(T, N), f(T, N), f(T, N)
        self.b_done = torch.zeros((T, N), dtype=torch.uint8, device=self.dev)
        self.b_adv, self.b_ret, self.last_val = f(T, N), f(T, N), f(N)
        self.stats = f(8)
        self.world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        self.gen = torch.Generator(device=self.dev).manual_seed(seed)
        self.obs = env.reset().clone()
        self.num_timesteps = 0
        self.ep_returns, self.ep_lengths, self.ep_scores = [], [], []
        self.launches = 0
        self.use_graph, self._graphs, self._idx, self._graph_gen = use_graph, None, None, 0

    def collect_rollouts(self):
        env, pol = self.env, self.pol
        c = env.core
        for t in range(self.n_steps):
            act_env, act_raw, logp, val = pol.forward(self.obs)
            self.b_obs[t].copy_(self.obs); self.b_act[t].copy_(act_raw); self.b_logp[t].copy_(logp); self.b_val[t].copy_(val)
            obs, rew, term, _ = env.step(act_env, with_infos=False)        # episode statistics stay on the device (env.monitor)
            self.b_rew[t].copy_(rew); self.b_done[t].copy_(c.terminated)
            self.obs.copy_(obs)
        pol.check_status()
        _, _, _, val = pol.forward(self.obs)
        self.last_val.copy_(val)
        s = torch.cuda.current_stream(self.dev).cuda_stream
        rc = self.lib.fm_gae(self.b_rew.data_ptr(), self.b_val.data_ptr(), self.b_done.data_ptr(), self.last_val.data_ptr(), self.n_steps,
                             env.num_envs, self.gamma, self.lam, self.b_adv.data_ptr(), self.b_ret.data_ptr(), s)
        pol._chk(rc)
        self.launches += 1
        self.num_timesteps += self.n_steps * env.num_envs * self.world

    def episode_stats(self):
        """`rollout/ep_rew_mean`, `rollout/ep_len_mean` (SB3 Monitor + logger) and `rollout/ep_score_mean_bucket{k}`
        (AdditionalMetricsCallback, learning.py:53-72), read from the on-device monitor once per rollout; with several ranks the
        sums are all-reduced so that every rank logs the job-wide means."""
        m = self.env.monitor.metrics()
        if self.world > 1:
            keys = ["rollout/ep_rew_mean", "rollout/ep_len_mean", "rollout/ep_score_mean_bucket0", "rollout/ep_score_mean_bucket1"]
            v = torch.tensor([m.get(k, 0.0) for k in keys] + [1.0 if k in m else 0.0 for k in keys], dtype=torch.float64, device=self.dev)
            torch.distributed.all_reduce(v)
            for i, k in enumerate(keys):
                if v[4 + i] > 0:
                    m[k] = float(v[i] / v[4 + i])
        return m

    # ---- checkpoints (SyncifiedCheckpointCallback / CheckpointCallback, learning.py:75-86,113-117)
    def save(self, path, include_env=True):
        """SB3-layout zip: `policy.pth` (ActorCriticPolicy state-dict), `policy.optimizer.pth` (torch Adam state-dict over the same
        13 tensors in state-dict order), `data` (hyper-parameters, JSON), `pytorch_variables.pth`, `_stable_baselines3_version`;
        plus `fm_env_state.npz` (env SoA state, PCG64 streams, monitor) so that a run can RESUME — the reference can only restart."""
        import json
        pol = self.pol
        sd = pol.sb3_state_dict()
        order = ["log_std"] if not pol.discrete else []
        order += ["mlp_extractor.policy_net.0.weight", "mlp_extractor.policy_net.0.bias", "mlp_extractor.policy_net.2.weight",
                  "mlp_extractor.policy_net.2.bias", "mlp_extractor.value_net.0.weight", "mlp_extractor.value_net.0.bias",
                  "mlp_extractor.value_net.2.weight", "mlp_extractor.value_net.2.bias", "action_net.weight", "action_net.bias",
                  "value_net.weight", "value_net.bias"]
        m_sd, v_sd = pol.sb3_state_dict(pol.adam_m), pol.sb3_state_dict(pol.adam_v)
        opt = {"state": {i: {"step": torch.tensor(float(pol.adam_step)), "exp_avg": m_sd[k], "exp_avg_sq": v_sd[k]} for i, k in enumerate(order)},
               "param_groups": [{"lr": self.lr, "betas": (0.9, 0.999), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                                 "params": list(range(len(order)))}]}
        data = {"policy_class": "ActorCriticPolicy", "n_steps": self.n_steps, "batch_size": self.batch_size, "n_epochs": self.n_epochs,
                "gamma": self.gamma, "gae_lambda": self.lam, "clip_range": self.clip, "vf_coef": self.vf, "ent_coef": self.ent,
                "learning_rate": self.lr, "max_grad_norm": self.max_norm, "normalize_advantage": self.norm_adv,
                "num_timesteps": self.num_timesteps, "n_envs": self.env.num_envs, "_n_updates": pol.adam_step,
                "policy_kwargs": {"net_arch": [128, 128]}, "env_class": self.env.env_class, "num_arms": self.env.num_arms,
                "adam_step": pol.adam_step, "rollout_seed_calls": pol._call, "policy_seed": pol.seed,
                "perm_generator_state": self.gen.get_state().cpu().numpy().tolist()}

        def pth(obj):
            b = io.BytesIO(); torch.save(obj, b); return b.getvalue()
        with zipfile.ZipFile(path, "w") as z:
            z.writestr("data", json.dumps(data, indent=1))
            z.writestr("policy.pth", pth(sd)); z.writestr("policy.optimizer.pth", pth(opt)); z.writestr("pytorch_variables.pth", pth({}))
            z.writestr("_stable_baselines3_version", "2.3.2")
            if include_env:
                c, mon = self.env.core, self.env.monitor
                b = io.BytesIO()
                np.savez_compressed(b, f=c.f.cpu().numpy(), d=c.d.cpu().numpy(), i=c.i.cpu().numpy(), rng=c.rng.cpu().numpy(),
                                    obs=self.obs.cpu().numpy(), hist=mon.hist.cpu().numpy(), first=mon.first.cpu().numpy(),
                                    count=mon.count.cpu().numpy(), ring=mon.ring.cpu().numpy(), ring_head=mon.ring_head.cpu().numpy())
                z.writestr("fm_env_state.npz", b.getvalue())

    def load(self, path, load_env=True):
        import json
        pol = self.pol
        with zipfile.ZipFile(path) as z:
            pol.set_sb3_state_dict(torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True))
            names = z.namelist()
            if "policy.optimizer.pth" in names:
                opt = torch.load(io.BytesIO(z.read("policy.optimizer.pth")), map_location="cpu", weights_only=True)
                order = (["log_std"] if not pol.discrete else []) + [k for k in pol.sb3_state_dict().keys() if k != "log_std"]
                if opt.get("state"):
                    pol.set_sb3_state_dict({k: opt["state"][i]["exp_avg"] for i, k in enumerate(order)}, into=pol.adam_m)
                    pol.set_sb3_state_dict({k: opt["state"][i]["exp_avg_sq"] for i, k in enumerate(order)}, into=pol.adam_v)
                    pol.adam_step = int(float(opt["state"][0]["step"]))
            data = json.loads(z.read("data"))
            self.num_timesteps = int(data.get("num_timesteps", 0))
            pol._call = int(data.get("rollout_seed_calls", pol._call)); pol.seed = int(data.get("policy_seed", pol.seed))
            if "perm_generator_state" in data:
                self.gen.set_state(torch.tensor(data["perm_generator_state"], dtype=torch.uint8))
            if load_env and "fm_env_state.npz" in names:
                st = np.load(io.BytesIO(z.read("fm_env_state.npz")))
                c, mon = self.env.core, self.env.monitor
                for dst, k in ((c.f, "f"), (c.d, "d"), (c.i, "i"), (self.obs, "obs"), (mon.hist, "hist"), (mon.first, "first"),
                               (mon.count, "count"), (mon.ring, "ring"), (mon.ring_head, "ring_head")):
                    dst.copy_(torch.from_numpy(st[k]).to(dst.device))
                c.rng.copy_(torch.from_numpy(st["rng"].view(np.int64)).to(c.rng.device))
        return self

    def _minibatch_eager(self, idx_ptr, step):
        pol = self.pol
        s = torch.cuda.current_stream(self.dev).cuda_stream
        total = self.n_steps * self.env.num_envs
        obs = self.b_obs.view(total, -1); act = self.b_act.view(total, -1)
        pol._chk(self.lib.fm_ppo_update(C.byref(pol.net), C.byref(pol._ws), obs.data_ptr(), act.data_ptr(), self.b_logp.data_ptr(),
                                        self.b_adv.data_ptr(), self.b_ret.data_ptr(), idx_ptr, self.batch_size, self.clip, self.vf, self.ent,
                                        int(self.norm_adv), self.stats.data_ptr(), s))
        _dist.allreduce_sum_(pol.grads)      # the only inter-GPU exchange of the path: one sum all-reduce of the flat gradient (NCCL / NVLink)
        pol._chk(self.lib.fm_adam_step(C.byref(pol.net), self.lr, 0.9, 0.999, self.eps, self.max_norm, 1.0 / self.world, step,
                                       self.stats.data_ptr(), s))

    def _capture_graphs(self):
        """One optimiser step = 30 short launches (forward: the fused tcgen05 MLP per net; backward: K <= 384, N = 128 mma.sync GEMMs): captured once into CUDA graphs and
        replayed with a static minibatch-index buffer; Adam's step count lives on the device (stats[7]).  With several ranks the
        NCCL all-reduce stays between two graphs (update | all-reduce | Adam)."""
        pol = self.pol
        total = self.n_steps * self.env.num_envs
        obs = self.b_obs.view(total, -1); act = self.b_act.view(total, -1)
        self._idx = torch.zeros(self.batch_size, dtype=torch.int32, device=self.dev)
        g_upd, g_adam = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        cs = lambda: torch.cuda.current_stream(self.dev).cuda_stream
        with torch.cuda.graph(g_upd):
            pol._chk(self.lib.fm_ppo_update(C.byref(pol.net), C.byref(pol._ws), obs.data_ptr(), act.data_ptr(), self.b_logp.data_ptr(),
                                            self.b_adv.data_ptr(), self.b_ret.data_ptr(), self._idx.data_ptr(), self.batch_size, self.clip,
                                            self.vf, self.ent, int(self.norm_adv), self.stats.data_ptr(), cs()))
            if self.world == 1:
                pol._chk(self.lib.fm_adam_step(C.byref(pol.net), self.lr, 0.9, 0.999, self.eps, self.max_norm, 1.0, 0, self.stats.data_ptr(), cs()))
        if self.world > 1:
            with torch.cuda.graph(g_adam):
                pol._chk(self.lib.fm_adam_step(C.byref(pol.net), self.lr, 0.9, 0.999, self.eps, self.max_norm, 1.0 / self.world, 0,
                                               self.stats.data_ptr(), cs()))
        self._graphs = (g_upd, g_adam)
        self._graph_gen = pol.ws_generation

    def train(self):
        pol = self.pol
        total = self.n_steps * self.env.num_envs
        nb = total // self.batch_size
        self.stats[:7].zero_()
        n_updates = 0
        # the first call runs launch by launch (it also sets the kernels' attributes); the graphs are captured after it
        if self._graphs is not None and self._graph_gen != pol.ws_generation:
            self._graphs = None                      # the workspace was reallocated under the captured graphs: recapture
        use_graph = self.use_graph and self._graphs is not None
        if use_graph:
            self.stats[7:8].fill_(float(pol.adam_step))
        for ep in range(self.n_epochs):
            perm = torch.randperm(total, device=self.dev, generator=self.gen, dtype=torch.int64).to(torch.int32)
            for b in range(nb):
                pol.adam_step += 1
                if use_graph:
                    self._idx.copy_(perm[b * self.batch_size:(b + 1) * self.batch_size])
                    self._graphs[0].replay()
                    if self.world > 1:
                        _dist.allreduce_sum_(pol.grads)
                        self._graphs[1].replay()
                else:
                    idx = perm[b * self.batch_size:(b + 1) * self.batch_size].contiguous()
                    self._minibatch_eager(idx.data_ptr(), pol.adam_step)
                n_updates += 1
                self.launches += 30
        if self.use_graph and self._graphs is None:
            self._capture_graphs()
        return {"n_updates": n_updates, "policy_loss": float(self.stats[0]) / max(n_updates, 1), "value_loss": float(self.stats[1]) / max(n_updates, 1),
                "approx_kl": float(self.stats[2]) / max(n_updates, 1), "clip_fraction": float(self.stats[3]) / max(n_updates, 1),
                "cuda_graph": bool(use_graph)}

    def learn(self, total_timesteps: int, log=None):
        it = 0
        while self.num_timesteps < total_timesteps:
            self.collect_rollouts()
            info = self.train()
            it += 1
            if log:
                log({"iteration": it, "timesteps": self.num_timesteps, **info, **self.episode_stats()})
        return self
