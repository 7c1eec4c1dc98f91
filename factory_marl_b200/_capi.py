"""ctypes view of the C-ABI in include/fm_api.h (libfm_b200.so).  No torch types cross the ABI: only raw
device pointers, sizes and a cudaStream_t.  `FmCore` is backend-agnostic over *where the arrays live*: the
product passes CUDA tensors (vec_env.py); the host-compiled logic-test build in tests/ passes numpy arrays."""
from __future__ import annotations

import ctypes as C

import numpy as np

ENV_CLASSES = {
    "FactoryManipulationEnv": 0, "SingleFullRLProgressRewardEnv": 1, "SingleDeltaProgressRewardEnv": 2,
    "AllFullRLProgressRewardEnv": 3, "AllDeltaProgressRewardEnv": 4, "PauseIKToggleEnv": 5, "BackupIKToggleEnv": 6,
}


class fm_config(C.Structure):
    _fields_ = [("n_envs", C.c_int32), ("n_arms", C.c_int32), ("env_class", C.c_int32), ("auto_reset", C.c_int32),
                ("pgs_outer", C.c_int32), ("pgs_inner", C.c_int32), ("envs_per_block", C.c_int32), ("reserved", C.c_int32),
                ("pgs_tol", C.c_float), ("reserved_f", C.c_float),
                ("initial_conveyor_speed", C.c_double), ("conveyor_acceleration", C.c_double), ("pt_time", C.c_double),
                ("force_contact_threshold", C.c_double), ("control_frequency", C.c_double), ("spawn_freq", C.c_double),
                ("spawn_freq_increase", C.c_double), ("gripper_to_closest_cube_reward_factor", C.c_double),
                ("closest_cube_to_bucket_reward_factor", C.c_double), ("small_action_norm_reward_factor", C.c_double),
                ("base_reward", C.c_double)]


class fm_state_view(C.Structure):
    _fields_ = [("env", C.c_int32), ("qpos", C.POINTER(C.c_double)), ("qvel", C.POINTER(C.c_double)),
                ("ctrl_target", C.POINTER(C.c_double))]


class fm_debug_view(C.Structure):
    _fields_ = [("env", C.c_int32), ("ncon", C.c_int32), ("flags", C.c_int32), ("reserved", C.c_int32),
                ("pair", (C.c_int32 * 2) * 64), ("pos", (C.c_float * 3) * 64), ("normal", (C.c_float * 3) * 64),
                ("force", (C.c_float * 4) * 64), ("site_xpos", (C.c_float * 3) * 6), ("mass_matrix", (C.c_float * 45) * 6),
                ("bias", (C.c_float * 9) * 6), ("qacc_smooth", C.c_float * 115), ("qacc", C.c_float * 115)]


EXPORTS = ["fm_create", "fm_destroy", "fm_obs_dim", "fm_act_dim", "fm_state_rows", "fm_bind_state", "fm_field_offset",
           "fm_reset", "fm_step", "fm_ik_policy", "fm_get_state", "fm_set_state", "fm_debug_substeps", "fm_last_error"]


def bind(lib: C.CDLL) -> C.CDLL:
    vp = C.c_void_p
    lib.fm_last_error.restype = C.c_char_p
    lib.fm_create.argtypes = [C.POINTER(fm_config), vp, C.c_size_t, C.POINTER(vp)]
    lib.fm_destroy.argtypes = [vp]
    lib.fm_obs_dim.argtypes = [vp]
    lib.fm_act_dim.argtypes = [vp]
    lib.fm_state_rows.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.fm_bind_state.argtypes = [vp, vp, vp, vp, vp]
    lib.fm_field_offset.argtypes = [vp, C.c_char_p]
    lib.fm_reset.argtypes = [vp, vp, vp, vp]
    lib.fm_set_env_order.argtypes = [vp, vp]
    lib.fm_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.fm_ik_policy.argtypes = [vp, vp, vp]
    lib.fm_get_state.argtypes = [vp, C.POINTER(fm_state_view)]
    lib.fm_set_state.argtypes = [vp, C.POINTER(fm_state_view)]
    lib.fm_debug_substeps.argtypes = [vp, C.c_int, C.POINTER(fm_debug_view), vp]
    return lib


def pcg64_state_words(seed) -> np.ndarray:
    """numpy default_rng(seed) initial PCG64 state as 4 x uint64 (state_hi, state_lo, inc_hi, inc_lo):
    the TaskManager RNG of task_utils.py:19."""
    st = np.random.PCG64(seed).state["state"]
    m = (1 << 64) - 1
    return np.array([st["state"] >> 64, st["state"] & m, st["inc"] >> 64, st["inc"] & m], dtype=np.uint64)


class FmCore:
    """Handle + caller-owned state arrays.  `backend` supplies zeros(shape, dtype)->array, ptr(array)->int,
    to_numpy(array)->np.ndarray, from_numpy(np)->array and stream()->int."""

    def __init__(self, lib, backend, blob: bytes, num_envs: int, num_arms: int, env_class, seeds=None, auto_reset=True,
                 pgs_outer=0, pgs_inner=0, pgs_tol=0.0, envs_per_block=0, sleep_hidden=True, initial_conveyor_speed=0.1,
                 conveyor_acceleration=0.001, pt_time=0.2, force_contact_threshold=200.0, control_frequency=10,
                 spawn_freq=1 / 10, spawn_freq_increase=1.001, gripper_to_closest_cube_reward_factor=0.0,
                 closest_cube_to_bucket_reward_factor=0.0, small_action_norm_reward_factor=0.0, base_reward=0.0):
        self.lib, self.be = lib, backend
        self.env_class = ENV_CLASSES[env_class] if isinstance(env_class, str) else int(env_class)
        self.num_envs, self.num_arms = int(num_envs), int(num_arms)
        self.cfg = fm_config(n_envs=num_envs, n_arms=num_arms, env_class=self.env_class, auto_reset=int(auto_reset),
                             pgs_outer=pgs_outer, pgs_inner=pgs_inner, envs_per_block=envs_per_block, pgs_tol=pgs_tol,
                             reserved=0 if sleep_hidden else 1,
                             initial_conveyor_speed=initial_conveyor_speed, conveyor_acceleration=conveyor_acceleration,
                             pt_time=pt_time, force_contact_threshold=force_contact_threshold,
                             control_frequency=control_frequency, spawn_freq=spawn_freq,
                             spawn_freq_increase=spawn_freq_increase,
                             gripper_to_closest_cube_reward_factor=gripper_to_closest_cube_reward_factor,
                             closest_cube_to_bucket_reward_factor=closest_cube_to_bucket_reward_factor,
                             small_action_norm_reward_factor=small_action_norm_reward_factor, base_reward=base_reward)
        self._blob = blob
        h = C.c_void_p()
        self._chk(lib.fm_create(C.byref(self.cfg), blob, len(blob), C.byref(h)))
        self.h = h
        self.obs_dim = lib.fm_obs_dim(h)
        self.act_dim = lib.fm_act_dim(h)
        self.act_is_int = self.env_class in (5, 6)
        self.uses_ik = self.env_class != 3
        self.is_toggle = self.env_class in (5, 6)
        nf, nd, ni = C.c_int32(), C.c_int32(), C.c_int32()
        lib.fm_state_rows(h, C.byref(nf), C.byref(nd), C.byref(ni))
        N = self.num_envs
        be = backend
        self.f = be.zeros((nf.value, N), np.float32)
        self.d = be.zeros((nd.value, N), np.float64)
        self.i = be.zeros((ni.value, N), np.int32)
        ig = self.off("I_IGNORE")
        ii = np.zeros((ni.value, N), np.int32)
        ii[ig:ig + num_arms * num_arms] = -2            # IKPolicy.ignore_objects = {} (ik_policy.py:79)
        ii[self.off("I_FSM_TGT"):self.off("I_FSM_TGT") + num_arms] = -1
        self.i = be.from_numpy(ii)
        ff = np.zeros((nf.value, N), np.float32)
        lc = self.off("F_IK_LAST_CTRL")
        ff[lc:lc + 8 * num_arms] = np.tile(np.array([-0.5, -0.5, 0.0, 1.0, 0.0, -1.6, 0.0, 0.06], np.float32), num_arms)[:, None]
        self.f = be.from_numpy(ff)
        if seeds is None:
            seeds = [42] * N                              # learning.py:26: every worker is built with seed=42
        elif np.isscalar(seeds):
            seeds = [int(seeds)] * N
        cache = {}
        rng = np.zeros((4, N), np.uint64)
        for e, s in enumerate(seeds):
            if s not in cache:
                cache[s] = pcg64_state_words(s)
            rng[:, e] = cache[s]
        self.rng = be.from_numpy(rng)
        self._chk(lib.fm_bind_state(h, be.ptr(self.f), be.ptr(self.d), be.ptr(self.i), be.ptr(self.rng)))
        # outputs
        self.obs = be.zeros((N, self.obs_dim), np.float32)
        self.reward = be.zeros((N,), np.float32)
        self.terminated = be.zeros((N,), np.uint8)
        self.scores = be.zeros((N, 2), np.int32)
        self.flags = be.zeros((N,), np.uint8)
        self.terminal_obs = be.zeros((N, self.obs_dim), np.float32)
        self.ep_info = be.zeros((N, 4), np.float32)

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("libfm: " + self.lib.fm_last_error().decode())

    def off(self, name: str) -> int:
        o = self.lib.fm_field_offset(self.h, name.encode())
        if o < 0:
            raise KeyError(name)
        return o

    def close(self):
        if getattr(self, "h", None):
            self.lib.fm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference-shaped calls
    def set_env_order(self, order=None):
        """order: int32[N] permutation owned by the backend (kept alive here), or None for the identity (fm_set_env_order)."""
        self._order = order
        self._chk(self.lib.fm_set_env_order(self.h, self.be.ptr(order) if order is not None else None))

    def reset(self, mask=None):
        be = self.be
        self._chk(self.lib.fm_reset(self.h, be.ptr(mask) if mask is not None else None, be.ptr(self.obs), be.stream()))
        if self.is_toggle:       # IKTogglingEnv._process_observation also runs at reset (environments.py:244)
            self._chk(self.lib.fm_ik_policy(self.h, be.ptr(self.obs), be.stream()))
        return self.obs

    def step(self, actions):
        """actions: array owned by the backend, f32[N, act_dim] or i32[N, n_arms]."""
        be = self.be
        self._chk(self.lib.fm_step(self.h, be.ptr(actions), be.ptr(self.obs), be.ptr(self.reward), be.ptr(self.terminated),
                                   be.ptr(self.scores), be.ptr(self.flags), be.ptr(self.terminal_obs), be.ptr(self.ep_info),
                                   be.stream()))
        return self.obs, self.reward, self.terminated

    def env_step(self, actions):
        """One FactoryManipulationEnv.step for every env, with the IK base policy run where the reference runs it:
        before the step for the base / Single* / *Delta classes (environments.py:120-127,419,456,531), inside
        _process_observation (after the step) for the toggling classes (environments.py:574-576)."""
        if self.env_class in (0, 1, 2, 4):
            self.ik_policy(write_obs=False)
        out = self.step(actions)
        if self.is_toggle:
            self.ik_policy(write_obs=True)
        return out

    def ik_policy(self, write_obs=True):
        be = self.be
        self._chk(self.lib.fm_ik_policy(self.h, be.ptr(self.obs) if write_obs else None, be.stream()))

    def get_state(self, env=0):
        nq, nv, nu = 71 + 9 * self.num_arms, 61 + 9 * self.num_arms, 1 + 8 * self.num_arms
        q, v, c = np.zeros(nq), np.zeros(nv), np.zeros(nu)
        view = fm_state_view(env=env, qpos=q.ctypes.data_as(C.POINTER(C.c_double)), qvel=v.ctypes.data_as(C.POINTER(C.c_double)),
                             ctrl_target=c.ctypes.data_as(C.POINTER(C.c_double)))
        self._chk(self.lib.fm_get_state(self.h, C.byref(view)))
        return q, v, c

    def set_state(self, env, qpos=None, qvel=None, ctrl_target=None):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float64) for a in (qpos, qvel, ctrl_target)]
        ptr = [None if a is None else a.ctypes.data_as(C.POINTER(C.c_double)) for a in arrs]
        view = fm_state_view(env=env, qpos=ptr[0], qvel=ptr[1], ctrl_target=ptr[2])
        self._chk(self.lib.fm_set_state(self.h, C.byref(view)))

    def debug_substeps(self, n, env=0):
        view = fm_debug_view(env=env)
        self._chk(self.lib.fm_debug_substeps(self.h, n, C.byref(view), self.be.stream()))
        nc = view.ncon
        return dict(ncon=nc, flags=view.flags, pair=np.ctypeslib.as_array(view.pair)[:nc].copy(),
                    pos=np.ctypeslib.as_array(view.pos)[:nc].copy(), normal=np.ctypeslib.as_array(view.normal)[:nc].copy(),
                    force=np.ctypeslib.as_array(view.force)[:nc].copy(),
                    mass_matrix=np.ctypeslib.as_array(view.mass_matrix)[:self.num_arms].copy(),
                    bias=np.ctypeslib.as_array(view.bias)[:self.num_arms].copy(),
                    qacc_smooth=np.ctypeslib.as_array(view.qacc_smooth)[:61 + 9 * self.num_arms].copy(),
                    qacc=np.ctypeslib.as_array(view.qacc)[:61 + 9 * self.num_arms].copy())
