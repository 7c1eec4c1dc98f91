"""ctypes binding of the TEST ORACLE (oracle/libfm_oracle.so).  Test infrastructure: import only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ENV_CLASSES = {"FactoryManipulationEnv": 0, "SingleFullRLProgressRewardEnv": 1, "SingleDeltaProgressRewardEnv": 2,
               "AllFullRLProgressRewardEnv": 3, "AllDeltaProgressRewardEnv": 4, "PauseIKToggleEnv": 5,
               "BackupIKToggleEnv": 6}


def build(force=False):
    so = os.path.join(_HERE, "libfm_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("fm_oracle.cpp", "fmo_physics.hpp", "fmo_env.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.fmo_create.restype = C.c_void_p
        _LIB.fmo_create.argtypes = [C.c_void_p, C.c_size_t]
        _LIB.fmo_last_error.restype = C.c_char_p
        _LIB.fmo_pcg64_next.restype = C.c_uint64
        for f in ("fmo_destroy", "fmo_sizes", "fmo_set_state", "fmo_get_state", "fmo_step1", "fmo_step2", "fmo_get_xpos",
                  "fmo_get_geom", "fmo_get_dyn", "fmo_ncon", "fmo_nefc", "fmo_get_contacts", "fmo_get_efc",
                  "fmo_solver_stats", "fmo_env_create", "fmo_env_reset", "fmo_env_step", "fmo_env_obs_dim",
                  "fmo_env_get", "fmo_env_set", "fmo_ik_solve", "fmo_set_warmstart"):
            getattr(_LIB, f).argtypes = None
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    def __init__(self, blob: bytes):
        self.L = lib()
        self._blob = blob
        self.h = C.c_void_p(self.L.fmo_create(blob, len(blob)))
        if not self.h:
            raise RuntimeError(self.L.fmo_last_error().decode())
        s = np.zeros(11, np.int32)
        self.L.fmo_sizes(self.h, _p(s))
        (self.nb, self.ng, self.nj, self.nq, self.nv, self.nu, self.n_arms, self.n_cubes, self.n_clusters,
         self.n_cpairs, self.n_geom_pairs) = [int(x) for x in s]
        self.obs_dim = None

    def __del__(self):
        try:
            self.L.fmo_destroy(self.h)
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.fmo_last_error().decode())

    # ---- physics level
    def set_state(self, qpos, qvel):
        qpos = np.ascontiguousarray(qpos, np.float64); qvel = np.ascontiguousarray(qvel, np.float64)
        self._chk(self.L.fmo_set_state(self.h, _p(qpos), _p(qvel)))

    def get_state(self):
        q = np.zeros(self.nq); v = np.zeros(self.nv)
        self.L.fmo_get_state(self.h, _p(q), _p(v))
        return q, v

    def set_warmstart(self, qacc):
        qacc = np.ascontiguousarray(qacc, np.float64)
        self.L.fmo_set_warmstart(self.h, _p(qacc))

    def step1(self):
        self._chk(self.L.fmo_step1(self.h))

    def step2(self, ctrl):
        ctrl = np.ascontiguousarray(ctrl, np.float64)
        self._chk(self.L.fmo_step2(self.h, _p(ctrl)))

    def xpos(self):
        p = np.zeros((self.nb, 3)); q = np.zeros((self.nb, 4))
        self.L.fmo_get_xpos(self.h, _p(p), _p(q))
        return p, q

    def geoms(self):
        p = np.zeros((self.ng, 3)); m = np.zeros((self.ng, 9))
        self.L.fmo_get_geom(self.h, _p(p), _p(m))
        return p, m.reshape(-1, 3, 3)

    def dyn(self):
        nv = self.nv
        out = dict(M=np.zeros((nv, nv)), bias=np.zeros(nv), qacc_smooth=np.zeros(nv), qacc=np.zeros(nv),
                   qfrc_constraint=np.zeros(nv), qfrc_actuator=np.zeros(nv))
        self.L.fmo_get_dyn(self.h, *[_p(out[k]) for k in ("M", "bias", "qacc_smooth", "qacc", "qfrc_constraint", "qfrc_actuator")])
        return out

    def contacts(self):
        n = self.L.fmo_ncon(self.h)
        g = np.zeros((n, 2), np.int32); dd = np.zeros((n, 20))
        if n:
            self.L.fmo_get_contacts(self.h, _p(g), _p(dd))
        return g, dd

    def efc(self, with_J=False):
        n = self.L.fmo_nefc(self.h)
        t = np.zeros(n, np.int32); pos = np.zeros(n); aref = np.zeros(n); R = np.zeros(n); f = np.zeros(n)
        J = np.zeros((n, self.nv)) if with_J else None
        if n:
            self.L.fmo_get_efc(self.h, _p(t), _p(J), _p(pos), _p(aref), _p(R), _p(f))
        return dict(type=t, J=J, pos=pos, aref=aref, R=R, force=f)

    def solver_stats(self):
        o = np.zeros(3); self.L.fmo_solver_stats(self.h, _p(o)); return o

    # ---- env level
    def env_create(self, env_class="AllFullRLProgressRewardEnv", seed=42, initial_conveyor_speed=0.1,
                   conveyor_acceleration=0.001, pt_time=0.2, force_contact_threshold=200.0, control_frequency=10,
                   spawn_freq=1 / 10, spawn_freq_increase=1.001, gripper_to_closest_cube_reward_factor=0.0,
                   closest_cube_to_bucket_reward_factor=0.0, small_action_norm_reward_factor=0.0, base_reward=0.0):
        st = np.random.PCG64(seed).state["state"]
        words = []
        for v in (st["state"] >> 64, st["state"] & (2 ** 64 - 1), st["inc"] >> 64, st["inc"] & (2 ** 64 - 1)):
            words += [v & 0xFFFFFFFF, v >> 32]
        icfg = np.array([ENV_CLASSES[env_class] if isinstance(env_class, str) else env_class] + words, np.uint32).astype(np.int32)
        dcfg = np.array([initial_conveyor_speed, conveyor_acceleration, pt_time, force_contact_threshold,
                         control_frequency, spawn_freq, spawn_freq_increase, gripper_to_closest_cube_reward_factor,
                         closest_cube_to_bucket_reward_factor, small_action_norm_reward_factor, base_reward], np.float64)
        self._chk(self.L.fmo_env_create(self.h, _p(icfg), _p(dcfg)))
        self.obs_dim = self.L.fmo_env_obs_dim(self.h)

    def env_reset(self):
        obs = np.zeros(self.obs_dim, np.float32)
        self._chk(self.L.fmo_env_reset(self.h, _p(obs)))
        return obs

    def env_step(self, action):
        obs = np.zeros(self.obs_dim, np.float32)
        rew = np.zeros(1); info = np.zeros(6, np.int32)
        a = np.asarray(action)
        af = np.ascontiguousarray(a, np.float32) if a.dtype.kind == "f" else np.zeros(8 * self.n_arms, np.float32)
        ai = np.ascontiguousarray(a, np.int32) if a.dtype.kind in "iu" else np.zeros(self.n_arms, np.int32)
        self._chk(self.L.fmo_env_step(self.h, _p(af), _p(ai), _p(obs), _p(rew), _p(info)))
        return obs, float(rew[0]), dict(terminated=bool(info[0]), out_of_reach=bool(info[1]), force_terminate=bool(info[2]),
                                        scores=(int(info[3]), int(info[4])), num_obj=int(info[5]))

    def env_get(self, what, n):
        o = np.zeros(n); self.L.fmo_env_get(self.h, what.encode(), _p(o)); return o

    def env_set(self, what, arr):
        arr = np.ascontiguousarray(arr, np.float64); self.L.fmo_env_set(self.h, what.encode(), _p(arr))

    def ik_solve(self, arm, tpos, tquat, q7):
        q = np.ascontiguousarray(q7, np.float64).copy(); steps = np.zeros(1, np.int32)
        ok = self.L.fmo_ik_solve(self.h, arm, _p(np.ascontiguousarray(tpos, np.float64)), _p(np.ascontiguousarray(tquat, np.float64)), _p(q), _p(steps))
        return bool(ok), q, int(steps[0])
