"""One interface over the two things the parity tests drive: the fp64 oracle (oracle/) and the kernel behind the C-ABI
(`FmCore`: CUDA library on the GPU, host-compiled phase build in the CPU suite).  Test infrastructure only."""
import numpy as np

from factory_marl_b200 import model as fm
from oracle.oracle_py import Oracle


class Scene:
    """A full env state in the reference's terms: MuJoCo qpos/qvel, BaseEnv.ctrl_target, TaskManager lists and counters."""

    def __init__(self, qpos, qvel, ctrl_target, in_scene, step, since_spawn=0, conveyor_speed=None, spawn_freq=None, n_arms=2):
        self.qpos, self.qvel, self.ctrl_target = np.asarray(qpos, float), np.asarray(qvel, float), np.asarray(ctrl_target, float)
        self.in_scene = list(in_scene)
        self.out_scene = [k for k in range(10) if k not in self.in_scene]
        self.step, self.since_spawn = int(step), int(since_spawn)
        self.conveyor_speed = 0.1 + 1e-4 * step if conveyor_speed is None else conveyor_speed
        self.spawn_freq = 0.1 * n_arms * 1.001 ** step if spawn_freq is None else spawn_freq
        self.play_time = 0.1 * step


class OracleSim:
    kind = "oracle"

    def __init__(self, n_arms, cls, seed=42, **kw):
        self.n_arms, self.cls = n_arms, cls
        self.M = fm.build_model(n_arms)
        self.o = Oracle(fm.build_blob(n_arms))
        self.o.env_create(cls, seed=seed, **kw)
        self.is_int = cls in ("PauseIKToggleEnv", "BackupIKToggleEnv")

    def reset(self):
        return self.o.env_reset()

    def step(self, action):
        a = np.asarray(action)
        a = a.astype(np.int32) if self.is_int else a.astype(np.float32)
        obs, r, info = self.o.env_step(a)
        return obs, r, info

    def transplant(self, sc: Scene):
        o = self.o
        o.set_state(sc.qpos, sc.qvel); o.step1()
        o.env_set("ctrl_target", sc.ctrl_target)
        lists = np.zeros(23); lists[0] = len(sc.in_scene); lists[1:1 + len(sc.in_scene)] = sc.in_scene
        lists[11] = len(sc.out_scene); lists[12:12 + len(sc.out_scene)] = sc.out_scene
        o.env_set("lists", lists)
        o.env_set("scalars", [sc.conveyor_speed, sc.spawn_freq, sc.play_time, sc.step, sc.since_spawn, 0, 0])

    def state(self):
        q, v = self.o.get_state()
        return q, v, self.o.env_get("ctrl_target", 1 + 8 * self.n_arms)

    def fsm(self):
        return self.o.env_get("fsm", 3 * self.n_arms).astype(int)


class CoreSim:
    """env 0 of a small batch of identical envs behind the C-ABI; `be` = backend of tests/test_{emu,gpu}_parity.py."""
    kind = "core"

    def __init__(self, be, n_arms, cls, n_envs=4, seed=42, **kw):
        self.be, self.n_arms, self.cls, self.n = be, n_arms, cls, n_envs
        self.M = fm.build_model(n_arms)
        kw.setdefault("auto_reset", False)
        self.core = be.mk(fm.build_blob(n_arms), n_envs, n_arms, cls, seeds=seed, **kw)
        self.is_int = cls in ("PauseIKToggleEnv", "BackupIKToggleEnv")

    def reset(self):
        return self.be.to_np(self.core.reset())[0].copy()

    def step(self, action):
        a = np.asarray(action)
        a = a.astype(np.int32) if self.is_int else a.astype(np.float32)
        if a.size == 0:
            a = np.zeros(1, np.float32)
        c, be = self.core, self.be
        c.env_step(be.from_np(np.tile(a[None], (self.n, 1))))
        fl = int(be.to_np(c.flags)[0])
        sc = be.to_np(c.scores)[0]
        info = dict(terminated=bool(be.to_np(c.terminated)[0]), out_of_reach=bool(fl & 1), force_terminate=bool(fl & 2),
                    scores=(int(sc[0]), int(sc[1])), num_obj=int(be.to_np(c.i)[c.off("I_NIN"), 0]), cap_flags=fl >> 2)
        return be.to_np(c.obs)[0].copy(), float(be.to_np(c.reward)[0]), info

    def transplant(self, sc: Scene):
        c, be = self.core, self.be
        ii = be.to_np(c.i).copy(); dd = be.to_np(c.d).copy()
        ii[c.off("I_IN"):c.off("I_IN") + 10] = 0; ii[c.off("I_OUT"):c.off("I_OUT") + 10] = 0
        ii[c.off("I_IN"):c.off("I_IN") + len(sc.in_scene)] = np.array(sc.in_scene, np.int32).reshape(-1, 1)
        ii[c.off("I_OUT"):c.off("I_OUT") + len(sc.out_scene)] = np.array(sc.out_scene, np.int32).reshape(-1, 1)
        ii[c.off("I_NIN")] = len(sc.in_scene); ii[c.off("I_NOUT")] = len(sc.out_scene)
        ii[c.off("I_STEP")] = sc.step; ii[c.off("I_SINCE")] = sc.since_spawn
        ii[c.off("I_FAIL")] = 0; ii[c.off("I_OUTCNT")] = 0
        dd[c.off("D_CONV")] = sc.conveyor_speed; dd[c.off("D_SPAWN_FREQ")] = sc.spawn_freq; dd[c.off("D_PLAY_TIME")] = sc.play_time
        be.assign(c.i, ii); be.assign(c.d, dd)
        for e in range(self.n):
            c.set_state(e, sc.qpos, sc.qvel, sc.ctrl_target)

    def state(self):
        return self.core.get_state(0)

    def fsm(self):
        c = self.core; gi = self.be.to_np(c.i)
        return np.array([[gi[c.off("I_FSM_STATE") + a, 0], gi[c.off("I_FSM_CNT") + a, 0], gi[c.off("I_FSM_TGT") + a, 0]]
                         for a in range(self.n_arms)]).reshape(-1)

    def cap_flags(self):
        return int(self.be.to_np(self.core.i)[self.core.off("I_FLAGS")].max())


def hidden_qpos(M, n_arms):
    """qpos0 with every cube resting in its hide slot on the floor (task_utils.py:146-156 after it has settled)."""
    q = fm.qpos0(M)
    for k in range(10):
        q[1 + 7 * k:8 + 7 * k] = [4.0, 0.2 * k, M["cube_size"][k], 1, 0, 0, 0]
    return q


def arm_qadr(M, a, k):
    return int(M["jnt_qposadr"][M["arm_jnt"][a, k]])


def arm_vadr(M, a, k):
    return int(M["jnt_dofadr"][M["arm_jnt"][a, k]])
