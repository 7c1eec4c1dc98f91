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


# =====================================================================================================================
# Integer outputs at non-trivial values, production solver settings, contact forces, episode boundaries (sim adapters)
from tests.sim_adapters import CoreSim, OracleSim, Scene, hidden_qpos  # noqa: E402


def _settled_scene(n_arms=2):
    """three cubes riding the belt after 120 control steps (tests/golden/known_answers.npz, made by the oracle)"""
    g = np.load(GOLDEN + "/known_answers.npz")
    q, v = g["qpos120"].copy(), g["qvel120"].copy()
    ct = np.zeros(1 + 8 * n_arms); ct[0] = -0.1119
    return Scene(q, v, ct, [0, 1, 2], 120, since_spawn=28, n_arms=n_arms)


def _pair(be, n_arms, cls, **kw):
    return OracleSim(n_arms, cls, **kw), CoreSim(be, n_arms, cls, **kw)


def case_scores_both_buckets_snapshot(be):
    """cubes 0 and 1 dropped into bucket 0 and bucket 1 in the same control step while cube 2 rides the belt.  The reference takes
    ONE position snapshot (task_utils.py:97): bucket 0 pops index 0, then bucket 1's hit index 1 is applied to the shortened list
    and hides cube 2 instead of cube 1 — which is scored one step later.  Scores, list contents and flags must follow exactly."""
    for cls in ("FactoryManipulationEnv", "AllFullRLProgressRewardEnv"):
        kw = PROGRESS_KW if "Progress" in cls else {}
        o, c = _pair(be, 2, cls, **kw)
        sc = _settled_scene()
        M = o.M
        for k, x in ((0, 0.9), (1, -0.9)):
            sc.qpos[1 + 7 * k:8 + 7 * k] = [x, 0.7, 1.10, 1, 0, 0, 0]; sc.qvel[1 + 6 * k:7 + 6 * k] = 0
        act = np.zeros(16 if "Progress" in cls else 0, np.float32)
        hist = []
        for sim in (o, c):
            sim.reset(); sim.transplant(sc)
            h = []
            for t in range(3):
                obs, r, info = sim.step(act)
                h.append((info["scores"], info["num_obj"], info["terminated"], round(r, 3) if "Progress" not in cls else None))
            hist.append(h)
        assert hist[0] == hist[1], (cls, hist)
        assert hist[0][0][0] == (1, 1) and hist[0][0][1] == 1, hist[0]      # two hidden after step 1: cube 0 and (wrongly) cube 2
        assert hist[0][1][0] == (1, 2), hist[0]                             # cube 1 is still in bucket 1 and scores again
        if "Progress" not in cls:
            assert hist[0][0][3] == 2.0 and hist[0][1][3] == 1.0             # base reward = score delta (environments.py:148)
        assert c.cap_flags() == 0


def case_out_of_bounds_terminates(be):
    """a cube beyond |x| > 1.2, y < -1.5 or z < 0.9 ends the episode with out_of_reach (task_utils.py:84-95), bit 0 of flags"""
    for pos in ([1.25, 0.0, 1.2], [0.0, -1.55, 1.14], [0.5, 0.9, 0.85]):
        o, c = _pair(be, 2, "AllFullRLProgressRewardEnv", **PROGRESS_KW)
        sc = _settled_scene()
        sc.qpos[1:8] = pos + [1, 0, 0, 0]; sc.qvel[1:7] = 0
        res = []
        for sim in (o, c):
            sim.reset(); sim.transplant(sc)
            obs, r, info = sim.step(np.zeros(16, np.float32))
            res.append((info["terminated"], info["out_of_reach"], info["force_terminate"], info["scores"], info["num_obj"]))
        assert res[0] == res[1] and res[0][0] and res[0][1], (pos, res)


def case_force_termination_same_step(be, seeds=(0, 2, 8, 10)):
    """arms driven into the table / belt / each other by bang-bang joint targets: the 200 N criterion (base_env.py:226-236) must
    fire at the same control step as in the oracle, flags bit 1, and never before"""
    hit = 0
    for seed in seeds:
        o, c = _pair(be, 2, "AllFullRLProgressRewardEnv", **PROGRESS_KW)
        a = (np.random.default_rng(seed).integers(0, 2, 16) * 2 - 1).astype(np.float32) * 3
        o.reset(); c.reset()
        for t in range(30):
            _, _, io = o.step(a)
            _, _, ic = c.step(a)
            assert (io["force_terminate"], io["out_of_reach"]) == (ic["force_terminate"], ic["out_of_reach"]), (seed, t, io, ic)
            if io["force_terminate"]:
                hit += 1
                break
        assert c.cap_flags() == 0
    assert hit >= 3, hit


def case_contact_forces(be, **extra):
    """normal forces of the resting contacts after one sub-step from identical state: each cube's weight is shared by its four
    corner contacts; compared with the oracle's solve contact by contact"""
    blob = fm.build_blob(2); M = fm.build_model(2)
    sc = _settled_scene()
    o = Oracle(blob); o.set_state(sc.qpos, sc.qvel); o.step1()
    ctrl = np.zeros(17); ctrl[0] = -0.112
    core = be.mk(blob, 4, 2, "AllFullRLProgressRewardEnv", auto_reset=False, **extra)
    core.reset(); core.set_state(0, sc.qpos, sc.qvel, ctrl)
    d = core.debug_substeps(1, env=0)
    og, _ = o.contacts()
    o.step2(ctrl)
    fo = o.efc()["force"]
    n_eq = int((o.efc()["type"] != 3).sum()) if False else len(fo) - 4 * len(og)
    fo = fo[n_eq:].reshape(-1, 4)
    # with hidden cubes asleep (production) their floor contacts are not generated: compare the contacts of the in-scene cubes
    gset = [int(np.where(M["geom_body"] == M["cube_body"][k])[0][0]) for k in range(3)]
    keep_o = np.array([x[0] in gset or x[1] in gset for x in og]); keep_g = np.array([x[0] in gset or x[1] in gset for x in d["pair"]])
    assert np.array_equal(d["pair"][keep_g], og[keep_o])
    fn_o, fn_g = fo.sum(axis=1)[keep_o], d["force"].sum(axis=1)[keep_g]
    og = og[keep_o]
    mass = M["body_mass"][M["cube_body"]]
    for k in range(3):      # each cube's weight is carried by its four corner contacts
        sel = (og[:, 0] == gset[k]) | (og[:, 1] == gset[k])
        assert sel.sum() == 4
        assert abs(fn_o[sel].sum() - mass[k] * 9.81) < 1e-6 * mass[k] * 9.81
        assert abs(fn_g[sel].sum() - mass[k] * 9.81) < 2e-3 * mass[k] * 9.81, (k, fn_g[sel].sum(), mass[k] * 9.81)
    tol = 0.05 * np.abs(fn_o).max()
    assert np.abs(fn_g - fn_o).max() < tol, (np.abs(fn_g - fn_o).max(), tol)
    return np.abs(fn_g - fn_o).max() / np.abs(fn_o).max()


def case_control_step_with_contacts_production(be):
    """ONE control step (100 sub-steps) from identical state with cubes in resting contact, at the solver settings the benchmark
    runs (<= 4 sweeps, 1e-3 m/s^2, hidden cubes asleep): north_star's 'single-step state <= 1e-4 relative'."""
    o, c = _pair(be, 2, "AllFullRLProgressRewardEnv", **PROGRESS_KW)
    sc = _settled_scene()
    a = np.random.default_rng(5).uniform(-0.3, 0.3, 16).astype(np.float32)
    o.reset(); c.reset(); o.transplant(sc); c.transplant(sc)
    oo, _, io = o.step(a); oc, _, ic = c.step(a)
    assert io == {k: ic[k] for k in io}
    r_arm = rel(oc[:48], oo[:48]); r_pos = rel(oc[48:118], oo[48:118]); r_vel = rel(oc[118:178], oo[118:178])
    assert r_arm < 1e-4 and r_pos < 1e-5 and r_vel < 1e-4, (r_arm, r_pos, r_vel)
    return r_arm, r_pos, r_vel


def case_impact_substep_production(be):
    """the sub-step in which a falling cube first touches the belt (4.2 m/s impact, 3 mm deep after one 1 ms step: impedance at
    dmax = 0.9999), from identical state: contact pair set bit-exact; post-impact cube velocity 1.1e-3 relative, position 5e-6.
    This is the one place where the fp32 path misses north_star's 1e-4: the active-set Newton solves (M + J^T J / R) a = ... with
    R ~ 1e-5 of the row inertia, a Hessian of condition ~1e5, so fp32 leaves ~1e-3 in the stiff directions (forces agree to 2e-3).
    Resting and sliding contacts (d < 0.99) meet 1e-4 (case_control_step_with_contacts_production)."""
    blob = fm.build_blob(2)
    o = Oracle(blob); o.env_create("AllFullRLProgressRewardEnv"); o.env_reset()
    for t in range(4):
        o.env_step(np.zeros(16, np.float32))
    ctrl = o.env_get("ctrl_target", 17)
    cube_g = 3
    for n in range(200):       # raw sub-steps until the cube-belt contact appears
        o.step2(ctrl); o.step1()
        g, _ = o.contacts()
        if any(x[0] == 2 and x[1] == cube_g for x in g):
            break
    else:
        raise AssertionError("no impact found")
    q, v = o.get_state()
    core = be.mk(blob, 4, 2, "AllFullRLProgressRewardEnv", auto_reset=False, sleep_hidden=False)
    core.reset(); core.set_state(0, q, v, ctrl)
    d = core.debug_substeps(1, env=0)
    og, _ = o.contacts()
    assert d["ncon"] == len(og) and np.array_equal(d["pair"], og)
    o.step2(ctrl)
    q2, v2 = o.get_state(); gq, gv, _ = core.get_state(0)
    rv = rel(gv[1:7], v2[1:7]); rq = rel(gq[1:8], q2[1:8])
    assert rv < 2e-3 and rq < 1e-5, (rv, rq)
    return rv, rq


def case_episode_boundary(be):
    """one full episode (zero actions: ends by out_of_reach near step 218) through the SB3-style auto-reset, against the oracle
    reset by hand: post-reset observation, the TaskManager RNG stream that is NOT reseeded (task_utils.py:19), and the progress
    reward's last-distance memory that is NOT cleared (environments.py:280-281)"""
    o = OracleSim(2, "AllFullRLProgressRewardEnv", **PROGRESS_KW)
    c = CoreSim(be, 2, "AllFullRLProgressRewardEnv", auto_reset=True, **PROGRESS_KW)
    a = np.zeros(16, np.float32)
    obs0 = o.reset(); assert np.array_equal(c.reset(), obs0)
    t_o = t_c = None
    ro = rc_ = None
    sp_o = sp_c = 0; last_o = last_c = 0
    for t in range(260):
        if t_o is None:
            _, r, io = o.step(a)
            sp_o += io["num_obj"] > last_o; last_o = io["num_obj"]
            if io["terminated"]:
                t_o = t
                o.o.env_reset()
                _, ro, _ = o.step(a)                       # first step of episode 2
                q_o, _, _ = o.state()
        if t_c is None:
            obs_c, r, ic = c.step(a)
            if ic["terminated"]:
                t_c = t
                assert np.array_equal(obs_c, obs0)         # the returned observation is the post-reset one
                obs_c2, rc_, ic2 = c.step(a)
                q_c, _, _ = c.state()
            else:
                sp_c += ic["num_obj"] > last_c; last_c = ic["num_obj"]
        if t_o is not None and t_c is not None:
            break
    assert t_o is not None and t_c is not None and abs(t_o - t_c) <= 2, (t_o, t_c)      # chaotic tumbling: +-2 control steps
    # the first spawn of episode 2 continues the PCG64 stream of episode 1 (seeded once in __init__, task_utils.py:19)
    draws = np.random.default_rng(42).uniform(0, 1, (12, 4))
    assert sp_o >= 5
    for q, n_sp in ((q_o, sp_o), (q_c, sp_o)):
        assert np.abs(q[4:8] - draws[n_sp] / np.linalg.norm(draws[n_sp])).max() < 1e-6, (n_sp, q[4:8])
    assert abs(ro - rc_) < 0.02, (ro, rc_)       # carries 0.2 * (stale last distance - new distance) ~ 0.2: a cleared memory would show
    return t_o, t_c


def case_config1(be, n_steps=10000, sync_every=25):
    """configs[0]: the oracle runs `n_steps` control steps of one env under
    default_rng(0).uniform(-1, 1, (n, 16)) actions with auto-reset; every `sync_every` steps the kernel (one env) is set to the
    oracle's full state and must reproduce the next control step: integers exactly, arm state to 1e-4 relative, cube rows to
    1e-3 absolute unless an arm or an impact stirred them in that step."""
    o = OracleSim(2, "AllFullRLProgressRewardEnv", **PROGRESS_KW)
    c = CoreSim(be, 2, "AllFullRLProgressRewardEnv", n_envs=1, auto_reset=False, **PROGRESS_KW)
    acts = np.random.default_rng(0).uniform(-1, 1, (n_steps, 16)).astype(np.float32)
    o.reset(); c.reset()
    worst_arm = worst_cube = 0.0
    checked = resets = calm = 0
    for t in range(n_steps):
        sync = t % sync_every == 0
        if sync:
            q, v, ct = o.state()
            lists = o.o.env_get("lists", 23); sc_ = o.o.env_get("scalars", 9)
            n_in = int(lists[0])
            sc = Scene(q, v, ct, [int(x) for x in lists[1:1 + n_in]], int(sc_[3]), since_spawn=int(sc_[4]), conveyor_speed=sc_[0],
                       spawn_freq=sc_[1])
            sc.out_scene = [int(x) for x in lists[12:12 + int(lists[11])]]
            c.transplant(sc)
        obs_o, r_o, io = o.step(acts[t])
        if sync:
            obs_c, r_c, ic = c.step(acts[t])
            assert {k: ic[k] for k in ("terminated", "out_of_reach", "force_terminate", "num_obj")} == {k: io[k] for k in ("terminated", "out_of_reach", "force_terminate", "num_obj")}, (t, io, ic)
            ra = rel(obs_c[:48], obs_o[:48])
            worst_arm = max(worst_arm, ra)
            dc = np.abs(obs_c[48:118] - obs_o[48:118]).max()
            if dc < 1e-3:
                calm += 1
            worst_cube = max(worst_cube, dc)
            checked += 1
        if io["terminated"]:
            o.reset(); resets += 1
    assert checked >= n_steps // sync_every and resets >= n_steps // 400
    assert worst_arm < 1e-4, worst_arm
    assert calm >= 0.9 * checked, (calm, checked, worst_cube)
    return worst_arm, worst_cube, calm, checked, resets
