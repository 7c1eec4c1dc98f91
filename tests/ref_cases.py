"""Parity against REFERENCE-HELD data: the real MuJoCo observations (`_last_obs`) and Monitor episodes (`ep_info_buffer`)
inside /root/reference/runs/*.zip, decoded by tools/make_ref_fixtures.py into tests/golden/ref_*.  These are outputs of the
reference's own env (dm_control + MuJoCo 3.2.0), not of this repo's oracle.  Each case takes a `mk(n_arms, cls, **kw)` factory
returning a tests/sim_adapters.py sim (OracleSim, or CoreSim over the host-compiled kernel build / the CUDA library)."""
import json
import os

import numpy as np

from factory_marl_b200 import model as fm
from tests.parity_util import GOLDEN, PROGRESS_KW
from tests.sim_adapters import Scene, arm_qadr, arm_vadr, hidden_qpos

RUNS = json.load(open(os.path.join(GOLDEN, "ref_runs.json")))
LAST_OBS = np.load(os.path.join(GOLDEN, "ref_last_obs.npz"))
N_ARMS = {r: RUNS[r]["env_kwargs"]["num_arms"] for r in ("r666unuv", "rk5rxnav", "xfwgqibb", "y6lp1j7k")}
BELT_TOP, TABLE_TOP = 1.09, 1.0
DEFAULT_POSE = np.array([-0.5, -0.5, 0.0, 1.0, 0.0, -1.6, 0.0, 0.06])


def ref_rows(run):
    """(env, row, pose7, vel6) of every cube row of the run's 8 reference observations."""
    na = N_ARMS[run]; nb = 24 * na
    for e in range(8):
        ob = LAST_OBS[run][e]
        poses = ob[nb:nb + 70].reshape(10, 7); vels = ob[nb + 70:nb + 130].reshape(10, 6)
        for k in range(10):
            if np.abs(poses[k]).sum() > 0:
                yield e, k, poses[k], vels[k]


def is_resting(pose, vel, vz_tol=1e-6):
    """at rest on a horizontal surface: no vertical / lateral / angular velocity and one cube face exactly horizontal"""
    if not (abs(vel[2]) < vz_tol and np.abs(vel[3:]).max() < 1e-5 and abs(vel[0]) < 1e-6 and 1.02 < pose[2] < 1.2):
        return False
    q = np.asarray(pose[3:7], float); q = q / np.linalg.norm(q)
    return np.abs(fm.q2mat(q)[2]).max() > 1 - 1e-6


def is_falling(pose, vel):
    return abs(pose[0]) < 1e-9 and abs(pose[1] - 1) < 1e-9 and abs(vel[0]) < 1e-9 and abs(vel[1]) < 1e-9 and np.abs(vel[3:]).max() < 1e-9


def cube_id_resting(sizes, z):
    """resting cube -> (id, support top) from its height: z = top + s_k - depth"""
    best = None
    for top in (BELT_TOP, TABLE_TOP):
        k = int(np.argmin(np.abs(z - (top + sizes))))
        err = abs(z - (top + sizes[k]))
        if best is None or err < best[2]:
            best = (k, top, err)
    return best


# ---------------------------------------------------------------------------------------------------------------------
def case_ref_free_fall(mk):
    """Every falling cube row of the reference observations ((x, y) = spawn point (0, 1), task_utils.py:44) is a point of one of
    two chains: the first spawn of an episode inherits v_z = -0.981 from 0.1 s of hidden free fall, later spawns start at rest
    (SURVEY.md A.3).  Our env must pass through the same (z, v_z) float32 pairs."""
    sim = mk(2, "AllFullRLProgressRewardEnv", **PROGRESS_KW)
    sim.reset()
    ours = []
    for t in range(53):
        obs, _, _ = sim.step(np.zeros(16, np.float32))
        poses = obs[48:118].reshape(10, 7); vels = obs[118:178].reshape(10, 6)
        for k in range(10):
            if abs(poses[k, 0]) < 1e-4 and abs(poses[k, 1] - 1) < 1e-4:
                ours.append((float(poses[k, 2]), float(vels[k, 2])))
    ours = np.array(ours)
    n = 0
    for run in N_ARMS:
        for e, k, pose, vel in ref_rows(run):
            if is_falling(pose, vel):
                d = np.abs(ours - np.array([pose[2], vel[2]])).max(axis=1).min()
                # fp64 oracle: exact to float32 print precision; fp32 kernel: 400 accumulated sub-steps, and a slept hidden cube
                # keeps < 1e-5 m/s (DESIGN.md deviation 5)
                assert d < (1e-6 if sim.kind == "oracle" else 3e-5), (run, e, pose[2], vel[2], d)
                n += 1
    assert n >= 8, n
    return n


def case_ref_rest_depth_and_belt_lag(mk, n_steps=262):
    """Resting cubes in the reference: z = support top + half size - depth (2.28e-6 on the belt for every mass) and v_y equal to the
    belt's, which lags the commanded -conveyor_speed by 5.8e-5 (velocity servo kv 1e4 on 1000 kg, speed raised 1e-4 per step)."""
    sim = mk(2, "AllFullRLProgressRewardEnv", **PROGRESS_KW)
    M = sim.M; sizes = M["cube_size"]
    fp64 = sim.kind == "oracle"
    sim.reset()
    belt_v, depth, ridden = [], [], 0
    for t in range(n_steps):
        obs, _, info = sim.step(np.zeros(16, np.float32))
        q, v, _ = sim.state()
        belt_v.append(v[0])
        poses = obs[48:118].reshape(10, 7); vels = obs[118:178].reshape(10, 6)
        if not info["terminated"] and 100 <= t < 200:
            for k in range(10):
                # fp32 kernel: z toggles between the two floats around the equilibrium (|v_z| ~ 2e-6), v_y within a few ulp of the belt's
                if np.abs(poses[k]).sum() > 0 and is_resting(poses[k], vels[k], 1e-6 if fp64 else 1e-5):
                    cid, top, err = cube_id_resting(sizes, poses[k, 2])
                    if top != BELT_TOP or abs(vels[k, 1] - np.float32(v[0])) > (1e-7 if fp64 else 5e-7):
                        continue                                          # still sliding after its landing
                    depth.append((top + sizes[cid]) - q[3 + 7 * cid])     # the cube rides the belt
                    ridden += 1
    belt_v = np.array(belt_v); depth = np.array(depth)
    assert ridden > 100
    n = lag_ok = 0
    for run in N_ARMS:
        for e, k, pose, vel in ref_rows(run):
            if not is_resting(pose, vel):
                continue
            cid, top, err = cube_id_resting(sizes, float(pose[2]))
            if top != BELT_TOP or abs(pose[0]) + 1.4143 * sizes[cid] > 0.3:
                continue                                              # on the table, or a corner may overhang the belt edge
            d_ref = (top + sizes[cid]) - float(pose[2])               # float32 quantisation of z: +-6e-8
            assert abs(d_ref - depth.mean()) < 2e-7, (run, e, cid, d_ref, depth.mean())
            # belt velocity: cube landings disturb the 1000 kg belt by a few 1e-7 m/s, an arm pressing on it by more (one scene)
            dv = np.abs(belt_v - float(vel[1])).min()
            assert dv < 1e-6 or N_ARMS[run] == 4, (run, e, float(vel[1]))   # four arms lean on the belt now and then
            lag_ok += dv < 1e-6
            n += 1
    assert n >= 80 and lag_ok >= 60, (n, lag_ok)
    assert abs(depth.mean() - 2.28e-6) < 5e-8 and depth.std() < (2e-8 if fp64 else 5e-8), (depth.mean(), depth.std())
    return depth.mean()


def case_ref_arm_droop(mk):
    """xfwgqibb / y6lp1j7k env 0, arm 0: an idle arm commanded to the IK default pose hangs at joint 2 = -0.528117 etc. under
    gravity against the kp = 2000 position servos (pins link masses, COMs, FK and the actuator model)."""
    sim = mk(4, "BackupIKToggleEnv")
    sim.reset()
    for t in range(40):
        obs, _, _ = sim.step(np.zeros(4, np.int32))
    for run in ("xfwgqibb", "y6lp1j7k"):
        ref = LAST_OBS[run][0]
        assert np.allclose(ref[16:24], DEFAULT_POSE, atol=1e-7)
        assert np.abs(obs[0:8] - ref[0:8]).max() < 2e-6, (run, obs[0:8], ref[0:8])
    assert np.abs(obs[16:24] - DEFAULT_POSE).max() < 1e-6
    assert np.abs(obs[226:234] - DEFAULT_POSE).max() < 1e-6        # IDLE arm: proposal = default pose (toggle obs tail)
    return obs[0:8]


def scene_from_ref_obs(M, run, e):
    """Rebuild a full env state from one reference observation whose cubes are all identifiable: resting ones by their
    height (the half sizes differ), falling ones take a free id.  Returns (Scene, ids in reference row order) or None."""
    na = N_ARMS[run]; nb = 24 * na
    ob = LAST_OBS[run][e].astype(np.float64)
    sizes = M["cube_size"]
    q = hidden_qpos(M, na); v = np.zeros(61 + 9 * na); ct = np.zeros(1 + 8 * na)
    poses = ob[nb:nb + 70].reshape(10, 7); vels = ob[nb + 70:nb + 130].reshape(10, 6)
    ids, vy = [], None
    rows = [k for k in range(10) if np.abs(poses[k]).sum() > 0]
    for k in rows:
        if is_resting(poses[k], vels[k]):
            cid, top, err = cube_id_resting(sizes, poses[k, 2])
            if err > 5e-6 or cid in ids:
                return None
            ids.append(cid)
            if top == BELT_TOP:
                vy = vels[k, 1]
        elif is_falling(poses[k], vels[k]):
            ids.append(-1)
        else:
            return None
    if vy is None:
        return None
    free = [k for k in range(10) if k not in ids]
    ids = [free.pop() if i < 0 else i for i in ids]
    for k, cid in zip(rows, ids):
        qq = poses[k].copy(); qq[3:] /= np.linalg.norm(qq[3:])
        q[1 + 7 * cid:8 + 7 * cid] = qq
        v[1 + 6 * cid:7 + 6 * cid] = vels[k]
    # after n steps the belt has tracked 0.1 + 1e-4 (n - 1) up to its lag of 5.8e-5; conveyor_speed is 0.1 + 1e-4 n (base_env.py:269)
    step = int(round((-vy - 0.1 + 5.82e-5) / 1e-4)) + 1
    v[0] = vy
    ct[0] = -(0.1 + 1e-4 * (step - 1))
    for a in range(na):
        blk = ob[24 * a:24 * a + 24]
        for j in range(8):
            q[arm_qadr(M, a, j)] = blk[j]; v[arm_vadr(M, a, j)] = blk[8 + j]
        q[arm_qadr(M, a, 8)] = blk[7]; v[arm_vadr(M, a, 8)] = blk[15]      # plate equality (gripper.xml:46-49)
        ct[1 + 8 * a:9 + 8 * a] = blk[16:24]
    return Scene(q, v, ct, ids, step, n_arms=na), ids


def case_ref_scene_step(mk, run, min_envs=2):
    """Transplant reference observations (all cubes at rest or in free fall) and advance ONE control step: resting cubes keep the
    reference height to 2e-7 (belt and table), pick up the belt's next velocity, and the observation lists the cubes in the
    reference's row order (ascending x, zero padded; task_utils.py:62-77)."""
    na = N_ARMS[run]
    cls = RUNS[run]["env_class"]
    kw = PROGRESS_KW if "Progress" in cls else {}
    if cls == "AllDeltaProgressRewardEnv":
        cls = "AllFullRLProgressRewardEnv"          # same physics and observation; the action map is not what is tested here
    if cls == "PauseIKToggleEnv":
        cls = "BackupIKToggleEnv"                   # action 0 = default pose needs no hidden last_arm_actions
    M = fm.build_model(na)
    nb = 24 * na
    done = 0
    for e in range(8):
        built = scene_from_ref_obs(M, run, e)
        if built is None:
            continue
        sc, ids = built
        if len(ids) < 2:
            continue
        sim = mk(na, cls, **kw)
        sim.reset()
        sim.transplant(sc)
        ref = LAST_OBS[run][e]
        if cls == "BackupIKToggleEnv":
            act = np.zeros(na, np.int32)
        else:       # hold every arm at its filtered command: action = atanh of the normalised ctrl (environments.py:98-102)
            lo, hi = M["act_ctrlrange"][1:9, 0], M["act_ctrlrange"][1:9, 1]
            u = np.clip((sc.ctrl_target[1:].reshape(na, 8) - lo) / (hi - lo) * 2 - 1, -0.999, 0.999)
            act = np.arctanh(u).reshape(-1).astype(np.float32)
        obs, _, info = sim.step(act)
        poses = obs[nb:nb + 70].reshape(10, 7); vels = obs[nb + 70:nb + 130].reshape(10, 6)
        rposes = ref[nb:nb + 70].reshape(10, 7); rvels = ref[nb + 70:nb + 130].reshape(10, 6)
        n_ref = len(ids)
        spawned = info["num_obj"] - n_ref
        assert spawned in (0, 1)
        rows = [k for k in range(10) if np.abs(poses[k]).sum() > 0]
        keep = [k for k in range(n_ref) if not is_falling(rposes[k], rvels[k])]
        ours = []                                                         # our row of every resting reference row
        for kr in keep:
            cand = [k for k in rows if abs(poses[k, 0] - rposes[kr, 0]) < 2e-6 and abs(poses[k, 2] - rposes[kr, 2]) < 1e-5]
            assert len(cand) == 1, (run, e, kr, cand)
            ours.append(cand[0])
        assert ours == sorted(ours)                                       # same order as the reference rows
        assert np.all(np.diff(poses[rows, 0]) >= 0)                       # ascending x
        assert np.all(poses[info["num_obj"]:] == 0) and np.all(vels[info["num_obj"]:] == 0)   # zero padding
        for ko, kr in zip(ours, keep):
            assert abs(poses[ko, 0] - rposes[kr, 0]) < 2e-6, (run, e, ko)
            assert abs(poses[ko, 2] - rposes[kr, 2]) < 2e-7, (run, e, ko, poses[ko, 2], rposes[kr, 2])
            cid, top, _ = cube_id_resting(M["cube_size"], float(rposes[kr, 2]))
            if top == BELT_TOP:
                vtol = (1e-6 if sim.kind == "oracle" else 3e-6) if len(keep) == n_ref else 2e-5   # a cube landing during this step jolts the belt
                assert abs(vels[ko, 1] - (rvels[kr, 1] - 1e-4)) < vtol, (run, e, vels[ko, 1], rvels[kr, 1])
                assert abs(poses[ko, 1] - (rposes[kr, 1] + 0.1 * rvels[kr, 1])) < 2e-5
            else:
                assert np.abs(vels[ko]).max() < 1e-6
        done += 1
    assert done >= min_envs, (run, done)
    return done


def case_ref_ik_proposals(n_min=15):
    """IK proposals in the toggling runs' observation tails (environments.py:574-576).  (1) The forward kinematics of every
    unclipped proposal puts the gripper site exactly pre_grasp_height (0.15) or grasp_offset (0.04) above the target cube plus
    0.3 s of its velocity, or at the bucket's pre-release point (ik_policy.py:59-69,214-254).  (2) The oracle's restatement of
    dm_control's qpos_from_site_pose, started from the observed arm pose, lands on the reference's joint values."""
    from oracle.oracle_py import Oracle
    M = fm.build_model(4)
    o = Oracle(fm.build_blob(4)); o.env_create("PauseIKToggleEnv"); o.env_reset()
    lo, hi = M["act_ctrlrange"][1:8, 0], M["act_ctrlrange"][1:8, 1]
    bucket = {1: np.array([-0.9, -0.3, 1.3]), 0: np.array([0.9, -0.3, 1.3])}

    def site_pose(a, q7):
        q = fm.qpos0(M)
        for k in range(7):
            q[arm_qadr(M, a, k)] = q7[k]
        xp, xq = fm.forward_kinematics(M, q)
        b = M["site_body"][a]
        return xp[b] + fm.q2mat(xq[b]) @ M["site_pos"][a], xq[b]

    n_fk = n_ik = 0
    miss = []
    for run in ("xfwgqibb", "y6lp1j7k"):
        for e in range(8):
            ob = LAST_OBS[run][e]
            cubes = ob[96:166].reshape(10, 7); vels = ob[166:226].reshape(10, 6); ik = ob[226:].reshape(4, 8)
            for a in range(4):
                p7 = ik[a, :7].astype(float)
                if np.allclose(p7, DEFAULT_POSE[:7]):
                    continue
                if np.any(np.isclose(p7, lo, atol=1e-6) | np.isclose(p7, hi, atol=1e-6)):
                    continue                                            # clipped to ctrlrange: the IK solution is not recoverable
                tp, tq = site_pose(a, p7)
                offs = [tp - cubes[k, :3] - np.r_[vels[k, :2] * 0.3, 0] for k in range(10) if np.abs(cubes[k]).sum() > 0]
                hit = any(np.abs(d[:2]).max() < 2e-6 and min(abs(d[2] - 0.15), abs(d[2] - 0.04)) < 2e-6 for d in offs)
                hit = hit or np.abs(tp - bucket[a % 2]).max() < 2e-6
                if ik[a, 7] > 0.03 and not hit:
                    continue        # POST_GRASP / GO_TO_RELEASE interpolate from a hidden start point; stale last_ctrl after an IK failure
                if hit:
                    n_fk += 1
                ok, qs, steps = o.ik_solve(a, tp, tq, ob[24 * a:24 * a + 7].astype(float))
                if hit:
                    if ok and np.abs(qs - p7).max() < 5e-7:
                        n_ik += 1
                    else:
                        miss.append((run, e, a, float(np.abs(qs - p7).max())))
    # one proposal (y6lp1j7k env 3 arm 2, gripper closed at the pre-release point) sits on another null-space branch than a
    # solve from the observed pose: most likely the reference's solve failed there and `last_ctrl` was returned (ik_policy.py:266)
    assert n_fk >= n_min and n_ik >= n_min and len(miss) <= 1, (n_fk, n_ik, miss)
    return n_fk, n_ik
