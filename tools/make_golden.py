"""Generate tests/golden/*.npz from the fp64 oracle (oracle/): seeded rollouts of every env class.

The reference itself cannot run in this image (MuJoCo / dm_control / SB3 absent), so these vectors pin the ORACLE
(a restatement) against regressions and give the `-m gpu` tests fixtures that travel to the GPU box.  They are
not outputs of the reference; the reference-held fixtures are tests/golden/ref_* (tools/make_ref_fixtures.py).
    python tools/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from factory_marl_b200 import model as fm  # noqa: E402
from oracle.oracle_py import Oracle, ENV_CLASSES  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
KW = dict(gripper_to_closest_cube_reward_factor=0.2, closest_cube_to_bucket_reward_factor=0.4, base_reward=0.4)


def actions_for(cls, n_arms, steps, rng):
    if cls in ("PauseIKToggleEnv", "BackupIKToggleEnv"):
        return rng.integers(0, 2, (steps, n_arms)).astype(np.int32)
    if cls == "FactoryManipulationEnv":
        return np.zeros((steps, 0), np.float32)
    d = 8 if cls.startswith("Single") else 8 * n_arms
    return rng.uniform(-1, 1, (steps, d)).astype(np.float32)


def main():
    os.makedirs(OUT, exist_ok=True)
    arms = tuple(int(a) for a in sys.argv[1:]) or (2, 4, 6)       # python tools/make_golden.py 6: only the six-arm scene
    for n_arms in arms:
        blob = fm.build_blob(n_arms)
        for cls in ENV_CLASSES:
            steps = 12 if n_arms == 2 else 6
            rng = np.random.default_rng(1234 + n_arms)
            o = Oracle(blob)
            kw = KW if "Progress" in cls else {}
            o.env_create(cls, **kw)
            acts = actions_for(cls, n_arms, steps, rng)
            obs = [o.env_reset()]
            rew, term, scores, qpos, qvel = [], [], [], [], []
            for t in range(steps):
                ob, r, info = o.env_step(acts[t])
                obs.append(ob); rew.append(r); term.append(info["terminated"]); scores.append(info["scores"])
                q, v = o.get_state(); qpos.append(q); qvel.append(v)
            np.savez_compressed(os.path.join(OUT, f"rollout_{cls}_{n_arms}arms.npz"), actions=acts, obs=np.array(obs, np.float32),
                                reward=np.array(rew), terminated=np.array(term), scores=np.array(scores, np.int32),
                                qpos=np.array(qpos), qvel=np.array(qvel))
            print(cls, n_arms, "ok", np.array(obs).shape)
    if 2 not in arms:
        return
    # known-answer scalars
    blob = fm.build_blob(2)
    o = Oracle(blob); o.env_create("AllFullRLProgressRewardEnv")
    o.env_reset()
    spawn_steps, nobj = [], []
    for t in range(120):
        _, _, info = o.env_step(np.zeros(16, np.float32))
        nobj.append(info["num_obj"]); spawn_steps.append(int(1.0 / (0.1 * o.env_get("scalars", 9)[1])))
    q, v = o.get_state()
    np.savez_compressed(os.path.join(OUT, "known_answers.npz"), nobj=np.array(nobj), spawn_steps=np.array(spawn_steps), qpos120=q, qvel120=v)


if __name__ == "__main__":
    main()
