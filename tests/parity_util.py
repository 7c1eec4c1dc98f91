"""Shared by the CPU logic tests (host-compiled kernel build) and the GPU tests (CUDA library through the C-ABI)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PROGRESS_KW = dict(gripper_to_closest_cube_reward_factor=0.2, closest_cube_to_bucket_reward_factor=0.4, base_reward=0.4)
ENV_CLASS_NAMES = ["FactoryManipulationEnv", "SingleFullRLProgressRewardEnv", "SingleDeltaProgressRewardEnv",
                   "AllFullRLProgressRewardEnv", "AllDeltaProgressRewardEnv", "PauseIKToggleEnv", "BackupIKToggleEnv"]


def load_golden(cls, n_arms):
    return np.load(os.path.join(GOLDEN, f"rollout_{cls}_{n_arms}arms.npz"))


def rollout_against_golden(core, g, to_np, from_np, n_envs):
    """Drive `core` (FmCore) with the golden action sequence in every env; returns per-step max errors of env 0 and
    checks that all envs (same seed) agree bit for bit."""
    obs = to_np(core.reset())
    errs = [np.abs(obs[0] - g["obs"][0]).max()]
    rew_err, term_ok, score_ok = [], [], []
    for t in range(len(g["actions"])):
        a = np.tile(g["actions"][t][None], (n_envs, 1))
        if a.shape[1] == 0:
            a = np.zeros((n_envs, 1), np.float32)
        core.env_step(from_np(a))
        obs = to_np(core.obs)
        assert np.array_equal(obs, np.tile(obs[0][None], (n_envs, 1))), "envs with identical seed/actions diverged"
        errs.append(np.abs(obs[0] - g["obs"][t + 1]).max())
        rew_err.append(abs(float(to_np(core.reward)[0]) - g["reward"][t]))
        term_ok.append(bool(to_np(core.terminated)[0]) == bool(g["terminated"][t]))
        score_ok.append(tuple(to_np(core.scores)[0]) == tuple(g["scores"][t]))
    return np.array(errs), np.array(rew_err), term_ok, score_ok
