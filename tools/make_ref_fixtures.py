"""Decode the reference-held data inside /root/reference/runs/*.zip into small committed fixtures (numpy + stdlib + torch.load).

Every SB3 checkpoint of the reference carries, next to `policy.pth`, a `data` JSON whose pickled fields were produced by the
reference's own MuJoCo env during the last training rollout:
  * `ep_info_buffer`  — the last 100 Monitor episodes (return `r`, length `l`) under the near-final stochastic policy
                        (src/learning.py:98-118; for the toggling classes the return IS the collected-object count,
                        src/environments.py:148);
  * `_last_obs`       — 8 real MuJoCo observations mid-episode (one per SubprocVecEnv worker).
These are the only outputs of the real reference that exist; tests/test_ref_fixtures.py pins the oracle and the CUDA path on
them.  Writes tests/golden/ref_runs.json (episodes + 95 % CIs + env kwargs), tests/golden/ref_last_obs.npz and
tests/golden/ref_policy_<run>.npz (the 13 state-dict tensors of policy.pth, fp32).
    python tools/make_ref_fixtures.py            # needs /root/reference (not present on the GPU box)
"""
import base64
import io
import json
import os
import pickle
import sys
import warnings
import zipfile

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REF = "/root/reference/runs"
RUNS = ["r666unuv", "rk5rxnav", "xfwgqibb", "y6lp1j7k"]


def _unpickle(field):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return pickle.loads(base64.b64decode(field[":serialized:"]))


def ci95(x):
    x = np.asarray(x, float)
    h = 1.96 * x.std(ddof=1) / np.sqrt(len(x))
    return [float(x.mean() - h), float(x.mean() + h)]


def main():
    import torch
    out, last_obs = {}, {}
    for run in RUNS:
        cfg = json.load(open(f"{REF}/{run}.json"))
        z = zipfile.ZipFile(f"{REF}/{run}.zip")
        data = json.loads(z.read("data"))
        eps = _unpickle(data["ep_info_buffer"])
        r = [float(e["r"]) for e in eps]
        ln = [int(e["l"]) for e in eps]
        kw = dict(cfg["env_kwargs"])
        kw.pop("render_mode", None)
        out[run] = {"env_class": cfg["env_class"], "env_kwargs": kw, "notes": cfg.get("notes", ""),
                    "n_episodes": len(eps), "returns": r, "lengths": ln,
                    "mean_return": float(np.mean(r)), "ci95_return": ci95(r),
                    "mean_length": float(np.mean(ln)), "ci95_length": ci95(ln),
                    "num_timesteps": int(data["num_timesteps"]), "n_envs": int(data["n_envs"]),
                    "source": f"runs/{run}.zip:data[ep_info_buffer], runs/{run}.json"}
        last_obs[run] = np.asarray(_unpickle(data["_last_obs"]), np.float32)
        sd = torch.load(io.BytesIO(z.read("policy.pth")), weights_only=True, map_location="cpu")
        np.savez(os.path.join(ROOT, "tests", "golden", f"ref_policy_{run}.npz"),
                 **{k: v.numpy().astype(np.float32) for k, v in sd.items()})
    # published task-quality means (report/report.tex:273-297), no variance published
    out["report_tex"] = {
        "ik_base_2arms": {"scores": [1.65, 1.17], "length": 208.8, "line": 281},
        "ik_base_4arms": {"scores": [1.18, 1.16], "length": 119.74, "line": 294},
        "pause_4arms": {"scores": [2.01, 1.5], "length": 173.04, "line": 292},
        "backup_4arms": {"scores": [0.0, 2.12], "length": 216.74, "line": 293},
        "allfull_2arms": {"scores": [0, 0], "length": 219.9, "line": 279},
        "alldelta_2arms": {"scores": [0, 0], "length": 220.5, "line": 280}}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "ref_runs.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_last_obs.npz"), **last_obs)
    for run in RUNS:
        print(run, out[run]["env_class"], out[run]["env_kwargs"]["num_arms"], "return %.3f CI %s length %.1f CI %s" % (
            out[run]["mean_return"], np.round(out[run]["ci95_return"], 3), out[run]["mean_length"], np.round(out[run]["ci95_length"], 1)))


if __name__ == "__main__":
    sys.exit(main())
