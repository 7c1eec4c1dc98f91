"""Episode statistics of the fp64 oracle under the reference's own trained policies (tests/golden/ref_policy_<run>.npz, made by
tools/make_ref_fixtures.py from runs/<run>.zip:policy.pth) or under the IK base policy, in the protocol that produced the
reference's numbers: stochastic actions (src/visualisation.py:65, `deterministic` not passed; ep_info_buffer comes from
training rollouts), uncapped episodes.  Compared with tests/golden/ref_runs.json by tests/test_ref_fixtures.py.
    python tools/oracle_policy_stats.py <run|ik2|ik4> [N episodes] [max steps] [--write]
"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


class NumpyPolicy:
    """SB3 ActorCriticPolicy (MlpPolicy, net_arch [128,128], tanh) forward + sampling, fp32 numpy."""

    def __init__(self, path):
        self.w = dict(np.load(path))
        self.discrete = "log_std" not in self.w

    def act(self, obs, rng, deterministic=False):
        w = self.w
        h = np.tanh(w["mlp_extractor.policy_net.0.weight"] @ obs + w["mlp_extractor.policy_net.0.bias"])
        h = np.tanh(w["mlp_extractor.policy_net.2.weight"] @ h + w["mlp_extractor.policy_net.2.bias"])
        out = w["action_net.weight"] @ h + w["action_net.bias"]
        if self.discrete:                         # MultiDiscrete([2]*n): independent 2-way categoricals
            lg = out.reshape(-1, 2)
            if deterministic:
                return lg.argmax(1).astype(np.int32)
            p1 = 1.0 / (1.0 + np.exp(lg[:, 0] - lg[:, 1]))
            return (rng.uniform(size=len(p1)) < p1).astype(np.int32)
        a = out if deterministic else out + np.exp(w["log_std"]) * rng.standard_normal(len(out)).astype(np.float32)
        return np.clip(a, -1, 1).astype(np.float32)   # SB3 clips Box actions before env.step


def episode(args):
    what, seed, max_steps = args
    from factory_marl_b200 import model as fm
    from oracle.oracle_py import Oracle
    runs = json.load(open(os.path.join(GOLD, "ref_runs.json")))
    if what in ("ik2", "ik4"):
        na, cls, kw, pol = int(what[2]), "FactoryManipulationEnv", {}, None
    else:
        r = runs[what]
        na, cls = r["env_kwargs"]["num_arms"], r["env_class"]
        kw = {k: v for k, v in r["env_kwargs"].items() if k not in ("seed", "num_arms", "max_num_objects")}
        pol = NumpyPolicy(os.path.join(GOLD, f"ref_policy_{what}.npz"))
    o = Oracle(fm.build_blob(na))
    o.env_create(cls, seed=seed, **kw)
    obs = o.env_reset()
    rng = np.random.default_rng(1000 + seed)
    ret, length, reason, scores = 0.0, max_steps, "horizon", (0, 0)
    for t in range(max_steps):
        a = pol.act(obs, rng) if pol is not None else np.zeros(0, np.float32)
        obs, rew, info = o.env_step(a)
        ret += rew
        scores = info["scores"]
        if info["terminated"]:
            length = t + 1
            reason = "force" if info["force_terminate"] else "out_of_reach"
            if info["force_terminate"] and info["out_of_reach"]:
                reason = "both"
            break
    return seed, ret, length, reason, scores[0], scores[1]


def main():
    what = sys.argv[1]
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        res = pool.map(episode, [(what, 42 + i, H) for i in range(N)], chunksize=1)
    ret = np.array([r[1] for r in res]); ln = np.array([r[2] for r in res], float)
    s0 = np.array([r[4] for r in res], float); s1 = np.array([r[5] for r in res], float)

    def ci(x):
        h = 1.96 * x.std(ddof=1) / np.sqrt(len(x))
        return [float(x.mean() - h), float(x.mean() + h)]
    out = {"what": what, "episodes": N, "mean_return": float(ret.mean()), "ci95_return": ci(ret), "mean_length": float(ln.mean()),
           "ci95_length": ci(ln), "mean_scores": [float(s0.mean()), float(s1.mean())], "mean_total_score": float((s0 + s1).mean()),
           "ci95_total_score": ci(s0 + s1),
           "reasons": {k: int(sum(1 for r in res if r[3] == k)) for k in ("horizon", "force", "out_of_reach", "both")},
           "lengths": [int(x) for x in ln], "returns": [float(x) for x in ret]}
    print(json.dumps({k: v for k, v in out.items() if k not in ("lengths", "returns")}))
    print("lengths", sorted(out["lengths"]))
    if "--write" in sys.argv:
        json.dump(out, open(os.path.join(GOLD, f"oracle_stats_{what}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
