"""Oracle statistics for the statistical parity test (BASELINE.json: "mean collected objects under a fixed policy must fall
within the reference's 95% confidence interval").  The reference cannot run here, so the interval is produced by N episodes of
the fp64 oracle under the IK base policy (FactoryManipulationEnv, 2 arms), TaskManager seeds 42..42+N-1,
fixed horizon H control steps (episodes that terminate earlier keep their score).  Writes tests/golden/ik_policy_stats.json.
    python tools/make_policy_stats.py [N] [H]
"""
import json, os, sys
import multiprocessing as mp
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def episode(seed_h):
    seed, H = seed_h
    from factory_marl_b200 import model as fm
    from oracle.oracle_py import Oracle
    o = Oracle(fm.build_blob(2)); o.env_create("FactoryManipulationEnv", seed=seed); o.env_reset()
    scores, length, reason = (0, 0), H, "horizon"
    for t in range(H):
        _, _, info = o.env_step(np.zeros(0, np.float32))
        scores = info["scores"]
        if info["terminated"]:
            length, reason = t + 1, ("force" if info["force_terminate"] else "out_of_reach")
            break
    return seed, scores[0], scores[1], length, reason


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        res = pool.map(episode, [(42 + i, H) for i in range(N)])
    s0 = np.array([r[1] for r in res], float); s1 = np.array([r[2] for r in res], float); ln = np.array([r[3] for r in res], float)
    tot = s0 + s1
    def ci(x):
        return [float(x.mean() - 1.96 * x.std(ddof=1) / np.sqrt(len(x))), float(x.mean() + 1.96 * x.std(ddof=1) / np.sqrt(len(x)))]
    out = {"policy": "IK base policy (FactoryManipulationEnv), 2 arms", "episodes": N, "horizon": H, "seeds": [42, 42 + N - 1],
           "mean_score": [float(s0.mean()), float(s1.mean())], "mean_total": float(tot.mean()), "std_total": float(tot.std(ddof=1)),
           "ci95_total": ci(tot), "ci95_bucket0": ci(s0), "ci95_bucket1": ci(s1), "mean_length": float(ln.mean()), "ci95_length": ci(ln),
           "reasons": {k: int(sum(1 for r in res if r[4] == k)) for k in ("horizon", "force", "out_of_reach")},
           "per_episode": [[int(r[0]), int(r[1]), int(r[2]), int(r[3])] for r in res]}
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ik_policy_stats.json"), "w"), indent=1)
    print({k: v for k, v in out.items() if k != "per_episode"})
