"""GPU suite (-m gpu): the host-side mirror of the reference interface — VecEnv contract (lazy infos, Monitor / ep_score_history on the
device), per-env gymnasium view, SB3-layout checkpoints that can resume, trajectory export, two-rank PPO step."""
import io
import json
import os
import zipfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from tests.parity_util import GOLDEN, PROGRESS_KW  # noqa: E402


def _env(n=64, cls="AllFullRLProgressRewardEnv", **kw):
    from factory_marl_b200.vec_env import FactoryVecEnv
    return FactoryVecEnv(cls, num_envs=n, num_arms=2, seeds=list(range(42, 42 + n)), **(PROGRESS_KW if "Progress" in cls else {}), **kw)


def test_lazy_infos_and_monitor_match_a_python_monitor():
    """bang-bang actions crash the arms (200 N) within ~15 steps: many short episodes.  The device monitor's ep_score_history /
    ep_rew_mean / ep_len_mean must equal what a Python Monitor computes from the per-step infos."""
    env = _env(64)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    ret = np.zeros(64); ln = np.zeros(64, int); eps = []; hist = [[] for _ in range(64)]
    for t in range(60):
        a = (torch.randint(0, 2, (64, 16), device="cuda", generator=g).float() * 2 - 1) * 3
        obs, rew, done, infos = env.step(a)
        assert len(infos) == 64
        r = rew.cpu().numpy(); d = done.cpu().numpy()
        ret += r; ln += 1
        for e in range(64):
            info = infos[e]
            assert set(info) >= {"scores", "play_time", "conveyor_speed", "out_of_reach", "force_terminate"}      # base_env.py:274-280
            if d[e]:
                assert abs(info["episode"]["r"] - ret[e]) < 1e-3 * max(1, abs(ret[e])) and info["episode"]["l"] == ln[e]
                assert info["terminal_observation"].shape == (178,) and info["force_terminate"] or info["out_of_reach"]
                eps.append((ret[e], ln[e])); hist[e].append(info["scores"]); ret[e] = 0; ln[e] = 0
            else:
                assert "episode" not in info
    assert len(eps) > 64
    assert env.get_attr("ep_score_history") == hist
    m = env.monitor.metrics()
    assert m["episodes"] == len(eps)
    if all(hist):       # learning.py:64: the bucket means exist once no env has an empty history (all scores are 0 here)
        assert m["rollout/ep_score_mean_bucket0"] == 0.0 and m["rollout/ep_score_mean_bucket1"] == 0.0
    else:
        assert "rollout/ep_score_mean_bucket0" not in m
    assert abs(m["rollout/ep_len_mean"] - np.mean([l for _, l in eps])) < 5      # same population up to which 100 the ring kept
    _, _, _, unread = env.step(a)
    env.step(a)
    with pytest.raises(RuntimeError):      # lazy infos are built from the step's device outputs: they must be read before the next step
        unread[0]


def test_monitor_bucket_means_reproduce_the_reference_indexing():
    """learning.py:64-67 averages history[-t] for t in range(min(100, len)): t = 0 is the FIRST episode.  Synthetic histories."""
    from factory_marl_b200.vec_env import EpisodeMonitor
    mon = EpisodeMonitor(3, torch.device("cuda:0"))
    rng = np.random.default_rng(0)
    hist = [[list(map(int, rng.integers(0, 5, 2))) for _ in range(n)] for n in (1, 7, 130)]
    for step in range(130):
        term = torch.tensor([1 if step < len(h) else 0 for h in hist], dtype=torch.uint8, device="cuda")
        ep = torch.zeros((3, 4), device="cuda")
        for e, h in enumerate(hist):
            if step < len(h):
                ep[e, 2], ep[e, 3] = h[step][0], h[step][1]
        mon.update(term, ep)
    m = mon.metrics()
    for b in range(2):
        vals = [h[-t][b] for h in hist for t in range(min(100, len(h)))]
        assert abs(m[f"rollout/ep_score_mean_bucket{b}"] - sum(vals) / len(vals)) < 1e-6
    got = mon.ep_score_history()
    assert got[0] == hist[0] and got[1] == hist[1] and got[2] == hist[2][-100:]


def test_gym_single_env_contract():
    from factory_marl_b200.vec_env import FactoryEnv
    env = FactoryEnv("PauseIKToggleEnv", num_arms=2)
    obs, info = env.reset(seed=1)
    assert obs.shape == (194,) and obs.dtype == np.float32 and info == {}
    obs, r, term, trunc, info = env.step(np.array([1, 0]))
    assert obs.shape == (194,) and isinstance(r, float) and term is False and trunc is False and info["scores"] == [0, 0]
    with pytest.raises(ValueError):
        env.vec.step(torch.zeros((1, 3), dtype=torch.int32))
    env.close()


def test_checkpoint_roundtrip_resumes_bit_exactly(tmp_path):
    """SB3-layout zip (policy.pth + policy.optimizer.pth + data) with the env state: a PPO restored from it continues exactly as
    the original (the reference can only restart from scratch, learning.py has no resume path)."""
    from factory_marl_b200.ppo import PPO
    a = PPO(_env(128), n_steps=4, batch_size=256, n_epochs=2, seed=3, use_graph=False)
    a.collect_rollouts(); a.train()
    p = str(tmp_path / "ck.zip")
    a.save(p)
    with zipfile.ZipFile(p) as z:
        assert {"data", "policy.pth", "policy.optimizer.pth", "pytorch_variables.pth", "_stable_baselines3_version", "fm_env_state.npz"} <= set(z.namelist())
        sd = torch.load(io.BytesIO(z.read("policy.pth")), weights_only=True)
        opt = torch.load(io.BytesIO(z.read("policy.optimizer.pth")), weights_only=True)
        assert list(sd)[0] == "log_std" and sd["mlp_extractor.policy_net.0.weight"].shape == (128, 178) and len(opt["state"]) == 13
        assert opt["state"][1]["exp_avg"].shape == (128, 178) and opt["param_groups"][0]["eps"] == 1e-5
        assert json.loads(z.read("data"))["num_timesteps"] == 512
    b = PPO(_env(128), n_steps=4, batch_size=256, n_epochs=2, seed=99, use_graph=False)
    b.load(p)
    assert torch.equal(a.pol.params, b.pol.params) and torch.equal(a.pol.adam_m, b.pol.adam_m) and torch.equal(a.pol.adam_v, b.pol.adam_v)
    assert torch.equal(a.env.core.f, b.env.core.f) and torch.equal(a.env.core.i, b.env.core.i) and torch.equal(a.env.core.rng, b.env.core.rng)
    for ppo in (a, b):
        ppo.collect_rollouts()
    # the resumed rollout is bit-identical: same env state, same PCG64 / action-noise streams, same parameters
    assert torch.equal(a.obs, b.obs) and torch.equal(a.b_obs, b.b_obs) and torch.equal(a.b_act, b.b_act) and torch.equal(a.b_adv, b.b_adv)
    for ppo in (a, b):
        ppo.train()
    # the update accumulates weight gradients with split-K float atomics: equal up to summation order
    assert float((a.pol.params - b.pol.params).abs().max()) < 1e-5


@pytest.mark.parametrize("run", ["r666unuv", "rk5rxnav", "xfwgqibb", "y6lp1j7k"])
def test_reference_policy_pth_roundtrip(run, tmp_path):
    """the reference's trained policies (policy.pth of runs/<run>.zip, committed as tests/golden/ref_policy_<run>.npz) load into
    the on-GPU MLP and come back out of save() tensor for tensor; the forward pass matches a torch fp32 MLP of the same weights"""
    from factory_marl_b200.ppo import PolicyValueNet
    w = dict(np.load(os.path.join(GOLDEN, f"ref_policy_{run}.npz")))
    discrete = "log_std" not in w
    D = w["mlp_extractor.policy_net.0.weight"].shape[1]
    A = w["action_net.weight"].shape[0] // (2 if discrete else 1)
    pol = PolicyValueNet(D, A, discrete, capacity=256)
    pol.set_sb3_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    sd = pol.sb3_state_dict()
    assert list(sd) == list(w) or set(sd) == set(w)
    for k in w:
        assert np.array_equal(sd[k].numpy(), w[k]), k
    obs = torch.randn((256, D), device="cuda") * 0.5
    _, _, _, val = pol.forward(obs, deterministic=True)
    x = obs.cpu().numpy()
    h = np.tanh(x @ w["mlp_extractor.value_net.0.weight"].T + w["mlp_extractor.value_net.0.bias"])
    h = np.tanh(h @ w["mlp_extractor.value_net.2.weight"].T + w["mlp_extractor.value_net.2.bias"])
    v = (h @ w["value_net.weight"].T + w["value_net.bias"])[:, 0]
    assert np.abs(val.cpu().numpy() - v).max() < 0.05 * max(1.0, np.abs(v).max())      # bf16 tensor-core GEMMs, fp32 accumulate
    if os.path.exists(f"/root/reference/runs/{run}.zip"):
        pol.load_sb3_zip(f"/root/reference/runs/{run}.zip")                                # the real file, where it exists
        assert np.array_equal(pol.sb3_state_dict()["action_net.weight"].numpy(), w["action_net.weight"])


def test_state_views_reject_an_env_index_out_of_range():
    from factory_marl_b200.vec_env import FactoryVecEnv
    env = FactoryVecEnv("PauseIKToggleEnv", num_envs=4, num_arms=2)
    env.reset()
    for bad in (4, -1):
        with pytest.raises(RuntimeError, match="out of range"):
            env.core.get_state(bad)
        with pytest.raises(RuntimeError, match="out of range"):
            env.core.set_state(bad, qpos=np.zeros(89))
        with pytest.raises(RuntimeError, match="out of range"):
            env.core.debug_substeps(1, env=bad)
    q, v, c = env.core.get_state(3)
    assert q.shape == (89,)


def test_dump_trajectory_layout(tmp_path):
    from factory_marl_b200.vec_env import dump_trajectory
    env = _env(4, "PauseIKToggleEnv")
    p = dump_trajectory(env, None, 6, str(tmp_path / "traj.npz"))
    d = np.load(p)
    assert d["qpos"].shape == (7, 89) and d["qvel"].shape == (7, 79) and d["ctrl"].shape == (7, 17) and d["reward"].shape == (6,)
    assert np.allclose(np.linalg.norm(d["qpos"][-1, 4:8]), 1.0, atol=1e-5)            # cube 0 was spawned: unit quaternion


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_ppo_step_keeps_parameters_identical():
    """env shards differ per rank; after one train() (gradient all-reduce per optimiser step, NCCL) every rank holds the same
    parameters"""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29641", os.path.join(root, "tools", "two_rank_ppo_check.py")],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "PARAMS_EQUAL True" in out.stdout, out.stdout[-2000:]
