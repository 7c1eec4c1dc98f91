"""GPU suite (-m gpu): PPO kernels (bf16 tensor-core MLP, heads, GAE, loss/backward, Adam) against plain PyTorch fp32
references of the same ops.  Tolerances: bf16 operands with fp32 accumulation -> 2e-2 relative on activations/gradients;
fp32 kernels (GAE, Adam) -> 1e-5."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _torch_nets(pol):
    sd = {k: v.cuda().float() for k, v in pol.sb3_state_dict().items()}

    def mlp(x, pre):
        h = torch.tanh(x @ sd[f"{pre}.0.weight"].t() + sd[f"{pre}.0.bias"])
        return torch.tanh(h @ sd[f"{pre}.2.weight"].t() + sd[f"{pre}.2.bias"])

    def fwd(x):
        pi = mlp(x, "mlp_extractor.policy_net") @ sd["action_net.weight"].t() + sd["action_net.bias"]
        v = mlp(x, "mlp_extractor.value_net") @ sd["value_net.weight"].t() + sd["value_net.bias"]
        return pi, v[:, 0]
    return sd, fwd


@pytest.mark.parametrize("obs_dim,act_dim,discrete", [(178, 16, False), (258, 4, True), (274, 48, False), (322, 6, True)])
def test_policy_forward_matches_fp32(obs_dim, act_dim, discrete):
    from factory_marl_b200.ppo import PolicyValueNet
    pol = PolicyValueNet(obs_dim, act_dim, discrete, seed=3, capacity=512)
    sd, fwd = _torch_nets(pol)
    x = torch.randn(512, obs_dim, device="cuda")
    act_env, act_raw, logp, val = pol.forward(x, deterministic=True)
    pi, v = fwd(x)
    assert torch.allclose(val, v, atol=2e-2, rtol=2e-2)
    if discrete:
        assert (act_env == (pi[:, 1::2] > pi[:, 0::2]).int()).float().mean() > 0.98
        lp = torch.log_softmax(pi.view(512, act_dim, 2), -1).gather(-1, act_env.long().unsqueeze(-1)).squeeze(-1).sum(-1)
        assert torch.allclose(logp, lp, atol=2e-2)
    else:
        assert torch.allclose(act_raw, pi, atol=2e-2, rtol=2e-2)
        assert torch.allclose(act_env, pi.clamp(-1, 1), atol=2e-2, rtol=2e-2)
        assert torch.allclose(logp, torch.full_like(logp, -act_dim * 0.9189385332046727), atol=1e-4)
    # stochastic head: sample moments
    _, raw, _, _ = pol.forward(x, deterministic=False)
    if not discrete:
        z = (raw - pi)
        assert abs(float(z.mean())) < 0.05 and abs(float(z.std()) - 1.0) < 0.05


def test_gae_matches_sb3_recurrence():
    from factory_marl_b200 import lib
    from factory_marl_b200.ppo import _bind
    L = _bind(lib.load())
    T, N = 37, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    rew = torch.randn(T, N, device="cuda", generator=g); val = torch.randn(T, N, device="cuda", generator=g)
    done = (torch.rand(T, N, device="cuda", generator=g) < 0.05).to(torch.uint8); last = torch.randn(N, device="cuda", generator=g)
    adv = torch.zeros(T, N, device="cuda"); ret = torch.zeros(T, N, device="cuda")
    assert L.fm_gae(rew.data_ptr(), val.data_ptr(), done.data_ptr(), last.data_ptr(), T, N, 0.99, 0.95, adv.data_ptr(), ret.data_ptr(), None) == 0
    torch.cuda.synchronize()
    r, v, d, lv = rew.double().cpu().numpy(), val.double().cpu().numpy(), done.cpu().numpy().astype(float), last.double().cpu().numpy()
    A = np.zeros((T, N)); lastgae = np.zeros(N)
    for t in reversed(range(T)):                       # SB3 RolloutBuffer.compute_returns_and_advantage [3P]
        nnt = 1.0 - d[t]
        nv = lv if t == T - 1 else v[t + 1]
        delta = r[t] + 0.99 * nv * nnt - v[t]
        lastgae = delta + 0.99 * 0.95 * nnt * lastgae
        A[t] = lastgae
    assert np.abs(adv.cpu().numpy() - A).max() < 1e-4 and np.abs(ret.cpu().numpy() - (A + v)).max() < 1e-4


@pytest.mark.parametrize("discrete", [False, True, "six_arms"])
def test_ppo_update_gradients_match_autograd(discrete):
    import ctypes as C
    from factory_marl_b200.ppo import PolicyValueNet
    D, A, n = (258, 4, 1024) if discrete is True else ((274, 48, 1024) if discrete == "six_arms" else (178, 16, 1024))
    discrete = discrete is True
    pol = PolicyValueNet(D, A, discrete, seed=5, capacity=n)
    sd, _ = _torch_nets(pol)
    if not discrete:
        sd["log_std"] = sd["log_std"] - 0.3
        pol.set_sb3_state_dict(sd)
    for k in sd:
        sd[k] = sd[k].clone().requires_grad_(True)
    g = torch.Generator(device="cuda").manual_seed(1)
    obs = torch.randn(n, D, device="cuda", generator=g)
    act = (torch.rand(n, A, device="cuda", generator=g) < 0.5).float() if discrete else torch.randn(n, A, device="cuda", generator=g) * 0.7
    old_logp = torch.randn(n, device="cuda", generator=g) * 0.1 + (-2.5 if discrete else -16.0)
    adv = torch.randn(n, device="cuda", generator=g); ret = torch.randn(n, device="cuda", generator=g)
    idx = torch.randperm(n, device="cuda", generator=g).int()
    stats = torch.zeros(8, device="cuda")
    pol.grads.zero_()
    rc = pol.lib.fm_ppo_update(C.byref(pol.net), C.byref(pol._ws), obs.data_ptr(), act.data_ptr(), old_logp.data_ptr(), adv.data_ptr(),
                               ret.data_ptr(), idx.data_ptr(), n, 0.2, 0.5, 0.0, 1, stats.data_ptr(), None)
    assert rc == 0
    torch.cuda.synchronize()

    def mlp(x, pre):
        h = torch.tanh(x @ sd[f"{pre}.0.weight"].t() + sd[f"{pre}.0.bias"])
        return torch.tanh(h @ sd[f"{pre}.2.weight"].t() + sd[f"{pre}.2.bias"])
    x = obs[idx.long()]; a = act[idx.long()]
    pi = mlp(x, "mlp_extractor.policy_net") @ sd["action_net.weight"].t() + sd["action_net.bias"]
    v = (mlp(x, "mlp_extractor.value_net") @ sd["value_net.weight"].t() + sd["value_net.bias"])[:, 0]
    if discrete:
        lp = torch.log_softmax(pi.view(n, A, 2), -1).gather(-1, a.long().unsqueeze(-1)).squeeze(-1).sum(-1)
    else:
        lp = torch.distributions.Normal(pi, sd["log_std"].exp()).log_prob(a).sum(-1)
    ad = adv[idx.long()]; ad = (ad - ad.mean()) / (ad.std() + 1e-8)
    ratio = torch.exp(lp - old_logp[idx.long()])
    pl = -torch.min(ratio * ad, ratio.clamp(0.8, 1.2) * ad).mean()
    vl = torch.nn.functional.mse_loss(ret[idx.long()], v)
    (pl + 0.5 * vl).backward()                          # SB3 PPO.train loss with ent_coef = 0 [3P]
    assert abs(float(stats[0]) - float(pl)) < 2e-2 * max(1.0, abs(float(pl))) and abs(float(stats[1]) - float(vl)) < 3e-2 * float(vl)
    got = pol.sb3_state_dict()
    gsd = PolicyValueNet.__new__(PolicyValueNet)        # view the gradient buffer through the same layout
    gsd.__dict__.update(pol.__dict__); gsd.params = pol.grads
    gg = gsd.sb3_state_dict()
    for k in sd:
        ref = sd[k].grad.cpu(); mine = gg[k]
        denom = ref.norm() + 1e-8
        assert float((mine - ref).norm() / denom) < 8e-2, (k, float((mine - ref).norm() / denom), float(denom))


def test_adam_step_matches_torch():
    import ctypes as C
    from factory_marl_b200.ppo import PolicyValueNet
    pol = PolicyValueNet(178, 16, False, seed=7, capacity=64)
    p0 = pol.params.clone()
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=3e-4, eps=1e-5)
    g = torch.Generator(device="cuda").manual_seed(2)
    stats = torch.zeros(8, device="cuda")
    for step in range(1, 4):
        grad = torch.randn(pol.n_params, device="cuda", generator=g) * 0.01
        pol.grads.copy_(grad * 2.0)                     # as if summed over 2 ranks
        ref.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([ref], 0.5)
        opt.step()
        assert pol.lib.fm_adam_step(C.byref(pol.net), 3e-4, 0.9, 0.999, 1e-5, 0.5, 0.5, step, stats.data_ptr(), None) == 0
        torch.cuda.synchronize()
        assert abs(float(stats[5]) - float(grad.norm())) < 1e-4 * float(grad.norm())     # global gradient norm before clipping
        assert torch.allclose(pol.params, ref.detach(), atol=2e-6, rtol=1e-5)
        assert float(pol.grads.abs().max()) == 0.0
    assert torch.allclose(pol.params_bf16.float(), pol.params, atol=1e-2, rtol=1e-2)


def test_ppo_learn_smoke_and_sb3_roundtrip(tmp_path):
    from factory_marl_b200.ppo import PPO, PolicyValueNet
    from factory_marl_b200.vec_env import FactoryVecEnv
    env = FactoryVecEnv("PauseIKToggleEnv", num_envs=128, num_arms=2)
    algo = PPO(env, n_steps=4, batch_size=256, n_epochs=2)
    p0 = algo.pol.params.clone()
    algo.learn(2 * 4 * 128)
    assert algo.pol.adam_step == 2 * 2 * 2 and not torch.equal(p0, algo.pol.params) and torch.isfinite(algo.pol.params).all()
    sd = algo.pol.sb3_state_dict()
    assert sd["mlp_extractor.policy_net.0.weight"].shape == (128, 194) and sd["action_net.weight"].shape == (4, 128)
    other = PolicyValueNet(194, 2, True, seed=9)
    other.set_sb3_state_dict({k: v.cuda() for k, v in sd.items()})
    for k, v in other.sb3_state_dict().items():
        assert torch.equal(v, sd[k])


def test_update_forward_on_tcgen05_is_equivalent_to_the_gemm_chain():
    """default (fm_ppo_use_tcgen05(3)): the update's forward runs the fused tcgen05 kernel (row gather, h1 / h2 kept for the backward
    pass); mode 1 keeps the mma.sync chain there; one PPO iteration from identical state must end at the same parameters up to bf16 noise"""
    from factory_marl_b200.ppo import PPO
    from factory_marl_b200.vec_env import FactoryVecEnv
    res = []
    for mode in (3, 1):
        env = FactoryVecEnv("PauseIKToggleEnv", num_envs=128, num_arms=2, seeds=list(range(42, 170)))
        algo = PPO(env, n_steps=4, batch_size=256, n_epochs=2, seed=1, use_graph=False)
        algo.pol.lib.fm_ppo_use_tcgen05(mode)
        try:
            algo.collect_rollouts(); info = algo.train()
        finally:
            algo.pol.lib.fm_ppo_use_tcgen05(3)
        algo.pol.check_status()
        res.append((algo.pol.params.clone(), info))
    d = (res[0][0] - res[1][0]).abs().max()
    assert float(d) < 2e-3 and torch.isfinite(res[1][0]).all(), float(d)        # 4 Adam steps of lr 3e-4: identical signs almost everywhere
    assert abs(res[0][1]["value_loss"] - res[1][1]["value_loss"]) < 2e-2 * max(1.0, abs(res[0][1]["value_loss"]))


def test_ppo_learn_six_arms():
    """num_arms = 6: 274-dim observations, 48-dim Box actions through rollout (tcgen05 forward, N = 48) and update"""
    from factory_marl_b200.ppo import PPO
    from factory_marl_b200.vec_env import FactoryVecEnv
    env = FactoryVecEnv("AllFullRLProgressRewardEnv", num_envs=128, num_arms=6, gripper_to_closest_cube_reward_factor=0.2,
                        closest_cube_to_bucket_reward_factor=0.4, base_reward=0.4)
    algo = PPO(env, n_steps=4, batch_size=256, n_epochs=2)
    p0 = algo.pol.params.clone()
    algo.learn(2 * 4 * 128)
    assert algo.pol.adam_step == 2 * 2 * 2 and not torch.equal(p0, algo.pol.params) and torch.isfinite(algo.pol.params).all()
    sd = algo.pol.sb3_state_dict()
    assert sd["mlp_extractor.policy_net.0.weight"].shape == (128, 274) and sd["action_net.weight"].shape == (48, 128) and sd["log_std"].shape == (48,)


@pytest.mark.parametrize("obs_dim,act_dim,discrete", [(178, 16, False), (258, 4, True), (226, 32, False), (274, 48, False), (322, 6, True)])
def test_tcgen05_mlp_matches_mma_sync_and_fp32(obs_dim, act_dim, discrete):
    """The fused tcgen05 forward (TMEM accumulators) against the mma.sync GEMM chain and a fp32 torch MLP."""
    from factory_marl_b200.ppo import PolicyValueNet
    pol = PolicyValueNet(obs_dim, act_dim, discrete, seed=11, capacity=640)
    sd, fwd = _torch_nets(pol)
    x = torch.randn(640, obs_dim, device="cuda") * 1.5
    pol.lib.fm_ppo_use_tcgen05(1)
    pol.forward(x, deterministic=True)
    out_tc5 = pol._bufs["out"][0][:640].clone(); val_tc5 = pol.value[:640].clone()
    status = int(pol._bufs["scal"][8:9].view(torch.int32)[0])
    pol.lib.fm_ppo_use_tcgen05(0)
    pol.forward(x, deterministic=True)
    out_mma = pol._bufs["out"][0][:640].clone(); val_mma = pol.value[:640].clone()
    pol.lib.fm_ppo_use_tcgen05(3)
    pi, v = fwd(x)
    n_out = pi.shape[1]
    assert status == 0, "mbarrier wait timed out in the tcgen05 kernel"
    assert torch.allclose(out_tc5[:, :n_out], pi, atol=2e-2, rtol=2e-2), float((out_tc5[:, :n_out] - pi).abs().max())
    assert torch.allclose(val_tc5, v, atol=2e-2, rtol=2e-2)
    assert torch.allclose(out_tc5, out_mma, atol=1e-2, rtol=1e-2) and torch.allclose(val_tc5, val_mma, atol=1e-2, rtol=1e-2)


def test_graph_replayed_optimizer_steps_match_launch_by_launch():
    """PPO.train replays one captured CUDA graph per optimiser step from the second call on (Adam's step count on the device);
    same rollouts, same minibatch permutations -> same parameters as the launch-by-launch path (split-K atomics: to 1e-5)."""
    from factory_marl_b200.vec_env import FactoryVecEnv
    from factory_marl_b200.ppo import PPO
    res = []
    for use_graph in (False, True):
        env = FactoryVecEnv("AllFullRLProgressRewardEnv", num_envs=64, num_arms=2, seed=3, gripper_to_closest_cube_reward_factor=0.2,
                            closest_cube_to_bucket_reward_factor=0.4, base_reward=0.4)
        algo = PPO(env, n_steps=4, batch_size=128, n_epochs=2, seed=1, use_graph=use_graph)
        infos = []
        for _ in range(3):
            algo.collect_rollouts(); infos.append(algo.train())
        assert [i["cuda_graph"] for i in infos] == [False, use_graph, use_graph]
        res.append((algo.pol.params.clone(), algo.pol.adam_step, infos[-1]))
        env.close()
    (p0, s0, i0), (p1, s1, i1) = res
    assert s0 == s1 == 3 * 2 * 2
    assert torch.isfinite(p1).all() and float((p0 - p1).abs().max()) < 1e-4, float((p0 - p1).abs().max())
    assert abs(i0["approx_kl"] - i1["approx_kl"]) < 1e-4
