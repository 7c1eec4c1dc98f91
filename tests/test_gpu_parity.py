"""GPU suite (-m gpu): the CUDA library through the C-ABI against the fp64 oracle and the golden vectors, plus
size-independent properties at BASELINE.json's full size (4096 envs)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from tests import parity_cases as pc  # noqa: E402


class GpuBackend:
    @staticmethod
    def mk(blob, n_envs, n_arms, cls, **kw):
        from factory_marl_b200 import _capi, lib
        from factory_marl_b200.vec_env import TorchCudaBackend
        return _capi.FmCore(lib.load(), TorchCudaBackend("cuda:0"), blob, n_envs, n_arms, cls, **kw)

    to_np = staticmethod(lambda x: x.detach().cpu().numpy())

    @staticmethod
    def from_np(x):
        return torch.from_numpy(np.ascontiguousarray(x)).cuda()

    @staticmethod
    def assign(dst, src):
        dst.copy_(torch.from_numpy(np.ascontiguousarray(src)).to(dst.device))


BE = GpuBackend()


def test_single_substep_free_2arms():
    pc.case_single_substep_free(BE, 2)


def test_single_substep_free_4arms():
    pc.case_single_substep_free(BE, 4)


def test_single_substep_contacts_pairset_bit_exact():
    pc.case_single_substep_contacts(BE)


def test_single_substep_contacts_pairset_one_block_per_sm_shape():
    """14 envs on 12 warps: the lane count is not a multiple of the env count (every cluster pair must be visited once)"""
    pc.case_single_substep_contacts(BE, envs_per_block=14)


def test_golden_rollouts_2arms():
    pc.case_golden_rollouts(BE, 2)


def test_golden_rollouts_strict_mode():
    pc.case_golden_rollouts(BE, 2, sleep_hidden=False)


def test_arm_contact_pairsets_bit_exact():
    pc.case_arm_contact_pairsets(BE)


def test_golden_rollouts_4arms():
    pc.case_golden_rollouts(BE, 4)


def test_single_substep_free_6arms():
    pc.case_single_substep_free(BE, 6)


def test_golden_rollouts_6arms():
    """num_arms = 6 (scene.py:114,153-158): every env class, default shape (one env per block below 2048 envs -> block of 6)"""
    pc.case_golden_rollouts(BE, 6)


def test_golden_rollouts_6arms_one_env_per_block():
    pc.case_golden_rollouts(BE, 6, ["AllFullRLProgressRewardEnv", "PauseIKToggleEnv"], envs_per_block=1)


def test_six_arm_vec_env_steps_with_ik_toggling():
    """six arms through the VecEnv mirror: spaces, 2048 envs (the block-of-6 shape with a ragged last block), 30 steps of random
    toggles with auto-reset; capacity flags stay rare and episodes end"""
    from factory_marl_b200.vec_env import FactoryVecEnv
    n = 2050
    env = FactoryVecEnv("PauseIKToggleEnv", num_envs=n, num_arms=6, seeds=list(range(42, 42 + n)))
    assert env.observation_space.shape == (322,) and list(env.action_space.nvec) == [2] * 6
    obs = env.reset()
    assert obs.shape == (n, 322) and bool(torch.isfinite(obs).all())
    gen = torch.Generator(device="cuda").manual_seed(0)
    for t in range(30):
        a = torch.randint(0, 2, (n, 6), device="cuda", generator=gen, dtype=torch.int32)
        obs, rew, term, _ = env.step(a, with_infos=False)
    assert bool(torch.isfinite(obs).all()) and bool(torch.isfinite(rew).all())
    flags = env.core.i[env.core.off("I_FLAGS")]
    assert float(((flags & 15) != 0).float().mean()) < 0.05
    assert int(env.monitor.ring_head.item()) > 0                 # six arms in 2 m of table: force terminations are frequent


def test_contact_free_100_step_rollout():
    pc.case_contact_free_100_steps(BE)


@pytest.mark.parametrize("cls", ["FactoryManipulationEnv", "PauseIKToggleEnv", "AllDeltaProgressRewardEnv"])
def test_ik_policy_from_settled_state(cls):
    if cls == "AllDeltaProgressRewardEnv":
        pytest.skip("delta map is covered by the golden rollouts")
    pc.case_ik_policy_from_settled_state(BE, cls)


def test_full_size_properties_4096_envs():
    """BASELINE.json configs[1] size: determinism, batch invariance, invariants of the state."""
    from factory_marl_b200.vec_env import FactoryVecEnv
    kw = dict(gripper_to_closest_cube_reward_factor=0.2, closest_cube_to_bucket_reward_factor=0.4, base_reward=0.4)
    gen = torch.Generator(device="cuda").manual_seed(1)
    # moderate joint targets: no arm crashes, so no contact-capacity overflow (overflow truncation depends on atomic order)
    acts = [(torch.rand((4096, 16), device="cuda", generator=gen) * 2 - 1) * 0.4 for _ in range(12)]

    def run(n):
        # 4096 envs select the one-block-per-SM shape (14 envs per block); the small batch is pinned to the same shape
        env = FactoryVecEnv("AllFullRLProgressRewardEnv", num_envs=n, num_arms=2, seeds=list(range(42, 42 + n)), envs_per_block=14, **kw)
        env.reset()
        out = []
        for a in acts:
            obs, rew, term, _ = env.step(a[:n])
            out.append((obs.clone(), rew.clone(), term.clone()))
        return env, out

    env, o1 = run(4096)
    assert int(env.core.i[env.core.off("I_FLAGS")].max()) == 0
    _, o2 = run(4096)
    _, o3 = run(8)
    for (a, b, c) in zip(o1, o2, o3):
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])   # run-to-run determinism
        assert torch.equal(a[0][:8], c[0][:8]) and torch.equal(a[1][:8], c[1][:8])               # independent of batch size
        assert torch.isfinite(a[0]).all() and torch.isfinite(a[1]).all()
    f = env.core.f
    cq = f[env.core.off("F_CUBE_Q"):env.core.off("F_CUBE_Q") + 70].reshape(10, 7, -1)
    qn = torch.linalg.norm(cq[:, 3:7], dim=1)
    in_scene_first = env.core.i[env.core.off("I_NIN")]
    assert (in_scene_first >= 0).all() and (in_scene_first <= 10).all()
    # hidden / settled cubes carry unit quaternions (the freshly spawned one is raw until its first sub-step)
    assert ((qn - 1).abs() < 1e-4).float().mean() > 0.85
    assert (env.core.i[env.core.off("I_NIN")] + env.core.i[env.core.off("I_NOUT")] == 10).all()
    assert env.gpu_launches >= 13


@pytest.mark.parametrize("n_arms,shapes", [(2, (1, 4, 14)), (4, (4, 7)), (6, (1, 6))])
def test_block_shapes_agree(n_arms, shapes):
    """every block shape built into the library (envs per block / warps per block) steps the same env to the same result;
    17 envs leave a ragged last block whose padding slots must not leak"""
    from factory_marl_b200.vec_env import FactoryVecEnv
    kw = dict(gripper_to_closest_cube_reward_factor=0.2, closest_cube_to_bucket_reward_factor=0.4, base_reward=0.4)
    n = 17
    gen = torch.Generator(device="cuda").manual_seed(5)
    acts = [torch.randn((n, 8 * n_arms), device="cuda", generator=gen).clamp(-1, 1) for _ in range(10)]
    outs = []
    for E in shapes:
        # pgs_outer = 1: the small shapes have no big-island phase, so every shape solves arm + contact islands by the same sweeps
        env = FactoryVecEnv("AllFullRLProgressRewardEnv", num_envs=n, num_arms=n_arms, seeds=list(range(7, 7 + n)), envs_per_block=E, pgs_outer=1, **kw)
        env.reset()
        for a in acts:
            obs, rew, term, _ = env.step(a)
        assert int(env.core.i[env.core.off("I_FLAGS")].max()) == 0
        outs.append((obs.clone(), rew.clone(), env.core.i[env.core.off("I_NIN")].clone()))
    for o in outs[1:]:
        assert torch.allclose(o[0], outs[0][0], rtol=0, atol=2e-4), float((o[0] - outs[0][0]).abs().max())
        assert torch.allclose(o[1], outs[0][1], rtol=0, atol=2e-4) and torch.equal(o[2], outs[0][2])


def test_env_order_does_not_change_results():
    """fm_set_env_order / FactoryVecEnv(sort_envs=True): grouping the envs whose arm is in contact into the same blocks changes the
    launch time only; 2050 envs (ragged last block), 25 steps of hard random actions with auto-reset, bit-identical outputs"""
    from factory_marl_b200.vec_env import FactoryVecEnv
    n = 2050
    gen = torch.Generator(device="cuda").manual_seed(11)
    acts = [torch.randn((n, 16), device="cuda", generator=gen).mul(1.5).clamp(-1, 1) for _ in range(25)]
    outs = []
    for sort in (False, True):
        env = FactoryVecEnv("AllFullRLProgressRewardEnv", num_envs=n, num_arms=2, seeds=list(range(42, 42 + n)), sort_envs=sort, **pc.PROGRESS_KW)
        env.reset()
        moved = 0
        for a in acts:
            obs, rew, term, _ = env.step(a, with_infos=False)
            if sort:
                moved += int((env.core._order != torch.arange(n, device="cuda", dtype=torch.int32)).any())
        outs.append((obs.clone(), rew.clone(), env.core.scores.clone(), env.core.i.clone()))
        if sort:
            assert moved > 0                                   # some arm did touch something: the order really was a permutation
    # envs that overflowed a contact capacity (arm pile-ups; cumulative flag row) keep an arbitrary subset of their contacts — which
    # one depends on the arrival order of the narrow-phase lanes, from run to run as well — and are left out
    fl = env.core.off("I_FLAGS")
    clean = (outs[0][3][fl] == 0) & (outs[1][3][fl] == 0)
    assert float(clean.float().mean()) > 0.98
    assert torch.equal(outs[0][0][clean], outs[1][0][clean]) and torch.equal(outs[0][1][clean], outs[1][1][clean])
    assert torch.equal(outs[0][2][clean], outs[1][2][clean]) and torch.equal(outs[0][3][:, clean], outs[1][3][:, clean])


def test_auto_reset_and_episode_info():
    from factory_marl_b200.vec_env import FactoryVecEnv
    env = FactoryVecEnv("AllFullRLProgressRewardEnv", num_envs=64, num_arms=2, force_contact_threshold=1e-3)
    env.reset()
    a = torch.ones((64, 16), device="cuda") * 3.0      # drives the arms into the table: force termination
    done_seen = False
    for t in range(30):
        obs, rew, term, infos = env.step(a, with_infos=True)
        if term.any():
            e = int(torch.nonzero(term)[0])
            assert "terminal_observation" in infos[e] and infos[e]["episode"]["l"] == t + 1
            assert torch.all(obs[e][:48] == 0)          # returned obs is the post-reset observation (qpos0, zero ctrl)
            done_seen = True
            break
    assert done_seen


def test_mean_collected_objects_within_oracle_ci():
    """BASELINE.json: mean collected objects under a fixed policy must fall within the reference's 95% confidence interval.
    Fixed policy = IK base policy, 2 arms; interval from 128 oracle episodes (tests/golden/ik_policy_stats.json, made by
    tools/make_policy_stats.py — the reference itself cannot run here); GPU: 1024 envs, same horizon and bookkeeping."""
    import json, os
    from factory_marl_b200.vec_env import FactoryVecEnv
    from tests.parity_util import GOLDEN
    st = json.load(open(os.path.join(GOLDEN, "ik_policy_stats.json")))
    N, H = 1024, st["horizon"]
    env = FactoryVecEnv("FactoryManipulationEnv", num_envs=N, num_arms=2, seeds=list(range(42, 42 + N)), auto_reset=True)
    env.reset()
    done = torch.zeros(N, dtype=torch.bool, device="cuda")
    total = torch.zeros(N, device="cuda"); length = torch.full((N,), float(H), device="cuda")
    dummy = torch.zeros((N, 1), device="cuda")
    for t in range(H):
        obs, rew, term, _ = env.step(dummy)
        first = term & ~done
        ep = env.core.ep_info
        total = torch.where(first, ep[:, 2] + ep[:, 3], total)
        length = torch.where(first, torch.full_like(length, t + 1.0), length)
        done |= term
    sc = env.core.scores.float().sum(dim=1)
    total = torch.where(done, total, sc)
    mean_total, mean_len = float(total.mean()), float(length.mean())
    lo, hi = st["ci95_total"]
    print("GPU mean total %.3f (oracle %.3f, CI [%.3f, %.3f]); mean length %.1f (oracle %.1f, CI %s)" %
          (mean_total, st["mean_total"], lo, hi, mean_len, st["mean_length"], st["ci95_length"]))
    assert lo <= mean_total <= hi, (mean_total, lo, hi)
    assert st["ci95_length"][0] <= mean_len <= st["ci95_length"][1], (mean_len, st["ci95_length"])


def test_four_arm_capacities_hold_under_ik_toggling():
    """four arms crash into the table and into each other far more often than two: with the four-arm capacities (FmCap<4>: 96
    contacts, 64 arm-side Jacobians, 160 queued pairs) no capacity flag may come up in 50 steps of random IK toggling"""
    from factory_marl_b200.vec_env import FactoryVecEnv
    n = 512
    env = FactoryVecEnv("PauseIKToggleEnv", num_envs=n, num_arms=4, seeds=list(range(42, 42 + n)))
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(0)
    for t in range(50):
        obs, rew, term, _ = env.step(torch.randint(0, 2, (n, 4), device="cuda", generator=gen, dtype=torch.int32))
    assert int(env.core.i[env.core.off("I_FLAGS")].max()) == 0
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    assert int(env.core.i[env.core.off("I_MAXCON")].max()) > 0


# ---- integer outputs at non-trivial values, production solver settings, contact forces, episode boundary (round 2)
def test_scores_both_buckets_snapshot():
    pc.case_scores_both_buckets_snapshot(BE)


def test_out_of_bounds_terminates():
    pc.case_out_of_bounds_terminates(BE)


def test_force_termination_same_step():
    pc.case_force_termination_same_step(BE)


def test_contact_forces_production_and_converged():
    pc.case_contact_forces(BE, **pc.CONVERGED)
    pc.case_contact_forces(BE)


def test_control_step_with_contacts_production():
    pc.case_control_step_with_contacts_production(BE)


def test_impact_substep_production():
    pc.case_impact_substep_production(BE)


def test_episode_boundary_leaks():
    pc.case_episode_boundary(BE)


@pytest.mark.slow
def test_config1_10000_steps_single_env():
    """BASELINE.json configs[0]: 1 env, seeded random actions, 10 000 steps with auto-reset on the CPU oracle; the GPU (N = 1) is
    re-synchronised to the oracle's state every 25 steps and must reproduce the next control step."""
    pc.case_config1(BE, n_steps=10000, sync_every=25)
