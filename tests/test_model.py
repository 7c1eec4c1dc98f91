"""Model compiler (factory_marl_b200/model.py) against hand-checkable anchors derived from the reference XML
 and against numpy/scipy sub-oracles that do exist in this image."""
import numpy as np
import pytest

from factory_marl_b200 import model as fm
from factory_marl_b200._capi import pcg64_state_words


def test_sizes_match_survey_a1():
    for n, (nb, ng, nq, nv, nu) in {2: (54, 163, 89, 79, 17), 4: (82, 303, 107, 97, 33), 6: (110, 443, 125, 115, 49)}.items():
        M = fm.build_model(n)
        assert tuple(M["sizes"][[0, 1, 3, 4, 5]]) == (nb, ng, nq, nv, nu)
        assert int((M["geom_collide"] == 1).sum()) == {2: 133, 4: 243, 6: 353}[n]
        assert int((M["geom_arm"] >= 0).sum()) == 70 * n            # base_env.py:121-126: 61 + 9 geoms per arm


def test_six_arm_scene_layout():
    """scene.py:114 table length 1 + 0.5 (n - 2) / 2, :140 bucket y = 0.7 - (n // 2 - 1), :153-158 arms 4 and 5 sit between the rows
    at 0.9 x the lateral offset"""
    M = fm.build_model(6)
    exp = [(0.7, -2.0), (-0.7, -2.0), (0.7, -0.6), (-0.7, -0.6), (0.63, 0.7), (-0.63, 0.7)]
    assert np.allclose(M["arm_base_pos"][:, :2], exp, atol=1e-12) and np.allclose(M["arm_base_pos"][:, 2], 1.0)
    bucket_floor = M["geom_body"][M["bucket_geom"]]
    assert np.allclose(sorted(M["body_pos"][bucket_floor][:, 0]), [-0.9, 0.9]) and np.allclose(M["body_pos"][bucket_floor][:, 1], 0.7 - 2)
    from oracle.oracle_py import Oracle
    o = Oracle(fm.build_blob(6)); o.env_create("PauseIKToggleEnv"); assert len(o.env_reset()) == 6 * 24 + 130 + 48


def test_cube_sizes_scene_seed_42(model2):
    # default_rng(42) interleaved with 4 colour draws (scene.py:121-128)
    exp = [0.045479120971119266, 0.04951244703273512, 0.03741596048465162, 0.034544774435695536, 0.04516175480170748]
    assert np.allclose(model2["cube_size"][:5], exp, rtol=0, atol=1e-15)
    assert np.allclose(model2["body_mass"][model2["cube_body"]], 1000 * model2["cube_size"] ** 3)


def test_fk_known_answers_a5(model2):
    M = model2
    q = fm.qpos0(M)
    xp, xq = fm.forward_kinematics(M, q)
    for a, exp in enumerate([(0.7, 0, 2.356), (-0.7, 0, 2.356)]):
        b = M["site_body"][a]
        site = xp[b] + fm.q2mat(xq[b]) @ M["site_pos"][a]
        assert np.allclose(site, exp, atol=1e-12)
    dp = [-0.5, -0.5, 0, 1.0, 0, -1.6, 0]                           # ik_policy.py:71
    for a in range(2):
        for k in range(7):
            q[M["jnt_qposadr"][M["arm_jnt"][a, k]]] = dp[k]
    xp, xq = fm.forward_kinematics(M, q)
    b = M["site_body"][0]
    site = xp[b] + fm.q2mat(xq[b]) @ M["site_pos"][0]
    assert np.allclose(site, (0.166715, 0.291335, 1.581032), atol=1e-6)


def test_arm_mass_and_mass_matrix(model2):
    M = model2
    b0 = int(M["arm_base_body"][0])
    assert abs(M["body_mass"][b0:b0 + 8].sum() - 30.61) < 1e-12    # iiwa14.xml:56-135
    assert abs(M["body_mass"][b0 + 9:b0 + 12].sum() - 0.76) < 1e-12  # gripper.xml:5,10,27
    H = fm.mass_matrix(M, fm.qpos0(M))
    assert np.allclose(H, H.T)
    assert np.linalg.eigvalsh(H).min() > 0
    assert abs(H[0, 0] - 1000.0) < 1e-9                              # belt
    assert np.allclose(np.diag(H)[1:4], M["body_mass"][M["cube_body"][0]])


def test_cube_group_matches_scipy():
    R = pytest.importorskip("scipy.spatial.transform").Rotation
    q = R.create_group("O").as_quat()
    assert np.allclose(q, fm.CUBE_GROUP_XYZW, atol=1e-15)


def test_pcg64_words_reproduce_numpy_stream():
    # TaskManager.rng = default_rng(seed) (task_utils.py:19); SURVEY.md A.3 first draws
    w = pcg64_state_words(42)
    st = np.random.PCG64(42).state["state"]
    assert (int(w[0]) << 64) | int(w[1]) == st["state"] and (int(w[2]) << 64) | int(w[3]) == st["inc"]
    assert np.allclose(np.random.default_rng(42).uniform(0, 1, 4), [0.7739560485559633, 0.4388784397520523, 0.8585979199113825, 0.6973680290593639])


def test_spawn_steps_float_quirk():
    # task_utils.py:131-133 with spawn_freq = 0.1 * num_arms (base_env.py:36): 49 (not 50) for 2 arms, 24 for 4
    assert int(1.0 / (0.1 * (0.1 * 2))) == 49 and int(1.0 / (0.1 * (0.1 * 4))) == 24
    assert int((1 / 10) / 0.001) == 100                               # base_env.py:134


def test_gpu_tables_consistent(model2):
    G = fm.gpu_tables(model2)
    assert G["g_cluster_pair_all"].shape[0] == 859 and G["g_cluster_pair"].shape[0] < 859 and G["g_arm_sph"].shape == (46, 6) and G["g_arm_box"].shape == (9, 8)
    # bounding spheres enclose their geoms
    for s in G["g_arm_sph"]:
        c = G["g_arm_cluster"][int(s[0])]
        assert np.linalg.norm(s[2:5] - c[:3]) + s[5] <= c[3] + 1e-9
    blob = fm.pack_blob({**model2, **G})
    assert blob[:4] == b"1BMF"
