"""CPU suite (-m "not gpu"): the device phase functions, host-compiled (tests/kernel_emu), against the fp64 oracle and
the golden vectors.  Same test bodies as the GPU suite (tests/parity_cases.py)."""
import numpy as np
import pytest

from tests import parity_cases as pc
from tests.emu_backend import make_core


class EmuBackend:
    mk = staticmethod(make_core)
    to_np = staticmethod(lambda x: np.asarray(x))
    from_np = staticmethod(lambda x: np.ascontiguousarray(x))

    @staticmethod
    def assign(dst, src):
        dst[...] = src


BE = EmuBackend()


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
    pc.case_golden_rollouts(BE, 2, ["AllFullRLProgressRewardEnv", "BackupIKToggleEnv"], sleep_hidden=False)


def test_arm_contact_pairsets_bit_exact():
    pc.case_arm_contact_pairsets(BE)


def test_golden_rollouts_4arms():
    pc.case_golden_rollouts(BE, 4, ["AllFullRLProgressRewardEnv", "PauseIKToggleEnv"])


def test_single_substep_free_6arms():
    """scene.py:114,153-158: table length 2.0, buckets at y = -1.3, arms 4 and 5 at (+-0.63, 0.7)"""
    pc.case_single_substep_free(BE, 6)


def test_golden_rollouts_6arms():
    pc.case_golden_rollouts(BE, 6, ["AllFullRLProgressRewardEnv", "PauseIKToggleEnv", "FactoryManipulationEnv"])


def test_golden_rollouts_6arms_block_of_6():
    pc.case_golden_rollouts(BE, 6, ["SingleDeltaProgressRewardEnv", "BackupIKToggleEnv"], envs_per_block=6)


def test_contact_free_100_step_rollout():
    pc.case_contact_free_100_steps(BE)


@pytest.mark.parametrize("cls", ["FactoryManipulationEnv", "PauseIKToggleEnv"])
def test_ik_policy_from_settled_state(cls):
    pc.case_ik_policy_from_settled_state(BE, cls)


def test_no_intra_interval_dependencies():
    """Race check for the barrier schedule: executing the thread ids of every barrier interval in ascending, descending
    and shuffled order must give bit-identical results (T = 128 is coprime with the shuffle stride)."""
    from factory_marl_b200 import model as fm
    from tests.emu_backend import emu_lib
    from tests.parity_util import load_golden, PROGRESS_KW
    blob = fm.build_blob(2)
    outs = []
    for order in (0, 1, 2):
        emu_lib().fm_emu_set_order(order)
        try:
            res = []
            for cls in ("AllFullRLProgressRewardEnv", "PauseIKToggleEnv"):
                g = load_golden(cls, 2)
                core = make_core(blob, 4, 2, cls, auto_reset=False, **(PROGRESS_KW if "Progress" in cls else {}))
                core.reset()
                for t in range(len(g["actions"])):
                    a = np.tile(g["actions"][t][None], (4, 1))
                    core.env_step(np.ascontiguousarray(a))
                    res.append(core.obs.copy()); res.append(core.reward.copy())
            outs.append(res)
        finally:
            emu_lib().fm_emu_set_order(0)
    for r in outs[1:]:
        for x, y in zip(outs[0], r):
            assert np.array_equal(x, y)


def test_box_on_face_fast_path_equals_general_box_box():
    """200k random cube poses over belt / table / bucket-floor sized boxes: the fast path of the narrow phase returns the
    same contacts (count, order, position, depth) as the general separating-axis routine whenever it applies."""
    import ctypes as C
    from tests.emu_backend import emu_lib
    applied = C.c_int()
    bad = emu_lib().fm_emu_check_box_on_face(200000, C.c_ulonglong(2024), C.byref(applied))
    assert applied.value > 100000 and bad == 0, (bad, applied.value)


def test_golden_rollouts_block_of_7_with_padding_slots():
    """E = 7 envs per block with 4 envs: three padding slots step a copy of the last env and store nothing."""
    pc.case_golden_rollouts(BE, 2, ["AllFullRLProgressRewardEnv"], envs_per_block=7)


def test_block_shapes_agree():
    """the per-env result does not depend on how envs are grouped into blocks (E = 1, 4, 7, and 14 envs on 12 warps: a lane
    count that is not a multiple of E; 5 envs -> ragged last block)"""
    from factory_marl_b200 import model as fm
    blob = fm.build_blob(2)
    outs = []
    for E in (1, 4, 7, 14):
        core = make_core(blob, 5, 2, "AllFullRLProgressRewardEnv", seeds=list(range(42, 47)), auto_reset=True, envs_per_block=E, **pc.PROGRESS_KW)
        core.reset()
        rng = np.random.default_rng(3)
        for t in range(12):
            core.env_step(np.clip(rng.standard_normal((5, 16)), -1, 1).astype(np.float32))
        outs.append((np.array(core.obs), np.array(core.reward), np.array(core.f)))
    for o in outs[1:]:
        assert np.allclose(o[0], outs[0][0], rtol=0, atol=1e-6) and np.allclose(o[1], outs[0][1], atol=1e-6)
        assert np.allclose(o[2], outs[0][2], rtol=0, atol=1e-5)


def test_broad_phase_visits_every_pair_once():
    """lane -> (env, pair) mapping of the broad phase: every env slot of a block visits every admissible cluster pair exactly
    once per queue rebuild (the first sub-step of a launch), also when the lane count (12 warps) is not a multiple of the envs per block (14)"""
    import ctypes as C
    from factory_marl_b200 import model as fm
    from tests.emu_backend import emu_lib
    blob = fm.build_blob(2)
    lib = emu_lib(); lib.fm_emu_counters2.argtypes = [C.c_void_p]

    def visits():
        o = (C.c_longlong * 32)(); lib.fm_emu_counters2(o); return o[31]
    per_slot = {}
    for E, slots in [(1, 4), (4, 4), (7, 7), (14, 14)]:
        core = make_core(blob, 4, 2, "AllFullRLProgressRewardEnv", auto_reset=False, envs_per_block=E)
        core.reset()
        v0 = visits()
        core.debug_substeps(1, env=0)
        per_slot[E] = (visits() - v0) / slots
    n_pairs = per_slot[1]
    assert 600 < n_pairs < 900 and n_pairs == int(n_pairs)
    assert all(v == n_pairs for v in per_slot.values()), per_slot


def test_scores_both_buckets_snapshot():
    pc.case_scores_both_buckets_snapshot(BE)


def test_out_of_bounds_terminates():
    pc.case_out_of_bounds_terminates(BE)


def test_force_termination_same_step():
    pc.case_force_termination_same_step(BE, seeds=(0, 2, 8))


def test_contact_forces_production_and_converged():
    pc.case_contact_forces(BE, **pc.CONVERGED)
    pc.case_contact_forces(BE)


def test_control_step_with_contacts_production():
    pc.case_control_step_with_contacts_production(BE)


def test_impact_substep_production():
    pc.case_impact_substep_production(BE)


def test_episode_boundary_leaks():
    pc.case_episode_boundary(BE)


def test_env_order_does_not_change_results():
    """fm_set_env_order: block slot k serves env order[k].  Same seeds, same actions, a shuffled order and 2 envs per block:
    every env must end exactly where it ends under the identity order; the arm-contact cue (I_ARMCON) is filled"""
    from factory_marl_b200 import model as fm
    blob = fm.build_blob(2)
    n = 6
    rng = np.random.default_rng(4)
    acts = [np.clip(rng.normal(0, 1.5, (n, 16)), -1, 1).astype(np.float32) for _ in range(12)]
    outs = []
    for order in (None, np.array([3, 0, 5, 1, 4, 2], np.int32)):
        core = BE.mk(blob, n, 2, "AllFullRLProgressRewardEnv", seeds=list(range(42, 42 + n)), auto_reset=True, envs_per_block=2, **pc.PROGRESS_KW)
        core.set_env_order(order)
        core.reset()
        for a in acts:
            core.env_step(a)
        outs.append((np.array(core.obs).copy(), np.array(core.reward).copy(), np.array(core.i).copy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
    assert outs[0][2][core.off("I_ARMCON")].max() >= 0
