"""C-ABI library: built in-tree, exports every symbol include/fm_api.h declares, and fails loudly without a GPU
(no CPU path).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "fm_api.h")).read()
    return sorted(set(re.findall(r"^\s*(?:const\s+char\*|int)\s+(fm_[a-z_]+)\s*\(", src, re.M)))


@pytest.fixture(scope="module")
def so():
    from factory_marl_b200 import lib
    return lib.build()


def test_exports_every_declared_symbol(so):
    names = _declared()
    assert {"fm_create", "fm_destroy", "fm_reset", "fm_step", "fm_ik_policy", "fm_get_state", "fm_set_state", "fm_last_error",
            "fm_gae", "fm_policy_forward", "fm_ppo_update"} <= set(names)
    dll = C.CDLL(so)
    for n in names:
        assert hasattr(dll, n), n


def test_no_cpu_fallback(so, blob2):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from factory_marl_b200 import _capi, lib
    dll = lib.load()
    cfg = _capi.fm_config(n_envs=4, n_arms=2, env_class=3, control_frequency=10, spawn_freq=0.1, spawn_freq_increase=1.001, pt_time=0.2)
    h = C.c_void_p()
    rc = dll.fm_create(C.byref(cfg), blob2, len(blob2), C.byref(h))
    assert rc != 0 and b"no CUDA device" in dll.fm_last_error()
    from factory_marl_b200.vec_env import FactoryVecEnv
    with pytest.raises(RuntimeError):
        FactoryVecEnv("AllFullRLProgressRewardEnv", num_envs=4)


def test_oracle_is_not_imported_by_the_product():
    pkg = os.path.join(ROOT, "factory_marl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle_py" not in txt and "libfm_oracle" not in txt and "kernel_emu" not in txt.replace("tests/kernel_emu", ""), f
