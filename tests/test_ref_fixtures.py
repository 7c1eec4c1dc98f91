"""CPU suite: the fp64 oracle AND the host-compiled kernel build against reference-held data (tests/ref_cases.py): real MuJoCo
observations and Monitor episodes decoded from /root/reference/runs/*.zip by tools/make_ref_fixtures.py."""
import json
import os

import numpy as np
import pytest

from tests import ref_cases as rc
from tests.parity_util import GOLDEN
from tests.sim_adapters import CoreSim, OracleSim
from tests.test_emu_parity import BE as EMU


def mk_oracle(n_arms, cls, **kw):
    return OracleSim(n_arms, cls, **kw)


def mk_emu(n_arms, cls, **kw):
    return CoreSim(EMU, n_arms, cls, n_envs=2, **kw)


MK = {"oracle": mk_oracle, "kernel_emu": mk_emu}


def test_fixtures_match_reference_zips_when_present():
    """the committed fixtures are what tools/make_ref_fixtures.py decodes (checked where /root/reference exists)"""
    if not os.path.exists("/root/reference/runs/y6lp1j7k.zip"):
        pytest.skip("reference not present on this box")
    import base64, pickle, warnings, zipfile
    for run in rc.N_ARMS:
        data = json.loads(zipfile.ZipFile(f"/root/reference/runs/{run}.zip").read("data"))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            eps = pickle.loads(base64.b64decode(data["ep_info_buffer"][":serialized:"]))
            obs = pickle.loads(base64.b64decode(data["_last_obs"][":serialized:"]))
        assert [int(e["l"]) for e in eps] == rc.RUNS[run]["lengths"]
        assert np.allclose([e["r"] for e in eps], rc.RUNS[run]["returns"])
        assert np.array_equal(np.asarray(obs, np.float32), rc.LAST_OBS[run])


@pytest.mark.parametrize("who", ["oracle", "kernel_emu"])
def test_ref_free_fall_chains(who):
    rc.case_ref_free_fall(MK[who])


@pytest.mark.parametrize("who", ["oracle", "kernel_emu"])
def test_ref_rest_depth_and_belt_lag(who):
    rc.case_ref_rest_depth_and_belt_lag(MK[who])


@pytest.mark.parametrize("who", ["oracle", "kernel_emu"])
def test_ref_arm_droop(who):
    rc.case_ref_arm_droop(MK[who])


@pytest.mark.parametrize("run", ["rk5rxnav", "r666unuv", "xfwgqibb", "y6lp1j7k"])
@pytest.mark.parametrize("who", ["oracle", "kernel_emu"])
def test_ref_scene_step(who, run):
    rc.case_ref_scene_step(MK[who], run, min_envs=1)


def test_ref_ik_proposals():
    rc.case_ref_ik_proposals()
