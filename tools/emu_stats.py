"""Host logic-test build: solver / contact statistics of the bench workload (random N(0,1) actions clipped, 2 arms).
Usage: python tools/emu_stats.py [n_envs] [steps] [skip]"""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, ".")
from factory_marl_b200 import model
from tests.emu_backend import make_core, emu_lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 30
blob = model.build_blob(num_arms=2, seed=42)
core = make_core(blob, n, 2, "AllFullRLProgressRewardEnv", seeds=list(range(42, 42 + n)), auto_reset=True, envs_per_block=1,
                 gripper_to_closest_cube_reward_factor=0.2, closest_cube_to_bucket_reward_factor=0.4, base_reward=0.4)
core.reset()
rng = np.random.default_rng(0)
lib = emu_lib()
lib.fm_emu_counters2.argtypes = [C.c_void_p]
def cnt():
    o = (C.c_longlong * 32)(); lib.fm_emu_counters2(o); return np.array(list(o), dtype=np.int64)
base = None; sbase = None
def satc():
    o = (C.c_longlong * 512)(); lib.fm_emu_sat.argtypes = [C.c_void_p]; lib.fm_emu_sat(o); return np.array(list(o))
for t in range(steps):
    if t == skip: base = cnt(); sbase = satc()
    a = np.clip(rng.standard_normal((n, 16)), -1, 1).astype(np.float32)
    core.env_step(a)
c = cnt() - base
ss = c[0]
print("broad-phase queue rebuilds per env-sub-step %.3f" % (c[31] / ss / 725.0))
print("env-sub-steps", ss, "contacts/substep %.2f queue/substep %.2f need_union frac %.3f awake cubes %.2f" % (c[1] / ss, c[2] / ss, c[3] / ss, c[26] / ss))
for k, nm in enumerate(["cube-only", "arm-only", "arm+contacts"]):
    isl, sw, vis, con = c[4 + 4 * k: 8 + 4 * k]
    print("%-13s islands/substep %.3f sweeps/island %.2f contacts/island %.2f sweep-visits/substep %.2f" % (nm, isl / ss, sw / max(isl, 1), con / max(isl, 1), vis / ss))
print("cube-only sweep histogram", (c[16:26] / max(c[16:26].sum(), 1)).round(3))
print("box-box per substep: fast path %.3f, far-sphere reject %.3f, general SAT %.3f (with an arm geom %.3f)" % (c[27] / ss, c[28] / ss, c[29] / ss, c[30] / ss))
o = (C.c_longlong * 512)(); lib.fm_emu_sat.argtypes = [C.c_void_p]; lib.fm_emu_sat(o)
sat = np.array(list(o)) - sbase; names = model.assemble(2)["geom_names"] if hasattr(model, "assemble") else None
print("general SAT by geom1 id (whole run):", {int(k): int(sat[k]) for k in np.nonzero(sat)[0]})
print("single-cube Newton: calls/substep %.3f, settled %.4f, iterations/call %.2f" % (sat[300] / ss, sat[301] / max(sat[300], 1), sat[302] / max(sat[300], 1)))
print("Newton iterations-to-settle histogram (1..5):", (sat[311:316] / max(sat[311:316].sum(), 1)).round(3))
