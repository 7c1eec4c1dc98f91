"""Build + load libfm_b200.so (hand-written sm_100a CUDA behind the C-ABI of include/fm_api.h).
There is no CPU fallback: loading fails loudly when the extension is missing, and fm_create fails without a GPU."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from . import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libfm_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]
_lib = None


_STEP_HDRS = ["fm_types.h", "fm_math.cuh", "fm_collide.cuh", "fm_step.cuh", "fm_task.cuh", "fm_driver.cuh", "fm_kernel.cuh"]
_SHAPES = ["2_1", "2_4", "2_14", "2_14n", "4_1", "4_4", "4_7", "4_7n", "6_1", "6_6", "6_6n"]        # arms _ envs per block (fm_kernel.cuh); n = with the big-island Newton
_DEPS = {
    "fm_lib.cu": ["fm_lib.cu", "fm_ik.cuh", "fm_host_model.h", "fm_host_common.h", "../../include/fm_api.h", "../../include/fm_blob.h"] + _STEP_HDRS,
    "fm_ppo.cu": ["fm_ppo.cu", "../../include/fm_api.h"],
    "fm_mlp_tc5.cu": ["fm_mlp_tc5.cu"],
}
for _s in _SHAPES:
    _DEPS[f"fm_step_shape_{_s}.cu"] = [f"fm_step_shape_{_s}.cu"] + _STEP_HDRS


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str = SO_PATH, objdir: str | None = None, only=None) -> str:
    """nvcc cross-compiles for sm_100a without a GPU (used by __graft_entry__.build()).  One object per .cu, compiled in
    parallel: the fused step kernel is fully inlined (~50 k SASS lines per block shape), so every shape is its own
    translation unit.  extra_flags / out / objdir: kernel-variant experiments (tools/ab_variant.py)."""
    from concurrent.futures import ThreadPoolExecutor
    src = os.path.join(_HERE, "csrc")
    bdir = objdir or os.path.join(src, "build")
    os.makedirs(bdir, exist_ok=True)
    objs, jobs = [], []
    for cu, deps in _DEPS.items():
        if only is not None and cu not in only:             # kernel-variant experiments: every other object comes from the main build
            objs.append(os.path.join(src, "build", cu.replace(".cu", ".o")))
            continue
        obj = os.path.join(bdir, cu.replace(".cu", ".o"))
        if force or _newer(obj, [os.path.join(src, d) for d in deps]):
            jobs.append(["nvcc"] + [f for f in NVCC_FLAGS if f != "-shared"] + list(extra_flags) + (["-Xptxas", "-v"] if verbose else [])
                        + ["-c", "-o", obj, os.path.join(src, cu)])
        objs.append(obj)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            list(ex.map(subprocess.check_call, jobs))
    if jobs or _newer(out, objs):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs)
    return out


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the CUDA extension is required; there is no CPU path)")
        # FM_B200_LIB: kernel-variant A/B experiments only (tools/ab_variant.sh); the product loads the in-tree library
        _lib = _capi.bind(C.CDLL(os.environ.get("FM_B200_LIB") or SO_PATH))
        _lib.fm_launch_count.restype = C.c_longlong
        _lib.fm_launch_count.argtypes = [C.c_void_p]
    return _lib
