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


_DEPS = {
    "fm_lib.cu": ["fm_lib.cu", "fm_types.h", "fm_math.cuh", "fm_collide.cuh", "fm_step.cuh", "fm_task.cuh", "fm_driver.cuh", "fm_ik.cuh",
                  "fm_host_model.h", "fm_host_common.h", "../../include/fm_api.h", "../../include/fm_blob.h"],
    "fm_ppo.cu": ["fm_ppo.cu", "../../include/fm_api.h"],
    "fm_mlp_tc5.cu": ["fm_mlp_tc5.cu"],
}


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc cross-compiles for sm_100a without a GPU (used by __graft_entry__.build()).  One object per .cu so that a
    change to the PPO kernels does not recompile the (large, fully inlined) step kernel."""
    src = os.path.join(_HERE, "csrc")
    bdir = os.path.join(src, "build")
    os.makedirs(bdir, exist_ok=True)
    objs, relink = [], force
    for cu, deps in _DEPS.items():
        obj = os.path.join(bdir, cu.replace(".cu", ".o"))
        if force or _newer(obj, [os.path.join(src, d) for d in deps]):
            cmd = ["nvcc"] + [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(src, cu)]
            subprocess.check_call(cmd)
            relink = True
        objs.append(obj)
    if relink or _newer(SO_PATH, objs):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", SO_PATH] + objs)
    return SO_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the CUDA extension is required; there is no CPU path)")
        _lib = _capi.bind(C.CDLL(SO_PATH))
        _lib.fm_launch_count.restype = C.c_longlong
        _lib.fm_launch_count.argtypes = [C.c_void_p]
    return _lib
