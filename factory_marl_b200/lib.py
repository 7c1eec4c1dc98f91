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
_SOURCES = ["fm_lib.cu", "fm_ppo.cu"]
_lib = None


def _stale() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    src = os.path.join(_HERE, "csrc")
    deps = [os.path.join(src, f) for f in os.listdir(src)] + [os.path.join(_HERE, "..", "include", f) for f in ("fm_api.h", "fm_blob.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc cross-compiles for sm_100a without a GPU (used by __graft_entry__.build())."""
    if force or _stale():
        srcs = [os.path.join(_HERE, "csrc", s) for s in _SOURCES if os.path.exists(os.path.join(_HERE, "csrc", s))]
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO_PATH] + srcs
        subprocess.check_call(cmd)
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
