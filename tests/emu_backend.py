"""Test-only: loads the host-compiled logic-test build of the kernel phases (tests/kernel_emu) behind the same
C-ABI bindings the product uses, with numpy arrays in place of device memory."""
import ctypes as C
import os
import subprocess

import numpy as np

from factory_marl_b200 import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))


class NumpyBackend:
    def zeros(self, shape, dtype):
        return np.zeros(shape, dtype)

    def from_numpy(self, a):
        return np.ascontiguousarray(a).copy()

    def to_numpy(self, a):
        return a

    def ptr(self, a):
        return None if a is None else a.ctypes.data

    def stream(self):
        return None


_lib = None


def emu_lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "kernel_emu"), "-s"])
        _lib = _capi.bind(C.CDLL(os.path.join(_HERE, "kernel_emu", "libfm_emu.so")))
    return _lib


def make_core(blob, num_envs, num_arms, env_class, **kw):
    return _capi.FmCore(emu_lib(), NumpyBackend(), blob, num_envs, num_arms, env_class, **kw)
