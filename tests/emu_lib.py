"""TEST HARNESS: compiles the kernel sources (chsimpy_b200/csrc/*.cu*) for the HOST with
-DCHS_EMU (one OS thread per CUDA thread, see csrc/emu.h) so the index logic of the real
kernels can be exercised through the real C ABI in the GPU-less build container.

This is not a fallback: the package itself only ever loads libchs_b200.so (nvcc, sm_100a);
only tests construct an `EmuBackend` and pass it in explicitly."""
import ctypes as C
import os
import subprocess

import numpy as np

from chsimpy_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "_build")
EMU_LIB = os.path.join(OUT, "libchs_emu.so")


def build_emu():
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(_lib.CSRC, s) for s in _lib.SOURCES + ("emu.h",)]
    if os.path.exists(EMU_LIB) and all(os.path.getmtime(s) <= os.path.getmtime(EMU_LIB) for s in srcs):
        return EMU_LIB
    cmd = ["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-DCHS_EMU", "-x", "c++", "-shared", "-fPIC",
           "-pthread", "-o", EMU_LIB] + [os.path.join(_lib.CSRC, u) for u in _lib.UNITS]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emu build failed:\n" + r.stderr)
    return EMU_LIB


class EmuBackend:
    name = "emu"

    def __init__(self):
        self.lib = _lib.bind(C.CDLL(build_emu()))

    def empty(self, shape, dtype="f8"):
        return np.zeros(shape, dtype=np.float64 if dtype == "f8" else np.uint8)

    def ptr(self, t):
        assert t.flags["C_CONTIGUOUS"]
        return t.ctypes.data

    def upload(self, t, arr):
        t[...] = arr

    def to_device(self, arr):
        return np.ascontiguousarray(arr)

    def download(self, t):
        return np.array(t, copy=True)

    def stream_handle(self):
        return None

    def device_index(self):
        return 0
