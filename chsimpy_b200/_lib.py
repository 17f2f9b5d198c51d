"""ctypes binding of libchs_b200.so (include/chs_b200.h).  There is no CPU fallback: if
the CUDA library is missing it is built with nvcc, and if that fails the import error
is raised to the caller."""
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libchs_b200.so")
SOURCES = ("chs_api.cu", "chs_ll.cu", "chs_kernels.cuh", "chs_slab.cuh", "chs_big.cuh", "chs_gemm.cuh", "dct_core.cuh", "fastlog.cuh", "chs_rt.h")
UNITS = ("chs_api.cu", "chs_ll.cu")          # translation units of the library
# (no -split-compile: it builds 2x faster but the kernels measured 6 % slower on B200)
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-shared", "-Xcompiler", "-fPIC"]


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("RT", "BRT", "B", "A0", "A1", "Amr", "kappa_tilde", "L", "delx", "delt", "delt_max",
                 "M_tilde", "threshold", "time_limit_s", "jitter")] + [("full_sim", C.c_int32),
                                                                        ("adaptive_time", C.c_int32)]


class State(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("delt", "time_delta_sum", "time_passed", "tau0", "t0")] + \
               [("computed_steps", C.c_int64), ("skip_check", C.c_int32), ("stop_reason", C.c_int32)]


STOP_NAMES = {0: 'None', 1: 'energy', 2: 'time-limit', 3: 'nan'}

# every symbol include/chs_b200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "chs_abi_version": (C.c_int32, []),
    "chs_last_error": (C.c_char_p, []),
    "chs_supports_n": (C.c_int32, [C.c_int32]),
    "chs_uses_gemm": (C.c_int32, [C.c_int32, C.c_int32]),
    "chs_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "chs_create": (C.c_void_p, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "chs_destroy": (None, [C.c_void_p]),
    "chs_set_params": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(Params)]),
    "chs_set_state": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(State)]),
    "chs_get_state": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(State)]),
    "chs_prepare": (C.c_int, [C.c_void_p, C.c_void_p]),
    "chs_begin": (C.c_int, [C.c_void_p]),
    "chs_steps": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32]),
    "chs_poll": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "chs_rewind_rows": (C.c_int, [C.c_void_p]),
    "chs_end": (C.c_int, [C.c_void_p]),
    "chs_dctn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "chs_idctn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "chs_pcg64_fill": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int64]),
    "chs_row_means": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "chs_lcg_fill": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_void_p]),
    "chs_debug_log": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "chs_launch_count": (C.c_int64, [C.c_void_p]),
    "chs_slab_supports_n": (C.c_int32, [C.c_int32]),
    "chs_slab_row_granularity": (C.c_int32, [C.c_int32]),
    "chs_slab_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "chs_slab_create": (C.c_void_p, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Params),
                                     C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "chs_slab_destroy": (None, [C.c_void_p]),
    "chs_slab_row": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double]),
    "chs_slab_transpose": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "chs_slab_transpose_peers": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "chs_slab_sums_peers": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "chs_slab_control_gathered": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "chs_slab_colsum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "chs_slab_control_dyn": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "chs_slab_step_x": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_void_p]),
    "chs_slab_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "chs_slab_pcg64_fill": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int64]),
    "chs_slab_row_means": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "chs_big_supports_n": (C.c_int32, [C.c_int32]),
    "chs_big_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "chs_big_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "chs_big_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32]),
    "chs_big_phys": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "chs_big_sums": (C.c_int, [C.c_void_p]),
    "chs_slab_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "chs_slab_yedge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "chs_slab_clear_yedge": (C.c_int, [C.c_void_p]),
    "chs_slab_reduce": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "chs_slab_vec": (C.c_void_p, [C.c_void_p]),
    "chs_slab_prepare": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double]),
    "chs_slab_control": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "chs_slab_begin": (C.c_int, [C.c_void_p]),
    "chs_slab_rewind_rows": (C.c_int, [C.c_void_p]),
    "chs_slab_get_state": (C.c_int, [C.c_void_p, C.POINTER(State), C.c_void_p, C.c_void_p]),
    "chs_slab_set_state": (C.c_int, [C.c_void_p, C.POINTER(State)]),
    "chs_slab_launch_count": (C.c_int64, [C.c_void_p]),
    "chs_slab_transpose_stage": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "chs_slab_copy_blocks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "chs_slab_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "chs_slab_sums": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "chs_set_timing": (C.c_int, [C.c_void_p, C.c_int32]),
    "chs_get_timing": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "chs_set_mix": (C.c_int, [C.c_void_p, C.c_int32]),
    "chs_get_timing_mix": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}


def _stale(lib_path):
    if not os.path.exists(lib_path):
        return True
    t = os.path.getmtime(lib_path)
    return any(os.path.getmtime(os.path.join(CSRC, s)) > t for s in SOURCES if os.path.exists(os.path.join(CSRC, s)))


def build(force=False, verbose=False):
    """Compiles csrc/chs_api.cu for sm_100a into chsimpy_b200/libchs_b200.so (in-tree).  Safe under
    torchrun: one process builds under an exclusive file lock into a temporary file that is renamed into
    place; the others wait for the lock and find the library fresh."""
    if not force and not _stale(LIB_PATH):
        return LIB_PATH
    import fcntl
    import tempfile
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale(LIB_PATH):          # another rank built it while we waited
                return LIB_PATH
            nvcc = os.environ.get("NVCC", "nvcc")
            fd, tmp = tempfile.mkstemp(prefix=".libchs_b200.", suffix=".so.tmp", dir=HERE)
            os.close(fd)
            cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp] + [os.path.join(CSRC, u) for u in UNITS]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libchs_b200.so:\n" + r.stdout + r.stderr)
            os.chmod(tmp, 0o755)
            os.replace(tmp, LIB_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


def bind(lib):
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def load():
    """Returns the bound CUDA library, building it on first use.  Never falls back."""
    global _lib
    if _lib is None:
        # CHS_B200_LIB: an alternative nvcc build of the SAME CUDA library for tuning experiments
        # (tools/build_variant.sh -> variants/*.so).  Only files inside this repository's chsimpy_b200/ or
        # variants/ directories are accepted, and never the host-emulation test harness: no CPU fallback.
        path = os.environ.get("CHS_B200_LIB")
        if path:
            real = os.path.realpath(path)
            roots = (os.path.realpath(HERE), os.path.realpath(os.path.join(HERE, "..", "variants")))
            if not any(real.startswith(r + os.sep) for r in roots) or "emu" in os.path.basename(real):
                raise RuntimeError(f"CHS_B200_LIB={path}: only CUDA builds under chsimpy_b200/ or variants/ may be loaded")
        else:
            path = build()
        _lib = bind(C.CDLL(path))
        if _lib.chs_abi_version() != 1:
            raise RuntimeError("libchs_b200.so ABI mismatch")
    return _lib


def check(lib, rc, what):
    if rc is None or (isinstance(rc, int) and rc < 0):
        raise RuntimeError(f"{what} failed: {lib.chs_last_error().decode()}")
    return rc
