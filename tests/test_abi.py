"""The C-ABI library builds for sm_100a without a GPU and exports every symbol that
include/chs_b200.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

from chsimpy_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "chs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(chs_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.PROTOTYPES)


def test_library_builds_and_exports_everything():
    path = _lib.build()
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    _lib.bind(lib)
    assert lib.chs_abi_version() == 1
    assert [n for n in (4, 16, 32, 64, 100, 200, 512, 1024, 2048) if lib.chs_supports_n(n)] == [16, 32, 64, 100, 512, 1024]
    assert [n for n in (16, 32, 100, 512) if lib.chs_uses_gemm(n, 1)] == [16, 32, 100]
    assert lib.chs_uses_gemm(32, 1024) == 0
    assert lib.chs_workspace_bytes(512, 4) > 0 and lib.chs_workspace_bytes(200, 4) < 0
    assert [n for n in (512, 2048, 16384, 100) if lib.chs_slab_supports_n(n)] == [512, 2048, 16384]


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.Params) == 15 * 8 + 2 * 4
    assert ctypes.sizeof(_lib.State) == 5 * 8 + 8 + 2 * 4


def test_kernels_are_sm100a_sass():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def build_c_demo(out_dir):
    """Compiles examples/c_abi_demo.cpp (a plain C++ client of include/chs_b200.h: dlopen + the CUDA
    runtime, no Python/PyTorch) and returns the path of the binary."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = os.path.join(str(out_dir), "c_abi_demo")
    cmd = ["g++", "-std=c++17", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "c_abi_demo.cpp"),
           "-o", exe, "-I" + os.path.join(cuda, "include"), "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-ldl",
           "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_plain_cpp_client_compiles_against_the_header(tmp_path):
    assert os.path.exists(build_c_demo(tmp_path))
