"""The C++20 mirror include/mpc_b200/LMPC.hpp: compiles with g++ against the C ABI; on a GPU box it reproduces the
reference's golden vector through the same calls as test/LMPC/test_common.cpp:89-237; without a GPU it fails loudly."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "quadrotor_kat")


def _build():
    import __graft_entry__ as g
    g.build()
    src = EXE + ".cpp"
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "include", "mpc_b200", "LMPC.hpp"))):
        subprocess.run(["g++", "-std=c++20", "-O1", "-I", os.path.join(ROOT, "include"), "-o", EXE, src, "-L",
                        os.path.join(ROOT, "libmpc_b200"), "-lb200mpc", "-Wl,-rpath," + os.path.join(ROOT, "libmpc_b200")], check=True)


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu():
    _build()
    import libmpc_b200 as L
    if L.load_library().b200mpc_device_count() > 0:
        pytest.skip("GPU present")
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_golden_vector_gpu():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
