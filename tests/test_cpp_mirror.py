"""The C++20 mirrors include/mpc_b200/{LMPC,NLMPC}.hpp: compile with g++ against the C ABI; on a GPU box they reproduce
the reference's golden vector through the same calls as test/LMPC/test_common.cpp:89-237 and the closed loop of
examples/vanderpol_ex.cpp; without a GPU they fail loudly."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROGRAMS = ["quadrotor_kat", "vanderpol_kat", "user_system_kat"]


def _build(name):
    import __graft_entry__ as g
    g.build()
    exe = os.path.join(ROOT, "tests", "cpp", name)
    src = exe + ".cpp"
    hdrs = [os.path.join(ROOT, "include", "mpc_b200", h) for h in ("LMPC.hpp", "NLMPC.hpp")] + [os.path.join(ROOT, "include", "b200mpc.h")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.run(["g++", "-std=c++20", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", exe, src, "-L",
                        os.path.join(ROOT, "libmpc_b200"), "-lb200mpc", "-Wl,-rpath," + os.path.join(ROOT, "libmpc_b200")], check=True)
    return exe


@pytest.mark.parametrize("name", PROGRAMS)
def test_cpp_mirror_compiles_and_fails_loudly_without_gpu(name):
    exe = _build(name)
    import libmpc_b200 as L
    if L.load_library().b200mpc_device_count() > 0:
        pytest.skip("GPU present")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name", PROGRAMS)
def test_cpp_mirror_known_answers_gpu(name):
    exe = _build(name)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
