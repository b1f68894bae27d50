"""include/mpc_b200/B200Optimizer.hpp is the binding a libmpc++ maintainer adds behind mpc::IOptimizer<sizer>
(INTEGRATION.md section 2).  It must compile against the reference's OWN headers.  Eigen is not in the image, so the check runs
with the minimal stand-in under tests/cpp/eigen_stub (a type-check with -fsyntax-only plus a full compile + link against the
C-ABI library); it needs /root/reference and is skipped where that tree is absent (the GPU box)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/include"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference headers not present")
def test_b200optimizer_compiles_against_reference_headers(tmp_path):
    import __graft_entry__ as g
    g.build()
    src = os.path.join(ROOT, "tests", "cpp", "b200optimizer_compile.cpp")
    exe = str(tmp_path / "b200optimizer_compile")
    cmd = ["g++", "-std=c++20", "-I", os.path.join(ROOT, "tests", "cpp", "eigen_stub"), "-I", REF, "-I",
           os.path.join(ROOT, "include"), "-o", exe, src, "-L", os.path.join(ROOT, "libmpc_b200"), "-lb200mpc",
           "-Wl,-rpath," + os.path.join(ROOT, "libmpc_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    import libmpc_b200 as L
    if L.load_library().b200mpc_device_count() == 0:
        run = subprocess.run([exe], capture_output=True, text=True)
        assert run.returncode != 0          # onInit throws: no CUDA device, no CPU fallback
        assert "no CUDA device" in (run.stderr + run.stdout)
