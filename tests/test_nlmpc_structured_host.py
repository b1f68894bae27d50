"""CPU: the stage-structured NLMPC solver (libmpc_b200/csrc/nlmpc_structured.cuh) run through HOST EMULATION -- the exact device
source compiled by g++ with a thread group of one (tests/cpp/nl_structured_host.cpp) -- against
  * the Python specification of the same algorithm (tests/nlmpc_sqp_reference.py with the per-stage block BFGS), and
  * the SLSQP oracle (oracle/nlmpc_slsqp.py; stand-in for NLopt's LD_SLSQP at NLOptimizer.hpp:519, PARITY UNPINNED upstream)
on the reference's example systems and on BASELINE configs[2] (unicycle nx3 nu2 Tph30).  The GPU run of the same source is
tests/test_gpu_nlmpc_structured.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from nlmpc_sqp_reference import QPADMM, sqp_solve, stage_groups
from oracle import nlmpc_slsqp as S
from oracle.nlmpc_formulation import oscnet_formulation, ugv_formulation, vanderpol_formulation

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def host_lib():
    global _lib
    if _lib is not None:
        return _lib
    from libmpc_b200 import workloads as W
    cdir = os.path.join(ROOT, "tests", "cpp")
    hdr = os.path.join(cdir, "_unicycle_user_system.inc")
    src = os.path.join(cdir, "nl_structured_host.cpp")
    so = os.path.join(cdir, "libnls_host.so")
    if not os.path.exists(hdr) or open(hdr).read() != W.UNICYCLE_SRC:
        open(hdr, "w").write(W.UNICYCLE_SRC)
    deps = [src, hdr] + [os.path.join(ROOT, "libmpc_b200", "csrc", f) for f in ("nlmpc_structured.cuh", "nlmpc_kernels.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "libmpc_b200", "csrc"),
                        f'-DUSER_SYSTEM_HEADER="{hdr}"', f"-DUSER_SYSTEM_TYPE={W.UNICYCLE_TYPE}", "-o", so, src], check=True)
    _lib = C.CDLL(so)
    _lib.nls_host_solve.argtypes = [C.c_int] * 3 + [C.c_void_p] * 5 + [C.c_int, C.c_int] + [C.c_void_p] * 2
    return _lib


def host_solve(system, ph, ch, z0, x0, params, lb, ub, max_sqp=100, max_qp=200):
    lib = host_lib()
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (z0, x0, params, lb, ub)]
    z = np.zeros_like(arrs[0]); out = np.zeros(5)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.nls_host_solve(system, ph, ch, *[vp(a) for a in arrs], max_sqp, max_qp, vp(z), vp(out))
    assert rc == 0, rc
    return dict(z=z, cost=out[0], viol=out[1], status=int(out[2]), nit=int(out[3]), qp_iters=int(out[4]))


CASES = [
    ("vanderpol_ex (ch < ph)", 0, vanderpol_formulation(), np.array([0.1]), True, np.array([0.0, 1.0])),
    ("ugv_ex, obstacle active", 3, ugv_formulation(10, 10, v_pref=(0.6, 0.8)), None, False, np.array([0.4, 0.5, 0.6, 0.8])),
    ("networked oscillators N=4 ph15 ch8 (BASELINE configs[3])", 1, oscnet_formulation(4, 15, 8), np.array([0.1, 1.0, 0.1]), True,
     np.array([0.3, -0.2, 0.5, 0.1, -0.4, 0.2, 0.1, -0.3])),
]


@pytest.mark.parametrize("name,sid,f,params,hard,x0", CASES)
def test_host_emulated_kernel_matches_specification_and_slsqp(name, sid, f, params, hard, x0):
    if params is None:
        params = f.params
    lb, ub = S.default_bounds(f, hard)
    if not hard:
        lb[-1] = 0.0
    z0 = S.initial_guess(f, x0, np.zeros(f.nu), lb=lb, ub=ub)
    out = host_solve(sid, f.ph, f.ch, z0, x0, params, lb, ub)
    spec = sqp_solve(f, x0, z0, lb, ub, bfgs_groups=stage_groups(f), qp=QPADMM(carry_rho=True))
    ref = S.solve(f, x0, z0, lb, ub)
    assert out["status"] == 0 and out["viol"] < 1e-8
    # same algorithm in two languages: the iterates differ by finite-difference noise only
    assert abs(out["nit"] - spec["nit"]) <= 3, (out["nit"], spec["nit"])
    assert np.abs(out["z"] - spec["z"]).max() < 2e-5
    assert abs(out["cost"] - spec["cost"]) < 1e-8 * max(1.0, abs(spec["cost"]))
    # and the optimum NLopt's SLSQP stand-in finds
    assert ref["success"]
    assert abs(out["cost"] - ref["cost"]) < 1e-7 * max(1.0, abs(ref["cost"]))
    cmd = out["z"][f.ph * f.nx:f.ph * f.nx + f.nu]
    assert np.abs(cmd - ref["cmd"]).max() < 1e-5 * max(1.0, np.abs(ref["cmd"]).max())


def test_host_emulated_kernel_unicycle_cfg2_matches_slsqp_fixture():
    """BASELINE configs[2] at its stated shape, cold start: first command 1e-5 / cost 1e-7 against the committed SLSQP fixture."""
    from libmpc_b200 import workloads as W
    g = np.load(os.path.join(ROOT, "tests", "golden", "nlmpc_unicycle.npz"))
    x0, params = W.unicycle_inputs(0, 3)
    lb, ub = W.soft_bounds(151)
    for b in range(3):
        z0 = W.cold_start(x0[b:b + 1], np.zeros(2), 30, 30)[0]
        out = host_solve(100, 30, 30, z0, x0[b], params[b], lb, ub, max_sqp=300)
        assert out["status"] == 0 and out["viol"] < 1e-6
        assert abs(out["cost"] - g["cost"][b]) < 1e-7 * max(1.0, abs(g["cost"][b]))
        cmd = out["z"][90:92]
        assert np.abs(cmd - g["cmd"][b]).max() < 1e-5 * max(1.0, np.abs(g["cmd"][b]).max())
