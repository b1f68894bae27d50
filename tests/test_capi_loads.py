"""CPU-side checks of the drop-in boundary: the in-tree C-ABI library loads and exports every symbol that
include/b200mpc.h declares; no compute is attempted (there is no GPU here and no CPU fallback by design)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200mpc.h")).read()
    return sorted(set(re.findall(r"\b(b200mpc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    import libmpc_b200 as L
    lib = ctypes.CDLL(L.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for s in declared:
        assert hasattr(lib, s), f"missing export {s}"
    assert sorted(L.EXPORTED_SYMBOLS) == declared


def test_fails_loudly_without_gpu():
    import libmpc_b200 as L
    lib = L.load_library()
    if lib.b200mpc_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        L.LMPC(2, 1, 0, 2, 3, 3)


def test_default_params_match_reference():
    """mpc::LParameters defaults (include/mpc/Types.hpp:99-160) + OSQP v0.6.3 defaults."""
    import libmpc_b200 as L
    lib = L.load_library()
    p = L._Params()
    lib.b200mpc_lmpc_default_params(ctypes.byref(p))
    assert (p.maximum_iteration, p.enable_warm_start, p.alpha, p.rho) == (100, 0, 1.6, 1e-6)
    assert (p.eps_rel, p.eps_abs, p.eps_prim_inf, p.eps_dual_inf) == (1e-4, 1e-4, 1e-3, 1e-3)
    assert (p.adaptive_rho, p.polish, p.sigma, p.delta) == (1, 1, 1e-6, 1e-6)
    assert (p.scaling, p.check_termination, p.adaptive_rho_interval, p.polish_refine_iter) == (10, 25, 25, 3)
