"""CPU: oracle/nlmpc_oracle.c (the compiled restatement used for the NLMPC CPU baseline) against the numpy oracle
(oracle/nlmpc_formulation.py, itself pinned to the reference's known-answer tests) and, through it, the reference's own
Van der Pol constraint KAT (test/NLMPC/test_constraints.cpp:60-142)."""
import json
import os

import numpy as np
import pytest

from oracle import nlmpc_c_oracle as CO
from oracle import nlmpc_slsqp as S
from oracle.nlmpc_formulation import oscnet_formulation, ugv_formulation, vanderpol_formulation

KATS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")))


def _cases():
    fv = vanderpol_formulation(); fv.params = np.array([0.1])
    fo = oscnet_formulation(4, 15, 8); fo.params = np.array([0.1, 1.0, 0.1])
    fu = ugv_formulation(10, 10, v_pref=(0.6, 0.8))
    return [(0, fv, 10, 5), (1, fo, 15, 8), (3, fu, 10, 10)]


def test_reference_vanderpol_constraint_kat():
    k = KATS["nlmpc_vanderpol_dynamics_constraint"]
    out = CO.evaluate(0, 2, 2, np.arange(7.0), np.zeros(2), np.array([0.01]), want=("ceq", "Jeq"))
    assert np.abs(out["ceq"] - np.array(k["c"])).max() < k["abs_tol"]
    assert np.abs(out["Jeq"] - np.array(k["J"])).max() < k["abs_tol"]


@pytest.mark.parametrize("case", range(3))
def test_c_eval_matches_numpy_oracle(case):
    system, f, ph, ch = _cases()[case]
    rng = np.random.default_rng(70 + system)
    for _ in range(3):
        z = rng.standard_normal(f.nz) * 0.7
        z[-1] = abs(z[-1]) * 0.1
        x0 = rng.uniform(-0.5, 0.5, f.nx)
        out = CO.evaluate(system, ph, ch, z, x0, f.params)
        fv, g = f.objective(z, x0)
        c, J = f.state_eq(z, x0)
        ci, Ji = f.ineq_con(z, x0)
        assert abs(out["f"] - fv) <= 1e-12 * max(1.0, abs(fv))
        assert np.allclose(out["ceq"], c, rtol=1e-12, atol=1e-13) and np.allclose(out["cin"], ci, rtol=1e-12, atol=1e-13)
        assert np.allclose(out["grad"], g, rtol=1e-6, atol=5e-7 * max(1.0, abs(fv)))
        assert np.allclose(out["Jeq"], J, rtol=1e-6, atol=5e-7) and np.allclose(out["Jin"], Ji, rtol=1e-6, atol=5e-7)
        assert np.array_equal(out["Jeq"] != 0, J != 0)


def test_c_callback_slsqp_reaches_the_oracle_optimum():
    f = vanderpol_formulation()
    lb, ub = S.default_bounds(f, True)
    for x0 in (np.array([0.0, 1.0]), np.array([0.8, -0.4])):
        z0 = S.initial_guess(f, x0, np.zeros(1), lb=lb, ub=ub)
        a = CO.solve(0, 10, 5, x0, z0, np.array([0.1]), lb, ub)
        b = S.solve(f, x0, z0, lb, ub)
        assert a["success"] and b["success"]
        assert abs(a["cost"] - b["cost"]) < 1e-8 * max(1.0, abs(b["cost"])) and np.abs(a["cmd"] - b["cmd"]).max() < 1e-6
