"""CPU pinning of the SQP specification (tests/nlmpc_sqp_reference.py, what the CUDA kernel implements) against the
SLSQP oracle (oracle/nlmpc_slsqp.py, the stand-in for the NLopt call of NLOptimizer.hpp:519) on the reference's example
problems.  Both converge to the same local optimum of the reference's (finite-difference) NLP; the gradients carry
~1e-8 finite-difference noise, so optima agree to ~1e-5 in z and ~1e-8 relative in cost."""
import numpy as np
import pytest

from oracle import nlmpc_slsqp as S
from oracle.nlmpc_formulation import ugv_formulation, vanderpol_formulation
from nlmpc_sqp_reference import sqp_solve


def test_vanderpol_example_optimum():
    """examples/vanderpol_ex.cpp:59-71: x0 = (0, 1), u0 = 0, hard constraints."""
    f = vanderpol_formulation()
    lb, ub = S.default_bounds(f, True)
    x0 = np.array([0.0, 1.0])
    z0 = S.initial_guess(f, x0, np.zeros(1), lb=lb, ub=ub)
    ref = S.solve(f, x0, z0, lb, ub)
    got = sqp_solve(f, x0, z0, lb, ub)
    assert ref["success"]
    assert np.abs(got["z"] - ref["z"]).max() < 1e-5
    assert abs(got["cost"] - ref["cost"]) < 1e-8 * abs(ref["cost"])
    assert np.abs(got["cmd"] - ref["cmd"]).max() < 1e-5 * max(1.0, np.abs(ref["cmd"]).max())
    assert got["viol"] < 1e-8


def test_ugv_example_optimum():
    """examples/ugv_ex.cpp with the soft obstacle constraints (slack free in [0, inf))."""
    f = ugv_formulation(10, 10, v_pref=(0.6, 0.8))
    lb, ub = S.default_bounds(f, False)
    lb[-1] = 0.0
    x0 = np.zeros(4)
    z0 = S.initial_guess(f, x0, np.zeros(2), lb=lb, ub=ub)
    ref = S.solve(f, x0, z0, lb, ub)
    got = sqp_solve(f, x0, z0, lb, ub)
    assert np.abs(got["z"] - ref["z"]).max() < 1e-4
    assert abs(got["cost"] - ref["cost"]) < 1e-6 * max(1.0, abs(ref["cost"]))
    assert got["viol"] < 1e-8


def test_block_diagonal_bfgs_variant_reaches_the_same_optimum():
    """Groundwork for the stage-structured kernel: with a per-stage block-diagonal damped BFGS the SQP reaches the
    optimum of the dense variant (and of SLSQP) in no more major iterations (profiles/r01_block_bfgs_experiment.txt)."""
    from nlmpc_sqp_reference import stage_groups
    f = vanderpol_formulation()
    lb, ub = S.default_bounds(f, True)
    for x0 in (np.array([0.0, 1.0]), np.array([0.8, -0.4])):
        z0 = S.initial_guess(f, x0, np.zeros(1), lb=lb, ub=ub)
        dense = sqp_solve(f, x0, z0, lb, ub)
        block = sqp_solve(f, x0, z0, lb, ub, bfgs_groups=stage_groups(f))
        assert np.abs(block["z"] - dense["z"]).max() < 5e-6 and abs(block["cost"] - dense["cost"]) < 1e-8 * max(1.0, abs(dense["cost"]))
        assert block["viol"] < 1e-8 and block["nit"] <= dense["nit"]
