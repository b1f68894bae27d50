"""Pins oracle/nlmpc_formulation.py against every known-answer test the reference holds for the NLMPC formulation
(SURVEY.md section 8c).  The reference has NO end-to-end NLMPC test (nothing in test/NLMPC calls optimize()), so solver
parity for NLMPC is unpinned upstream; what is pinned is everything that is handed to the solver."""
import numpy as np
import pytest

from oracle.nlmpc_formulation import NLMPCFormulation, vanderpol_field, vanderpol_formulation, sum_squares_cost


@pytest.mark.parametrize("nx,nu,ph,ch", [(1, 1, 1, 1), (5, 1, 1, 1), (5, 3, 1, 1), (5, 3, 7, 1), (5, 3, 7, 4), (5, 3, 7, 7)])
def test_unwrap_vector_layout(nx, nu, ph, ch):
    """test/NLMPC/test_common.cpp:46-106."""
    f = NLMPCFormulation(nx, nu, 1, ph, ch, 1, 1)
    z = np.arange(ph * nx + nu * ch + 1, dtype=float)
    x0 = -np.arange(nx, dtype=float) - 1
    X, U, e = f.unwrap(z, x0)
    assert np.array_equal(X[0], x0)
    for i in range(1, ph + 1):
        assert np.array_equal(X[i], z[(i - 1) * nx:i * nx])
    u_index = 0
    for i in range(ph + 1):
        if i < ch:
            u_index = ph * nx + i * nu
        assert np.array_equal(U[i], z[u_index:u_index + nu])
    assert e == z[-1]


def test_objective_value_65730():
    """test/NLMPC/test_objective.cpp:9-63 (nx5 nu3 ph=ch=7, z=0..56, x0=0)."""
    f = NLMPCFormulation(5, 3, 1, 7, 7, 0, 0)
    f.obj = sum_squares_cost
    z = np.arange(7 * 5 + 3 * 7 + 1, dtype=float)
    val, _ = f.objective(z, np.zeros(5), want_grad=False)
    assert val == 65730.0
    # the forward-difference gradient approximates 2*z on X and (with the duplicated last row) 2*u / 4*u_last on U
    _, g = f.objective(z, np.zeros(5))
    assert np.allclose(g[:35], 2 * z[:35], rtol=1e-5, atol=1e-3)
    assert np.allclose(g[35:53], 2 * z[35:53], rtol=1e-5, atol=1e-3)
    assert np.allclose(g[53:56], 4 * z[53:56], rtol=1e-5, atol=1e-3)


def test_vanderpol_trapezoidal_residual_and_jacobian():
    """test/NLMPC/test_constraints.cpp:60-142."""
    f = NLMPCFormulation(2, 1, 1, 2, 2, 0, 0)
    f.continuous, f.Ts = True, 0.01
    f.f = vanderpol_field
    z = np.arange(7, dtype=float)
    c, J = f.state_eq(z, np.zeros(2), want_jac=True)
    assert np.all(np.abs(c - np.array([0.035, -1, -2.05, -1.99])) < 1e-3)
    Jexp = np.array([-1, -0.005, 0, 0, 0.01, 0, 0, 0.005, -1, 0, 0, 0, 0, 0, 1, -0.005, -1.04, -0.065, 0, 0.01, 0,
                     0.005, 1, 0.005, -1, 0, 0, 0]).reshape(4, 7)
    assert np.all(np.abs(J - Jexp) < 1e-3)
    c2, J2 = f.state_eq(z, np.zeros(2), want_jac=False)
    assert np.array_equal(c, c2) and not J2.any()


def test_user_constraints_plumbing():
    """test/NLMPC/test_constraints.cpp:144-274: value = x0[0] on every row, Jacobian zero."""
    f = NLMPCFormulation(2, 1, 1, 3, 3, 2, 2)
    f.ineq = lambda X, Y, U, e: np.full(2, X[0, 0])
    f.eq = lambda X, U: np.full(2, X[0, 0])
    z = np.arange(f.nz, dtype=float)
    v, J = f.ineq_con(z, np.array([3.0, 4.0]))
    assert np.all(v == 3.0) and not J.any()
    v, J = f.eq_con(z, np.array([3.0, 4.0]))
    assert np.all(v == 3.0) and not J.any()


def test_move_blocking_gradient_chain():
    """ch<ph: the objective gradient w.r.t. the last control block collects every blocked stage (Iz2u')."""
    f = vanderpol_formulation()
    rng = np.random.default_rng(0)
    z = rng.standard_normal(f.nz)
    _, g = f.objective(z, np.array([0.0, 1.0]))
    X, U, e = f.unwrap(z, np.array([0.0, 1.0]))
    # stages 4..9 share block 4; the reference's last-row pairing puts 2 extra u^2 terms on stage 9
    expected_last = 2 * U[4, 0] * (6 + 1)
    assert abs(g[20 + 4] - expected_last) < 1e-4 * max(1, abs(expected_last))
    vi, Ji = f.ineq_con(z, np.array([0.0, 1.0]))
    assert np.allclose(vi, U[:, 0] - 0.5)
    # ineq Jacobian: row r depends on u block min(r,4); rows 4..9 all on block 4, row 10 (duplicate) on nothing
    assert np.allclose(Ji[:4, 20:24], np.eye(4), atol=1e-6)
    assert np.allclose(Ji[4:10, 24], 1.0, atol=1e-6) and abs(Ji[10, 24]) < 1e-12
