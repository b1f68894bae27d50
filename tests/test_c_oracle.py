"""Pins the C oracle (oracle/lmpc_oracle.c, the CPU baseline bench.py times) against the reference's golden vector and
against the numpy oracle: same statuses / iteration counts / polish flags, solutions equal to rounding."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle.lmpc_formulation import LMPCFormulation, discretization, quadrotor_formulation
from oracle.osqp_restated import Settings, lmpc_optimize


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()


def test_golden_vector_c_oracle():
    """test/LMPC/test_common.cpp:226-236."""
    f = quadrotor_formulation(10, kat_scalar_rows=True)
    out = c_oracle.solve_batch(f, np.zeros((1, 12)), np.zeros((1, 4)), c_oracle.default_params(max_iter=250))
    golden = np.array([-0.9916, 1.74839, -0.9916, 1.74839])
    assert np.linalg.norm(out["cmd"][0] - golden) <= 1e-4 * np.linalg.norm(golden)
    assert out["solver_status"][0] == 1 and out["status_polish"][0] == 1


@pytest.mark.parametrize("ph", [10, 20])
def test_c_oracle_matches_numpy_oracle(ph):
    f = quadrotor_formulation(ph)
    rng = np.random.default_rng(ph)
    B = 5
    x0 = rng.uniform(-1, 1, (B, 12)) * np.array([0.2, 0.2, 0.5, 0.5, 0.5, 0.5] + [0.3] * 6)
    r = rng.uniform(0.5, 1.5, B)
    yref = np.zeros((B, ph, 12)); yref[:, :, 2] = r[:, None]
    out = c_oracle.solve_batch(f, x0, np.zeros((B, 4)), c_oracle.default_params(max_iter=250), yref_batch=yref, nthreads=3)
    for b in range(B):
        yr = np.zeros(12); yr[2] = r[b]
        f.set_references(yr, np.zeros(4), np.zeros(4))
        ref = lmpc_optimize(f, x0[b], np.zeros(4), Settings(max_iter=250))
        assert out["solver_status"][b] == ref["solver_status"] and out["iters"][b] == ref["iter"]
        assert out["rho_updates"][b] == ref["rho_updates"] and out["status_polish"][b] == ref["status_polish"]
        assert np.abs(out["x"][b] - ref["x"]).max() < 1e-9 and np.abs(out["y"][b] - ref["y"]).max() < 1e-8
        assert abs(out["cost"][b] - ref["cost"]) < 1e-8 * max(1, abs(ref["cost"]))


def test_c_oracle_small_system_disturbance_scalar_movblock():
    nx, nu, ndu, ny, ph, ch = 2, 1, 2, 3, 7, 4
    A = np.array([[0, 1.0], [0, 2.0]]); Bc = np.array([[0.0], [1.0]])
    Ad, Bdm = discretization(A, Bc, 0.05)
    rng = np.random.default_rng(3)
    f = LMPCFormulation(nx, nu, ndu, ny, ph, ch)
    f.set_state_space_model(Ad, Bdm, rng.standard_normal((ny, nx)))
    f.set_disturbances(0.1 * rng.standard_normal((nx, ndu)), 0.1 * rng.standard_normal((ny, ndu)))
    f.set_objective_weights(rng.uniform(0.5, 2, (ny, ph)), rng.uniform(0.05, 0.2, (nu, ph)), rng.uniform(0, 0.1, (nu, ph)))
    f.set_input_bounds(np.full((nu, ch), -2.0), np.full((nu, ch), 2.0))
    f.set_state_bounds(np.full(nx, -5.0), np.full(nx, 5.0))
    f.set_scalar_constraint(-0.5, 0.6, np.ones(nx), np.ones(nu))
    f.set_references(rng.standard_normal((ny, ph)), 0.1 * rng.standard_normal((nu, ph)), 0.01 * rng.standard_normal((nu, ph)))
    f.set_exogenous_inputs(0.3 * rng.standard_normal((ndu, ph)))
    x0 = rng.uniform(-0.3, 0.3, (4, nx)); u0 = rng.uniform(-0.2, 0.2, (4, nu))
    out = c_oracle.solve_batch(f, x0, u0, c_oracle.default_params(max_iter=4000))
    for b in range(4):
        ref = lmpc_optimize(f, x0[b], u0[b], Settings(max_iter=4000))
        assert out["solver_status"][b] == ref["solver_status"] and out["iters"][b] == ref["iter"]
        assert np.abs(out["cmd"][b] - ref["cmd"]).max() < 1e-9
