"""GPU parity of the batched NLMPC problem evaluation (K5) against oracle/nlmpc_formulation.py for the reference's three
example systems.  Values must agree to rounding (1e-12 relative); finite-difference derivatives are the difference of
two O(1) numbers divided by ~1.5e-8, so two correct evaluations in different summation orders differ by ~1e-16/1.5e-8:
the tolerance for gradients / Jacobians is 5e-7 absolute + 1e-6 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.nlmpc_formulation import oscnet_formulation, ugv_formulation, vanderpol_formulation


def _check(f, system, params, ph, ch, B, seed, x0_scale=1.0):
    import libmpc_b200 as L
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((B, f.nz)) * 0.7
    z[:, -1] = np.abs(z[:, -1]) * 0.1
    x0 = rng.uniform(-1, 1, (B, f.nx)) * x0_scale
    out = L.nlmpc_eval(system, ph, ch, z, x0, params)
    for b in range(B):
        fv, g = f.objective(z[b], x0[b])
        c, J = f.state_eq(z[b], x0[b])
        ci, Ji = f.ineq_con(z[b], x0[b])
        assert abs(out["f"][b] - fv) <= 1e-12 * max(1, abs(fv))
        assert np.allclose(out["ceq"][b], c, rtol=1e-12, atol=1e-13)
        assert np.allclose(out["cin"][b], ci, rtol=1e-12, atol=1e-13)
        tol = dict(rtol=1e-6, atol=5e-7 * max(1.0, abs(fv)))
        assert np.allclose(out["grad"][b], g, **tol), np.abs(out["grad"][b] - g).max()
        assert np.allclose(out["Jeq"][b], J, rtol=1e-6, atol=5e-7), np.abs(out["Jeq"][b] - J).max()
        assert np.allclose(out["Jin"][b], Ji, rtol=1e-6, atol=5e-7), np.abs(out["Jin"][b] - Ji).max()
        # bit-exact bookkeeping: identical sparsity pattern of the dynamics Jacobian (which blocks are touched)
        assert np.array_equal(out["Jeq"][b] != 0, J != 0)


def test_vanderpol_kat_on_gpu():
    """test/NLMPC/test_constraints.cpp:60-142 through the CUDA path (nx2 nu1 ph=ch=2, Ts=0.01, z=0..6, x0=0)."""
    import libmpc_b200 as L
    out = L.nlmpc_eval(L.SYS_VANDERPOL, 2, 2, np.arange(7.0)[None], np.zeros((1, 2)), np.array([0.01]), want=("ceq", "Jeq"))
    assert np.all(np.abs(out["ceq"][0] - np.array([0.035, -1, -2.05, -1.99])) < 1e-3)
    Jexp = np.array([-1, -0.005, 0, 0, 0.01, 0, 0, 0.005, -1, 0, 0, 0, 0, 0, 1, -0.005, -1.04, -0.065, 0, 0.01, 0,
                     0.005, 1, 0.005, -1, 0, 0, 0]).reshape(4, 7)
    assert np.all(np.abs(out["Jeq"][0] - Jexp) < 1e-3)


def test_vanderpol_example_shape():
    import libmpc_b200 as L
    _check(vanderpol_formulation(), L.SYS_VANDERPOL, np.array([0.1]), 10, 5, 16, 1)


@pytest.mark.parametrize("N,system,ph,ch", [(4, 1, 15, 8), (6, 2, 20, 10)])
def test_networked_oscillators(N, system, ph, ch):
    _check(oscnet_formulation(N, ph, ch), system, np.array([0.1, 1.0, 0.1]), ph, ch, 6, 40 + N)


def test_ugv_with_obstacles():
    import libmpc_b200 as L
    f = ugv_formulation(10, 10, v_pref=(0.6, 0.8))
    _check(f, L.SYS_UGV, f.params, 10, 10, 8, 30, x0_scale=0.5)
