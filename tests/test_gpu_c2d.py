"""GPU: batched zero-order-hold discretisation (b200mpc_c2d) against the reference's known answer
(test/test_utils.cpp:10-63) and against scipy.linalg.expm (what oracle/lmpc_formulation.discretization uses).
Tolerance: 1e-13 relative to the largest entry (both are FP64 scaling-and-squaring exponentials)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.lmpc_formulation import discretization as c2d_oracle


def test_reference_double_integrator_kat():
    import libmpc_b200 as L
    dof = 6
    A = np.zeros((12, 12)); A[:dof, dof:] = np.eye(dof)
    B = np.zeros((12, dof)); B[dof:, :] = np.eye(dof)
    Ad, Bd = L.discretization(A, B, 0.02)
    Ad_t = np.eye(12); Ad_t[:dof, dof:] = 0.02 * np.eye(dof)
    Bd_t = np.vstack([0.0002 * np.eye(dof), 0.02 * np.eye(dof)])
    assert np.allclose(Ad, Ad_t, rtol=1e-12, atol=1e-15) and np.allclose(Bd, Bd_t, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("nx,nu", [(2, 1), (4, 2), (12, 4), (20, 6)])
def test_batched_c2d_matches_expm(nx, nu):
    import libmpc_b200 as L
    rng = np.random.default_rng(nx * 10 + nu)
    Bn = 37
    A = rng.standard_normal((Bn, nx, nx)) * 2.0
    A[0] *= 40.0                                            # a stiff one: needs many squarings
    Bm = rng.standard_normal((Bn, nx, nu))
    Ts = rng.uniform(0.001, 0.5, Bn)
    Ad, Bd = L.discretization(A, Bm, Ts)
    for b in range(Bn):
        ra, rb = c2d_oracle(A[b], Bm[b], Ts[b])
        scale = max(np.abs(ra).max(), np.abs(rb).max(), 1.0)
        assert np.abs(Ad[b] - ra).max() <= 1e-12 * scale, (b, np.abs(Ad[b] - ra).max(), scale)
        assert np.abs(Bd[b] - rb).max() <= 1e-12 * scale
    # shared model, per-instance sampling time; and the ugv_ex double integrator (examples/ugv_ex.cpp:36-57)
    A1 = np.zeros((4, 4)); A1[0:2, 2:4] = np.eye(2)
    B1 = np.zeros((4, 2)); B1[2:4, :] = np.eye(2)
    Ad2, Bd2 = L.discretization(A1, B1, np.array([0.1, 0.2]))
    for k, ts in enumerate((0.1, 0.2)):
        ra, rb = c2d_oracle(A1, B1, ts)
        assert np.allclose(Ad2[k], ra, atol=1e-15) and np.allclose(Bd2[k], rb, atol=1e-15)
