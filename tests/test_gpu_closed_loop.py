"""GPU: the device-resident closed loop (b200mpc_lmpc_closed_loop) against the same loop run through the
CPU oracle -- optimize(x_k, u_{k-1}) -> u_k -> x_{k+1} = A x_k + B u_k -- with and without OSQP warm start
(LOptimizer.hpp:268-281).  Tolerances: every command of the trajectory within 1e-5 relative (north_star), iteration
counts exact, states within 1e-6."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.lmpc_formulation import quadrotor_formulation, quadrotor_model
from oracle.osqp_restated import Settings, lmpc_optimize


def _controller(L, ph, B, warm):
    f = quadrotor_formulation(ph)
    c = L.LMPC(12, 4, 4, 12, ph, ph, batch=B)
    Ad, Bd = quadrotor_model()
    assert c.setStateSpaceModel(Ad, Bd, np.eye(12))
    assert c.setObjectiveWeights(f.wOutput[:, 1], f.wU[:, 1], f.wDeltaU[:, 0], (0, ph))
    assert c.setStateBounds(f.minX[:, 1], f.maxX[:, 1], (0, ph))
    assert c.setInputBounds(f.minU[:, 0], f.maxU[:, 0], (0, ph))
    assert c.setReferences(f.yRef[:, 0], np.zeros(4), np.zeros(4), (0, ph))
    c.setOptimizerParameters(L.LParameters(maximum_iteration=250, enable_warm_start=warm))
    return f, c, Ad, Bd


@pytest.mark.parametrize("warm", [False, True])
def test_closed_loop_matches_oracle_loop(warm):
    import libmpc_b200 as L
    ph, B, steps = 10, 3, 4
    f, c, Ad, Bd = _controller(L, ph, B, warm)
    rng = np.random.default_rng(21)
    x0 = rng.uniform(-1, 1, (B, 12)) * 0.1
    out = c.closed_loop(x0, np.zeros((B, 4)), steps)
    assert out["x"].shape == (steps + 1, B, 12) and np.array_equal(out["x"][0], x0)
    for b in range(B):
        x, u, prev = x0[b].copy(), np.zeros(4), None
        for k in range(steps):
            r = lmpc_optimize(f, x, u, Settings(max_iter=250, warm_start=warm), warm=prev if warm else None)
            assert out["iterations"][k, b] == r["iter"], (b, k, out["iterations"][k, b], r["iter"])
            assert out["status"][k, b] == r["status"]
            assert np.abs(out["u"][k, b] - r["cmd"]).max() < 1e-5 * max(np.abs(r["cmd"]).max(), 1e-12), (b, k)
            prev = (r["x"], r["y"])
            u = r["cmd"]
            x = Ad @ x + Bd @ u
            assert np.abs(out["x"][k + 1, b] - x).max() < 1e-6


def test_closed_loop_with_a_different_plant_and_batch_order():
    """A perturbed per-instance plant (model mismatch) and the property that the loop of instance b does not depend on the
    rest of the batch."""
    import libmpc_b200 as L
    ph, B, steps = 10, 6, 3
    f, c, Ad, Bd = _controller(L, ph, B, True)
    rng = np.random.default_rng(22)
    x0 = rng.uniform(-1, 1, (B, 12)) * 0.1
    Ap = np.broadcast_to(Ad, (B, 12, 12)) * (1.0 + 0.01 * rng.standard_normal((B, 1, 1)))
    Bp = np.broadcast_to(Bd, (B, 12, 4)).copy()
    out = c.closed_loop(x0, np.zeros((B, 4)), steps, plant=(Ap, Bp))
    for k in range(steps):
        for b in range(B):
            assert np.abs(out["x"][k + 1, b] - (Ap[b] @ out["x"][k, b] + Bp[b] @ out["u"][k, b])).max() < 1e-12
    perm = rng.permutation(B)
    f2, c2, _, _ = _controller(L, ph, B, True)
    out2 = c2.closed_loop(x0[perm], np.zeros((B, 4)), steps, plant=(Ap[perm], Bp[perm]))
    assert np.array_equal(out2["u"], out["u"][:, perm]) and np.array_equal(out2["iterations"], out["iterations"][:, perm])
