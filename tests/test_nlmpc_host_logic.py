"""CPU: host-side logic of the NLMPC mirror (no GPU call is made): the bound vectors handed to the solver have the layout the
reference's own tests pin (test/NLMPC/test_nloptimizer.cpp:9-121), the slack bound follows `hard_constraints`
(NLOptimizer.hpp:160-190), and the initial guess / bound repair / one-stage shift follow NLOptimizer::run (:431-510,705-716)
-- checked against the oracle's restatement of the same lines (oracle/nlmpc_slsqp.initial_guess)."""
import numpy as np
import pytest

from oracle import nlmpc_slsqp as S
from oracle.nlmpc_formulation import NLMPCFormulation

SYSTEMS = [(0, 2, 1), (1, 8, 4), (3, 4, 2)]          # (system id, nx, nu) of the built-in device functors
HORIZONS = [(1, 1), (7, 1), (7, 4), (7, 7)]           # the (Tph, Tch) pairs of the reference test


@pytest.mark.parametrize("system,nx,nu", SYSTEMS)
@pytest.mark.parametrize("ph,ch", HORIZONS)
def test_default_and_set_bounds_layout(system, nx, nu, ph, ch):
    import libmpc_b200 as L
    c = L.NLMPC(system, ph, ch, batch=3)
    assert (c.nx, c.nu) == (nx, nu) and c.lb.size == ph * nx + ch * nu + 1
    # "Checking default state and input bounds": +-infinity everywhere but the slack
    assert np.all(c.lb[:-1] == -np.inf) and np.all(c.ub[:-1] == np.inf)
    # "Checking state and input bounds": vector + HorizonSlice::all
    assert c.setStateBounds(np.full(nx, -1.0), np.full(nx, 1.0), L.HorizonSlice.all())
    assert c.setInputBounds(np.full(nu, -1.0), np.full(nu, 1.0), L.HorizonSlice.all())
    for i in range(ph):
        for j in range(nx):
            assert c.lb[i * nx + j] == -1.0 and c.ub[i * nx + j] == 1.0
    for i in range(ch):
        for j in range(nu):
            assert c.lb[ph * nx + i * nu + j] == -1.0 and c.ub[ph * nx + i * nu + j] == 1.0
    # slices and the matrix form
    if ph >= 2:
        assert c.setStateBounds(np.full(nx, -2.0), np.full(nx, 2.0), L.HorizonSlice(1, 2))
        assert c.lb[nx] == -2.0 and c.lb[0] == -1.0
        assert not c.setStateBounds(np.full(nx, -2.0), np.full(nx, 2.0), L.HorizonSlice(2, 1))
    lo = -np.arange(1.0, nu * ch + 1).reshape(ch, nu).T
    assert c.setInputBounds(lo, -lo)
    assert np.array_equal(c.lb[ph * nx:ph * nx + ch * nu], lo.T.ravel())
    assert c.setOutputBounds(np.zeros(2), np.zeros(2)) is False                  # ignored upstream too (NLMPC.hpp:342-349)


def test_slack_bound_follows_hard_constraints():
    import libmpc_b200 as L
    c = L.NLMPC(L.SYS_UGV, 10, 10)
    assert c.lb[-1] == 0.0 and c.ub[-1] == 0.0                                    # hard constraints pin the slack
    c.setOptimizerParameters(L.NLParameters(hard_constraints=False))
    assert c.lb[-1] == 0.0 and c.ub[-1] == np.inf


@pytest.mark.parametrize("ph,ch", [(10, 5), (7, 7), (6, 1)])
def test_initial_guess_repair_and_shift_match_the_oracle(ph, ch):
    import libmpc_b200 as L
    B = 4
    c = L.NLMPC(L.SYS_UGV, ph, ch, batch=B)
    c.setOptimizerParameters(L.NLParameters(enable_warm_start=True, hard_constraints=False))
    c.setInputBounds(np.full(2, -0.5), np.full(2, 0.8), L.HorizonSlice.all())
    f = NLMPCFormulation(4, 2, 4, ph, ch)
    rng = np.random.default_rng(3)
    x0, u0 = rng.standard_normal((B, 4)), rng.uniform(-2, 2, (B, 2))          # some u0 outside the bounds -> repaired
    z_cold = c._initial_guess(x0, u0)
    for b in range(B):
        assert np.array_equal(z_cold[b], S.initial_guess(f, x0[b], u0[b], lb=c.lb, ub=c.ub))
    # warm: previous optimum, shifted by one stage, slack carried over
    prev = rng.uniform(-0.4, 0.7, (B, c.nz))
    c.opt_vector, c.is_first_iteration, c.current_slack = prev.copy(), False, prev[:, -1].copy()
    z_warm = c._initial_guess(x0, u0)
    for b in range(B):
        assert np.allclose(z_warm[b], S.initial_guess(f, x0[b], u0[b], prev=prev[b], slack=prev[b, -1], lb=c.lb, ub=c.ub), rtol=0, atol=0)
