"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs.  Tolerances: control outputs within 1e-5 relative error (BASELINE.json north_star);
status / iteration / polish bookkeeping exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.lmpc_formulation import LMPCFormulation, discretization, quadrotor_formulation, quadrotor_model
from oracle.osqp_restated import Settings, lmpc_optimize

REL = 1e-5


def _mk_quadrotor(L, ph, batch, kat_scalar_rows=False, max_iter=250):
    f = quadrotor_formulation(ph, kat_scalar_rows=kat_scalar_rows)
    c = L.LMPC(12, 4, 4, 12, ph, ph, batch=batch)
    Ad, Bd = quadrotor_model()
    assert c.setStateSpaceModel(Ad, Bd, np.eye(12))
    assert c.setDisturbances(np.zeros((12, 4)), np.zeros((12, 4)))
    assert c.setObjectiveWeights(f.wOutput[:, 1], f.wU[:, 1], f.wDeltaU[:, 0], (0, ph))
    assert c.setStateBounds(f.minX[:, 1], f.maxX[:, 1], (0, ph))
    assert c.setInputBounds(f.minU[:, 0], f.maxU[:, 0], (0, ph))
    assert c.setOutputBounds(np.full(12, -np.inf), np.full(12, np.inf), (0, ph))
    if kat_scalar_rows:
        assert c.setScalarConstraint(-np.inf, np.inf, np.ones(12), np.ones(4), (-1, -1))
    assert c.setReferences(f.yRef[:, 0], np.zeros(4), np.zeros(4), (0, ph))
    c.setOptimizerParameters(L.LParameters(maximum_iteration=max_iter))
    return f, c


def _relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


def test_quadrotor_golden_vector_gpu():
    """test/LMPC/test_common.cpp:89-237 through the CUDA path."""
    import libmpc_b200 as L
    f, c = _mk_quadrotor(L, 10, 1, kat_scalar_rows=True)
    res = c.optimize(np.zeros(12), np.zeros(4))
    golden = np.array([-0.9916, 1.74839, -0.9916, 1.74839])
    assert np.linalg.norm(res.cmd[0] - golden) <= 1e-4 * np.linalg.norm(golden)
    r = lmpc_optimize(f, np.zeros(12), np.zeros(4), Settings(max_iter=250))
    assert _relerr(res.cmd[0], r["cmd"]) < REL
    assert res.solver_status[0] == r["solver_status"] and res.status[0] == r["status"]
    assert res.iterations[0] == r["iter"] and res.rho_updates[0] == r["rho_updates"] and res.status_polish[0] == r["status_polish"]
    assert abs(res.cost[0] - r["cost"]) < 1e-7 * max(1, abs(r["cost"]))
    seq = c.getOptimalSequence()
    assert np.abs(seq.state[0] - r["state"]).max() < 1e-6
    assert np.abs(seq.input[0] - r["input"]).max() < 1e-6
    assert np.abs(seq.output[0] - r["output"]).max() < 1e-6


@pytest.mark.parametrize("ph,B", [(10, 24), (20, 12)])
def test_quadrotor_batch_vs_oracle(ph, B):
    import libmpc_b200 as L
    f, c = _mk_quadrotor(L, ph, B)
    rng = np.random.default_rng(20 + ph)
    x0 = rng.uniform(-1, 1, (B, 12)) * np.array([0.2, 0.2, 0.5, 0.5, 0.5, 0.5] + [0.3] * 6)
    x0[:, 0:2] = np.clip(x0[:, 0:2], -np.pi / 6, np.pi / 6)
    r_ref = rng.uniform(0.5, 1.5, B)
    yref = np.zeros((B, 12, ph))
    yref[:, 2, :] = r_ref[:, None]
    assert c.setReferences(yref, np.zeros((4, ph)), np.zeros((4, ph)))
    res = c.optimize(x0, np.zeros((B, 4)))
    wx, wy = c.getSolverWarmStartPrimal(), c.getSolverWarmStartDual()
    for b in range(B):
        yr = np.zeros(12); yr[2] = r_ref[b]
        f.set_references(yr, np.zeros(4), np.zeros(4))
        r = lmpc_optimize(f, x0[b], np.zeros(4), Settings(max_iter=250))
        assert _relerr(res.cmd[b], r["cmd"]) < REL, (b, res.cmd[b], r["cmd"])
        assert res.solver_status[b] == r["solver_status"], (b, res.solver_status[b], r["solver_status"])
        assert res.iterations[b] == r["iter"], (b, res.iterations[b], r["iter"])
        assert res.rho_updates[b] == r["rho_updates"] and res.status_polish[b] == r["status_polish"]
        assert np.abs(wx[b] - r["x"]).max() < 1e-6 * max(1, np.abs(r["x"]).max())
        assert np.abs(wy[b] - r["y"]).max() < 1e-5 * max(1, np.abs(r["y"]).max())
        # active-set bookkeeping: same sign pattern of the polished duals
        assert np.array_equal(np.sign(np.round(wy[b], 9)), np.sign(np.round(r["y"], 9)))


def test_max_iter_and_unpolished_path():
    """maximum_iteration=100 (the reference default) stops some instances before convergence: OSQP then skips polish;
    the raw ADMM iterate must still match the oracle's."""
    import libmpc_b200 as L
    ph, B = 10, 8
    f, c = _mk_quadrotor(L, ph, B, max_iter=50)
    rng = np.random.default_rng(5)
    x0 = rng.uniform(-1, 1, (B, 12)) * 0.2
    res = c.optimize(x0, np.zeros((B, 4)))
    for b in range(B):
        r = lmpc_optimize(f, x0[b], np.zeros(4), Settings(max_iter=50))
        assert res.solver_status[b] == r["solver_status"]
        assert res.status[b] == r["status"] and bool(res.is_feasible[b]) == r["is_feasible"]
        assert res.iterations[b] == r["iter"]
        assert _relerr(res.cmd[b], r["cmd"]) < 1e-6


def test_small_system_with_disturbance_and_scalar_constraint():
    """nx=2,nu=1 system of test/LMPC/test_constraints.cpp:95-167 + measured disturbance + ch<ph + per-stage weights."""
    import libmpc_b200 as L
    nx, nu, ndu, ny, ph, ch = 2, 1, 2, 3, 7, 4
    A = np.array([[0, 1.0], [0, 2.0]]); Bc = np.array([[0.0], [1.0]])
    Ad, Bdm = discretization(A, Bc, 0.05)
    rng = np.random.default_rng(3)
    Cm = rng.standard_normal((ny, nx))
    Bdist = 0.1 * rng.standard_normal((nx, ndu)); Ddist = 0.1 * rng.standard_normal((ny, ndu))
    OW = rng.uniform(0.5, 2, (ny, ph)); UW = rng.uniform(0.05, 0.2, (nu, ph)); DUW = rng.uniform(0, 0.1, (nu, ph))
    umeas = 0.3 * rng.standard_normal((ndu, ph))
    yref = rng.standard_normal((ny, ph)); uref = 0.1 * rng.standard_normal((nu, ph)); duref = 0.01 * rng.standard_normal((nu, ph))
    f = LMPCFormulation(nx, nu, ndu, ny, ph, ch)
    f.set_state_space_model(Ad, Bdm, Cm); f.set_disturbances(Bdist, Ddist)
    f.set_objective_weights(OW, UW, DUW)
    f.set_input_bounds(np.full((nu, ch), -2.0), np.full((nu, ch), 2.0))
    f.set_state_bounds(np.full(nx, -5.0), np.full(nx, 5.0))
    f.set_scalar_constraint(-0.5, 0.6, np.ones(nx), np.ones(nu))
    f.set_references(yref, uref, duref); f.set_exogenous_inputs(umeas)
    B = 6
    c = L.LMPC(nx, nu, ndu, ny, ph, ch, batch=B)
    assert c.setStateSpaceModel(Ad, Bdm, Cm) and c.setDisturbances(Bdist, Ddist)
    assert c.setObjectiveWeights(OW, UW, DUW)
    assert c.setInputBounds(np.full((nu, ch), -2.0), np.full((nu, ch), 2.0))
    assert c.setStateBounds(np.full(nx, -5.0), np.full(nx, 5.0), L.HorizonSlice.all())
    assert c.setScalarConstraint(-0.5, 0.6, np.ones(nx), np.ones(nu), L.HorizonSlice.all())
    assert c.setReferences(yref, uref, duref) and c.setExogenousInputs(umeas)
    c.setOptimizerParameters(L.LParameters(maximum_iteration=4000))
    x0 = rng.uniform(-0.3, 0.3, (B, nx)); u0 = rng.uniform(-0.2, 0.2, (B, nu))
    res = c.optimize(x0, u0)
    seq = c.getOptimalSequence()
    for b in range(B):
        r = lmpc_optimize(f, x0[b], u0[b], Settings(max_iter=4000))
        assert res.solver_status[b] == r["solver_status"] and res.iterations[b] == r["iter"], (b, res.iterations[b], r["iter"])
        assert _relerr(res.cmd[b], r["cmd"]) < REL
        assert np.abs(seq.output[b] - r["output"]).max() < 1e-6


def test_warm_start_roundtrip():
    import libmpc_b200 as L
    ph, B = 10, 4
    f, c = _mk_quadrotor(L, ph, B)
    p = L.LParameters(maximum_iteration=250, enable_warm_start=True)
    c.setOptimizerParameters(p)
    rng = np.random.default_rng(9)
    x0 = rng.uniform(-1, 1, (B, 12)) * 0.1
    r1 = c.optimize(x0, np.zeros((B, 4)))
    wx, wy = c.getSolverWarmStartPrimal(), c.getSolverWarmStartDual()
    x1 = x0 + 0.01
    r2 = c.optimize(x1, r1.cmd)
    for b in range(B):
        o = lmpc_optimize(f, x1[b], r1.cmd[b], Settings(max_iter=250, warm_start=True), warm=(wx[b], wy[b]))
        assert r2.solver_status[b] == o["solver_status"] and r2.iterations[b] == o["iter"]
        assert _relerr(r2.cmd[b], o["cmd"]) < REL


def test_infeasible_and_error_paths():
    import libmpc_b200 as L
    ph, B = 5, 2
    f, c = _mk_quadrotor(L, ph, B, max_iter=4000)
    # out-of-range lastU -> infeasible QP (ProblemBuilder.hpp:735-749); with mpc::inf bounds present the IEEE
    # inf*0=NaN quirk suppresses the primal certificate, exactly as the oracle reproduces it
    u0 = np.array([[0.0] * 4, [50.0] * 4])
    res = c.optimize(np.zeros((B, 12)), u0)
    for b in range(B):
        r = lmpc_optimize(f, np.zeros(12), u0[b], Settings(max_iter=4000))
        assert res.solver_status[b] == r["solver_status"], (b, res.solver_status[b], r["solver_status"])
        assert res.status[b] == r["status"] and bool(res.is_feasible[b]) == r["is_feasible"]
    # lower bound > upper bound: osqp_setup refuses the data -> ERROR, previous command kept, sequences zeroed
    prev = res.cmd.copy()
    assert c.setStateBounds(np.full(12, 1.0), np.full(12, -1.0), L.HorizonSlice.all())
    res2 = c.optimize(np.zeros((B, 12)), np.zeros((B, 4)))
    assert np.all(res2.status == L.ERROR) and np.all(np.isinf(res2.cost))
    assert np.array_equal(np.nan_to_num(res2.cmd, nan=-7), np.nan_to_num(prev, nan=-7))
    assert np.all(c.getOptimalSequence().state == 0)


def test_infeasible_stage0_box_runs_to_max_iter():
    """x0 outside the stage-0 state box: the QP is infeasible, but the du rows of stages <= ch always carry mpc::inf
    (ProblemBuilder.hpp:784-792), so OSQP's certificate sum is NaN and the reference runs to MAX_ITER; same here."""
    import libmpc_b200 as L
    nx, nu, ndu, ny, ph, ch = 2, 1, 0, 2, 4, 2
    Ad = np.array([[1, 0.1], [0, 1.0]]); Bd = np.array([[0.0], [0.1]])
    f = LMPCFormulation(nx, nu, ndu, ny, ph, ch)
    f.set_state_space_model(Ad, Bd, np.eye(2))
    f.set_objective_weights(np.ones(2), np.ones(1), np.zeros(1))
    f.set_state_bounds(np.array([1.0, -10]), np.array([2.0, 10]))
    f.set_input_bounds(np.full((1, ch), -1.0), np.full((1, ch), 1.0))
    c = L.LMPC(nx, nu, ndu, ny, ph, ch, batch=1)
    c.setStateSpaceModel(Ad, Bd, np.eye(2)); c.setObjectiveWeights(np.ones(2), np.ones(1), np.zeros(1), L.HorizonSlice.all())
    c.setStateBounds(np.array([1.0, -10]), np.array([2.0, 10]), L.HorizonSlice.all())
    c.setInputBounds(np.full((1, ch), -1.0), np.full((1, ch), 1.0))
    c.setOptimizerParameters(L.LParameters(maximum_iteration=200))
    res = c.optimize(np.zeros(2), np.zeros(1))
    r = lmpc_optimize(f, np.zeros(2), np.zeros(1), Settings(max_iter=200))
    assert res.solver_status[0] == r["solver_status"] == -2 and res.iterations[0] == r["iter"]
    assert res.status[0] == L.MAX_ITERATION and res.is_feasible[0]      # LOptimizer.hpp:344 quirk
    assert _relerr(res.cmd[0], r["cmd"]) < 1e-4


def test_no_fallback_symbols_loaded():
    """The product path is the CUDA library: it must be the in-tree .so and report a GPU."""
    import libmpc_b200 as L
    lib = L.load_library()
    assert lib.b200mpc_device_count() >= 1
    assert L.LIB_PATH.endswith("libmpc_b200/libb200mpc.so")


def test_input_bound_tail_after_slice_set_with_ch_lt_ph():
    """the matrix setter replicates column ch-1 into the tail ch..ph-1 (ProblemBuilder.hpp:397-413) but the per-index
    setter writes ONE column and leaves the tail alone (ProblemBuilder.hpp:469-477).  Tight bounds everywhere, then step ch-1
    loosened through the slice setter: the tail must stay tight on the device, as in the reference."""
    import libmpc_b200 as L
    nx, nu, ph, ch = 2, 1, 6, 3
    Ad, Bdm = discretization(np.array([[0, 1.0], [0, 0.0]]), np.array([[0.0], [1.0]]), 0.1)
    f = LMPCFormulation(nx, nu, 0, nx, ph, ch)
    f.set_state_space_model(Ad, Bdm, np.eye(nx))
    f.set_objective_weights(np.array([10.0, 1.0]), np.array([0.01]), np.array([0.0]))
    f.set_input_bounds(np.full((nu, ch), -0.2), np.full((nu, ch), 0.2))
    f.set_input_bounds_at(ch - 1, np.array([-5.0]), np.array([5.0]))
    f.set_references(np.array([1.0, 0.0]), np.zeros(nu), np.zeros(nu))
    c = L.LMPC(nx, nu, 0, nx, ph, ch, batch=2)
    assert c.setStateSpaceModel(Ad, Bdm, np.eye(nx))
    assert c.setObjectiveWeights(np.array([10.0, 1.0]), np.array([0.01]), np.array([0.0]), L.HorizonSlice.all())
    assert c.setInputBounds(np.full((nu, ch), -0.2), np.full((nu, ch), 0.2))
    assert c.setInputBounds(np.array([-5.0]), np.array([5.0]), (ch - 1, ch))
    assert c.setReferences(np.array([1.0, 0.0]), np.zeros(nu), np.zeros(nu), L.HorizonSlice.all())
    assert np.array_equal(c._st["UMin"][..., 0, :], f.minU[0]) and np.array_equal(c._st["UMax"][..., 0, :], f.maxU[0])
    assert f.maxU[0, ch - 1] == 5.0 and (f.maxU[0, ch:] == 0.2).all()
    c.setOptimizerParameters(L.LParameters(maximum_iteration=2000))
    x0 = np.array([[0.0, 0.0], [0.3, -0.1]]); u0 = np.zeros((2, nu))
    res = c.optimize(x0, u0)
    seq = c.getOptimalSequence()
    for b in range(2):
        r = lmpc_optimize(f, x0[b], u0[b], Settings(max_iter=2000))
        assert res.solver_status[b] == r["solver_status"] and res.iterations[b] == r["iter"]
        assert _relerr(res.cmd[b], r["cmd"]) < REL
        assert np.abs(seq.input[b] - r["input"]).max() < 1e-6
    # the tail obeys the TIGHT bound (it would reach 5 if the device had re-replicated column ch-1)
    # (an un-polished ADMM solution honours a bound to ~eps_abs = 1e-4, hence the 1e-3)
    assert np.abs(seq.input[:, ch + 1:, 0]).max() <= 0.2 + 1e-3 and np.abs(seq.input[:, :, 0]).max() > 0.25


def test_two_live_handles_of_different_sizes_interleaved():
    """the dynamic-shared-memory attribute belongs to the kernel function, not to a handle; a second, smaller handle
    must not break the first one's next launch."""
    import libmpc_b200 as L
    rng = np.random.default_rng(5)

    def make(ph, B):
        nx, nu = 3, 2
        A = np.eye(nx) + 0.1 * rng.standard_normal((nx, nx)); Bm = rng.standard_normal((nx, nu))
        f = LMPCFormulation(nx, nu, 0, nx, ph, ph)
        f.set_state_space_model(A, Bm, np.eye(nx)); f.set_objective_weights(np.ones(nx), 0.1 * np.ones(nu), np.zeros(nu))
        f.set_input_bounds(np.full(nu, -1.0), np.full(nu, 1.0)); f.set_references(np.ones(nx) * 0.2, np.zeros(nu), np.zeros(nu))
        c = L.LMPC(nx, nu, 0, nx, ph, ph, batch=B)
        c.setStateSpaceModel(A, Bm, np.eye(nx)); c.setObjectiveWeights(np.ones(nx), 0.1 * np.ones(nu), np.zeros(nu), L.HorizonSlice.all())
        c.setInputBounds(np.full(nu, -1.0), np.full(nu, 1.0), L.HorizonSlice.all())
        c.setReferences(np.ones(nx) * 0.2, np.zeros(nu), np.zeros(nu), L.HorizonSlice.all())
        c.setOptimizerParameters(L.LParameters(maximum_iteration=1000))
        return f, c, rng.uniform(-0.3, 0.3, (B, nx))
    fa, ca, xa = make(12, 5)
    ra0 = ca.optimize(xa, np.zeros((5, 2)))
    fb, cb, xb = make(2, 3)                                   # much smaller shared-memory footprint
    rb = cb.optimize(xb, np.zeros((3, 2)))
    ra1 = ca.optimize(xa, np.zeros((5, 2)))                   # the older handle again
    rb1 = cb.optimize(xb, np.zeros((3, 2)))
    assert np.array_equal(ra0.cmd, ra1.cmd) and np.array_equal(rb.cmd, rb1.cmd)
    for f, res, x in ((fa, ra1, xa), (fb, rb1, xb)):
        for b in range(len(x)):
            r = lmpc_optimize(f, x[b], np.zeros(2), Settings(max_iter=1000))
            assert res.iterations[b] == r["iter"] and _relerr(res.cmd[b], r["cmd"]) < REL
