"""`libmpc_b200.pympcxx` has the surface of the reference's Python module (python/pybind_export.cpp:13-213): same class,
method, field and enum names.  The lists below are transcribed from the `.def(...)` / `.def_readwrite(...)` /
`.value(...)` calls of that file.  The GPU tests run the reference's python/examples through it."""
import numpy as np
import pytest

LMPC_METHODS = ["setOptimizerParameters", "setLoggerLevel", "setLoggerPrefix", "optimize", "getLastResult", "getOptimalSequence",
                "getExecutionStats", "resetStats", "setStateBounds", "setInputBounds", "setOutputBounds", "setStateSpaceModel",
                "setDisturbances", "getSolverWarmStartPrimal", "getSolverWarmStartDual", "setSolverWarmStart", "setObjectiveWeights",
                "setScalarConstraint", "setExogenousInputs", "setReferences"]                       # pybind_export.cpp:93-123
NLMPC_METHODS = ["setDiscretizationSamplingTime", "setInputScale", "setStateScale", "setOptimizerParameters", "setLoggerLevel",
                 "setLoggerPrefix", "optimize", "getLastResult", "getOptimalSequence", "getExecutionStats", "resetStats",
                 "setStateBounds", "setInputBounds", "setOutputBounds", "setObjectiveFunction", "setStateSpaceFunction",
                 "setOutputFunction", "setIneqConFunction", "setEqConFunction"]                      # pybind_export.cpp:59-84


def test_module_surface():
    import libmpc_b200.pympcxx as mpc
    for name in ("LMPC", "NLMPC", "Parameters", "LParameters", "NLParameters", "HorizonSlice", "Result", "OptSequence", "SolutionStats",
                 "ResultStatus", "LoggerLevel"):
        assert hasattr(mpc, name), name
    for m in LMPC_METHODS:
        assert callable(getattr(mpc.LMPC, m)), m
    for m in NLMPC_METHODS:
        assert callable(getattr(mpc.NLMPC, m)), m
    assert [e.name for e in mpc.ResultStatus] == ["SUCCESS", "MAX_ITERATION", "INFEASIBLE", "ERROR", "UNKNOWN"]      # :193-199
    assert mpc.SUCCESS is mpc.ResultStatus.SUCCESS and mpc.UNKNOWN is mpc.ResultStatus.UNKNOWN                       # export_values()
    assert [e.name for e in mpc.LoggerLevel] == ["DEEP", "NORMAL", "ALERT", "NONE"]                                    # :163-167
    p = mpc.LParameters()
    for f in ("maximum_iteration", "time_limit", "enable_warm_start", "alpha", "rho", "eps_rel", "eps_abs", "eps_prim_inf",
              "eps_dual_inf", "verbose", "adaptive_rho", "polish"):                                                   # :129-145
        assert hasattr(p, f), f
    q = mpc.NLParameters()
    for f in ("relative_ftol", "relative_xtol", "absolute_ftol", "absolute_xtol", "hard_constraints"):                # :148-154
        assert hasattr(q, f), f
    s = mpc.HorizonSlice.all()
    assert (s.start, s.end) == (-1, -1) and mpc.HorizonSlice(2, 5).end == 5                                           # :202-204
    st = mpc.SolutionStats([0.1, 0.3], {})
    for f in ("totalSolutionTime", "numberOfSolutions", "minSolutionTime", "maxSolutionTime", "averageSolutionTime",
              "standardDeviation", "solutionsStates"):                                                                # :183-190
        assert hasattr(st, f), f
    assert abs(st.averageSolutionTime.total_seconds() - 0.2) < 1e-9


def test_nlmpc_python_callbacks_are_refused_with_the_way_out():
    import libmpc_b200.pympcxx as mpc
    c = mpc.NLMPC(2, 1, 2, 10, 5, 11, 0)
    for setter in (c.setStateSpaceFunction, c.setObjectiveFunction, c.setIneqConFunction, c.setEqConFunction, c.setOutputFunction):
        with pytest.raises(RuntimeError, match="setSystemSource"):
            setter(lambda *a: None)


@pytest.mark.gpu
def test_reference_example_py_runs_through_the_compat_module():
    """python/examples/example.py (double integrator, ph=10) for 5 closed-loop steps against the CPU oracle."""
    import libmpc_b200.pympcxx as mpc
    from oracle.lmpc_formulation import LMPCFormulation
    from oracle.osqp_restated import Settings, lmpc_optimize
    nx, nu, ndu, ny, N = 4, 2, 1, 4, 10
    lmpc = mpc.LMPC(nx, nu, ndu, ny, N, N)
    A = np.array([[1.0, 0.1, 0, 0], [0, 1.0, 0.1, 0], [0, 0, 1.0, 0.1], [0, 0, 0, 1.0]])
    B = np.array([[0.0, 0], [0, 0], [1.0, 0], [0, 1.0]])
    C = np.eye(4)
    lmpc.setStateSpaceModel(A, B, C)
    lmpc.setObjectiveWeights(np.ones(4), np.ones(2), np.ones(2), mpc.HorizonSlice.all())
    f = LMPCFormulation(nx, nu, ndu, ny, N, N)
    f.set_state_space_model(A, B, C)
    f.set_objective_weights(np.ones((ny, N)), np.ones((nu, N)), np.ones((nu, N)))
    x, u = np.array([2.0, 10.0, 0.0, 0.0]), np.zeros(2)
    for k in range(5):
        res = lmpc.optimize(x, u)
        ref = lmpc_optimize(f, x, u, Settings())
        assert res.status == mpc.ResultStatus(ref["status"]) and res.cmd.shape == (nu,)
        assert np.abs(res.cmd - ref["cmd"]).max() <= 1e-5 * max(np.abs(ref["cmd"]).max(), 1e-12), (k, res.cmd, ref["cmd"])
        u = res.cmd
        x = A @ x + B @ u
    seq = lmpc.getOptimalSequence()
    assert seq.state.shape == (N + 1, nx) and seq.input.shape == (N + 1, nu) and seq.output.shape == (N + 1, ny)
    stats = lmpc.getExecutionStats()
    assert stats.numberOfSolutions == 5 and stats.totalSolutionTime.total_seconds() > 0


@pytest.mark.gpu
def test_reference_example_nl_py_with_cuda_source():
    """python/examples/example_nl.py: the three Python callbacks become one CUDA source; first command of the restated
    example (oracle/nlmpc_slsqp.py: 0.09098442)."""
    import libmpc_b200.pympcxx as mpc
    from user_systems import VANDERPOL_SRC
    nlmpc = mpc.NLMPC(2, 1, 2, 10, 5, 11, 0)
    nlmpc.setLoggerLevel(mpc.LoggerLevel.NORMAL)
    nlmpc.setDiscretizationSamplingTime(0.1)
    p = mpc.NLParameters(); p.maximum_iteration = 100; p.hard_constraints = True
    nlmpc.setOptimizerParameters(p)
    nlmpc.setSystemSource(VANDERPOL_SRC, "UserVanDerPol", [0.1])
    res = nlmpc.optimize(np.array([0.0, 1.0]), np.array([0.0]))
    assert res.status == mpc.SUCCESS and res.cmd.shape == (1,)
    assert abs(res.cmd[0] - 0.09098442) < 1e-6 and abs(res.cost - 11.1952468) < 1e-6
