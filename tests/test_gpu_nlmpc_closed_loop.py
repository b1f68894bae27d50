"""NLMPC closed loop + warm-start shift on the device (second half) and the RK4 helper (N3, second half).

  * b200mpc_nlmpc_closed_loop (guess / repair / shift kernel -> solve -> plant kernel, no host round trip) against the
    reference's usage pattern written out on the host -- optimize() in a loop with the model stepped by the command
    (examples/vanderpol_ex.cpp:76-85) -- whose every step is checked against the SLSQP oracle;
  * b200mpc_nlmpc_rk4 against a line-by-line numpy restatement of mpc::RK4<N>::run (include/mpc/Integrator.hpp:38-56).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import nlmpc_slsqp as S
from oracle.nlmpc_formulation import vanderpol_formulation


def _vdp(x, u):
    return np.array([(1 - x[1] ** 2) * x[0] - x[1] + u[0], x[0]])


def _rk4_reference(f, x, h, steps):
    """Integrator.hpp:38-56 restated (the time argument is not advanced between sub-steps there either)."""
    sol = x.copy()
    for _ in range(steps):
        k1 = f(sol); k2 = f(sol + (h / 2.0) * k1); k3 = f(sol + (h / 2.0) * k2); k4 = f(sol + h * k3)
        sol = sol + h * (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0
    return sol


def test_rk4_matches_integrator_hpp():
    import libmpc_b200 as L
    rng = np.random.default_rng(3)
    x = rng.uniform(-1.5, 1.5, (37, 2)); u = rng.uniform(-0.5, 0.5, (37, 1))
    for h, n in ((0.1, 1), (0.01, 10), (0.05, 3)):
        out = L.nlmpc_rk4(L.SYS_VANDERPOL, x, u, np.array([0.1]), h, n)
        ref = np.stack([_rk4_reference(lambda s, b=b: _vdp(s, u[b]), x[b], h, n) for b in range(len(x))])
        assert np.abs(out - ref).max() < 1e-14 * max(1.0, np.abs(ref).max()) * 10
    # a linear (discrete-matrix) field as dx/dt: exact solution known -> 4th-order accuracy
    from oracle.nlmpc_formulation import ugv_formulation
    f = ugv_formulation(10, 10)
    xs = rng.uniform(-1, 1, (5, 4)); us = rng.uniform(-1, 1, (5, 2))
    A = f.params[:16].reshape(4, 4); Bm = f.params[16:24].reshape(4, 2)
    out = L.nlmpc_rk4(L.SYS_UGV, xs, us, f.params, 1e-3, 4)
    ref = np.stack([_rk4_reference(lambda s, b=b: A @ s + Bm @ us[b], xs[b], 1e-3, 4) for b in range(5)])
    assert np.abs(out - ref).max() < 1e-13


@pytest.mark.parametrize("warm", [True, False])
def test_device_closed_loop_equals_host_loop_and_oracle(warm):
    import libmpc_b200 as L
    f = vanderpol_formulation()
    lb, ub = S.default_bounds(f, True)
    steps, Ts = 5, 0.1
    x0 = np.array([[0.0, 1.0], [1.0, -0.5], [-0.7, 0.3], [0.2, 0.2]])
    B = len(x0)
    dev = L.nlmpc_closed_loop(L.SYS_VANDERPOL, 10, 5, x0, np.zeros((B, 1)), np.array([Ts]), lb, ub, steps, warm_start=warm,
                              plant_mode=L.PLANT_EULER, plant_h=Ts)
    assert (dev["status"] == 0).all()
    # the same loop with the host-side glue (libmpc_b200.NLMPC.optimize, itself pinned to the oracle in test_gpu_nlmpc_solve.py)
    ctl = L.NLMPC(L.SYS_VANDERPOL, 10, 5, batch=B)
    ctl.setSystemParameters(np.array([Ts]))
    p = L.NLParameters(); p.enable_warm_start = warm
    ctl.setOptimizerParameters(p)
    x = x0.copy(); u = np.zeros((B, 1)); prev = [None] * B
    for k in range(steps):
        assert np.abs(dev["x"][k] - x).max() < 1e-6
        r = ctl.optimize(x, u)
        assert np.abs(dev["u"][k] - r.cmd).max() < 1e-6, (k, dev["u"][k], r.cmd)     # finite-difference noise: x differs in the last bit
        assert np.abs(dev["cost"][k] - r.cost).max() < 1e-6 * max(1.0, np.abs(r.cost).max())
        for b in range(B):          # and the SLSQP oracle from the same guess (cold guess when warm start is off)
            z0 = S.initial_guess(f, x[b], u[b], prev=prev[b] if warm else None, slack=0.0, lb=lb, ub=ub)
            ref = S.solve(f, x[b], z0, lb, ub)
            if ref["success"]:
                assert np.abs(dev["u"][k, b] - ref["cmd"]).max() < 1e-5 * max(1.0, np.abs(ref["cmd"]).max())
            prev[b] = ctl.opt_vector[b].copy()
        u = r.cmd.copy()
        x = np.stack([x[b] + Ts * _vdp(x[b], u[b]) for b in range(B)])
    assert np.abs(dev["x"][steps] - x).max() < 1e-6


def test_device_closed_loop_discrete_plant_with_move_blocking():
    """ugv_ex.cpp:143-166 (discrete model stepped with the command), soft constraints, and a ch < ph oscillator network through
    the RK4 plant: trajectories equal the host-side loop."""
    import libmpc_b200 as L
    from oracle.nlmpc_formulation import oscnet_formulation
    f = oscnet_formulation(4, 6, 3)
    params = np.array([0.1, 1.0, 0.1])          # [Ts, mu, k] of oscnet_formulation's defaults
    lb, ub = S.default_bounds(f, True)
    rng = np.random.default_rng(11)
    x0 = rng.uniform(-0.5, 0.5, (3, 8)); steps = 3
    dev = L.nlmpc_closed_loop(L.SYS_OSCNET4, 6, 3, x0, np.zeros((3, 4)), params, lb, ub, steps, warm_start=True,
                              plant_mode=L.PLANT_RK4, plant_substeps=2, plant_h=0.05)
    ctl = L.NLMPC(L.SYS_OSCNET4, 6, 3, batch=3)
    ctl.setSystemParameters(params)
    p = L.NLParameters(); p.enable_warm_start = True
    ctl.setOptimizerParameters(p)
    x = x0.copy(); u = np.zeros((3, 4))
    for k in range(steps):
        r = ctl.optimize(x, u)
        assert np.abs(dev["u"][k] - r.cmd).max() < 1e-7
        u = r.cmd.copy()
        x = L.nlmpc_rk4(L.SYS_OSCNET4, x, u, params, 0.05, 2)
        assert np.abs(dev["x"][k + 1] - x).max() < 1e-7


def test_device_closed_loop_user_system_unicycle():
    """A user-defined (NVRTC) discrete system through the device loop: the unicycle of BASELINE configs[2] at a short horizon;
    the plant kernel is compiled from the same source as the solver kernels."""
    import libmpc_b200 as L
    from libmpc_b200 import workloads as W
    sid = L.register_system(W.UNICYCLE_SRC, W.UNICYCLE_TYPE)
    ph = ch = 8
    x0, params = W.unicycle_inputs(0, 4)
    lb, ub = W.soft_bounds(ph * 3 + ch * 2 + 1)
    dev = L.nlmpc_closed_loop(sid, ph, ch, x0, np.zeros((4, 2)), params, lb, ub, 3, warm_start=True, plant_mode=L.PLANT_DISCRETE,
                              max_sqp=200)
    Ts = 0.1
    for k in range(3):          # x+ = x + Ts [v cos(th), v sin(th), w]  
        x, u = dev["x"][k], dev["u"][k]
        xn = x + Ts * np.stack([u[:, 0] * np.cos(x[:, 2]), u[:, 0] * np.sin(x[:, 2]), u[:, 1]], axis=1)
        assert np.abs(dev["x"][k + 1] - xn).max() < 1e-12
    assert np.isfinite(dev["cost"]).all()
