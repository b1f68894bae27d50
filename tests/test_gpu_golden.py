"""GPU: the CUDA path, through the C ABI, against the committed golden fixtures (tests/golden/): the reference's own
known-answer values (reference_kats.json) and the frozen oracle vectors (*.npz).  No oracle code runs here.
Tolerances: commands 1e-5 relative (north_star); status / iteration / rho-update / polish bookkeeping exact; NLMPC
values 1e-12 relative, finite-difference derivatives 5e-7 abs + 1e-6 rel (see test_gpu_nlmpc.py)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KATS = json.load(open(os.path.join(GOLD, "reference_kats.json")))
W_Y = np.array([0, 0, 10, 10, 10, 10, 0, 0, 0, 5, 5, 5], float)


def _quadrotor(L, ph, batch):
    """quadrotor_ex.cpp:19-83 through the mirrored mpc::LMPC<> setters (model matrices stored in the fixture)."""
    g = np.load(os.path.join(GOLD, "lmpc_quadrotor.npz"))
    Ad, Bd = g["Ad"], g["Bd"]
    c = L.LMPC(12, 4, 4, 12, ph, ph, batch=batch)
    assert c.setStateSpaceModel(Ad, Bd, np.eye(12))
    assert c.setObjectiveWeights(W_Y, np.full(4, 0.1), np.zeros(4), (0, ph))
    xmin = np.full(12, -np.inf); xmax = np.full(12, np.inf)
    xmin[:2] = -np.pi / 6; xmax[:2] = np.pi / 6; xmin[5] = -1
    assert c.setStateBounds(xmin, xmax, (0, ph))
    assert c.setInputBounds(np.full(4, 9.6 - 10.5916), np.full(4, 13 - 10.5916), (0, ph))
    c.setOptimizerParameters(L.LParameters(maximum_iteration=250))
    return c


def test_reference_quadrotor_command_gpu():
    import libmpc_b200 as L
    k = KATS["lmpc_quadrotor_first_command"]
    c = _quadrotor(L, 10, 1)
    yr = np.zeros(12); yr[2] = 1.0
    assert c.setReferences(yr, np.zeros(4), np.zeros(4), (0, 10))
    res = c.optimize(np.zeros(12), np.zeros(4))
    g = np.array(k["cmd"])
    assert np.linalg.norm(res.cmd[0] - g) <= k["rel_tol"] * np.linalg.norm(g)


@pytest.mark.parametrize("ph", [10, 20])
def test_lmpc_fixture_gpu(ph):
    import libmpc_b200 as L
    g = np.load(os.path.join(GOLD, "lmpc_quadrotor.npz"))
    x0, r = g[f"ph{ph}_x0"], g[f"ph{ph}_r"]
    B = len(x0)
    c = _quadrotor(L, ph, B)
    yref = np.zeros((B, 12, ph)); yref[:, 2, :] = r[:, None]
    assert c.setReferences(yref, np.zeros((4, ph)), np.zeros((4, ph)))
    res = c.optimize(x0, np.zeros((B, 4)))
    gc = g[f"ph{ph}_cmd"]
    assert np.abs(res.cmd - gc).max(axis=1).max() < 1e-5 * np.abs(gc).max()
    for b in range(B):
        assert np.abs(res.cmd[b] - gc[b]).max() <= 1e-5 * max(np.abs(gc[b]).max(), 1e-12), (b, res.cmd[b], gc[b])
    meta = g[f"ph{ph}_meta"]
    assert np.array_equal(res.solver_status, meta[:, 0]) and np.array_equal(res.status, meta[:, 1])
    assert np.array_equal(res.iterations, meta[:, 2]) and np.array_equal(res.rho_updates, meta[:, 3])
    assert np.array_equal(res.status_polish, meta[:, 4])
    assert np.abs(res.cost - g[f"ph{ph}_cost"]).max() < 1e-7 * max(1.0, np.abs(g[f"ph{ph}_cost"]).max())
    seq = c.getOptimalSequence()
    assert np.abs(seq.state - g[f"ph{ph}_state"]).max() < 1e-6
    assert np.abs(seq.input - g[f"ph{ph}_input"]).max() < 1e-6


def test_reference_vanderpol_constraint_gpu():
    import libmpc_b200 as L
    k = KATS["nlmpc_vanderpol_dynamics_constraint"]
    out = L.nlmpc_eval(L.SYS_VANDERPOL, 2, 2, np.arange(7.0)[None], np.zeros((1, 2)), np.array([0.01]), want=("ceq", "Jeq"))
    assert np.abs(out["ceq"][0] - np.array(k["c"])).max() < k["abs_tol"]
    assert np.abs(out["Jeq"][0] - np.array(k["J"])).max() < k["abs_tol"]


@pytest.mark.parametrize("name", ["vanderpol", "oscnet4", "ugv"])
def test_nlmpc_eval_fixture_gpu(name):
    import libmpc_b200 as L
    g = np.load(os.path.join(GOLD, "nlmpc_eval.npz"))
    system, ph, ch = (int(v) for v in g[f"{name}_dims"])
    out = L.nlmpc_eval(system, ph, ch, g[f"{name}_z"], g[f"{name}_x0"], g[f"{name}_params"])
    fv = g[f"{name}_f"]
    assert np.allclose(out["f"], fv, rtol=1e-12, atol=0)
    assert np.allclose(out["ceq"], g[f"{name}_ceq"], rtol=1e-12, atol=1e-13)
    assert np.allclose(out["cin"], g[f"{name}_cin"], rtol=1e-12, atol=1e-13)
    for b in range(len(fv)):
        assert np.allclose(out["grad"][b], g[f"{name}_grad"][b], rtol=1e-6, atol=5e-7 * max(1.0, abs(fv[b])))
    assert np.allclose(out["Jeq"], g[f"{name}_Jeq"], rtol=1e-6, atol=5e-7)
    assert np.allclose(out["Jin"], g[f"{name}_Jin"], rtol=1e-6, atol=5e-7)
    assert np.array_equal(out["Jeq"] != 0, g[f"{name}_Jeq"] != 0)      # bit-exact sparsity bookkeeping


def test_nlmpc_solve_fixture_gpu():
    """Van der Pol example from 8 seeded states: same local optimum as the SLSQP oracle froze in the fixture."""
    import libmpc_b200 as L
    g = np.load(os.path.join(GOLD, "nlmpc_solve.npz"))
    x0 = g["x0"]
    B, ph, ch, nx, nu = len(x0), 10, 5, 2, 1
    nz = ph * nx + ch * nu + 1
    z0 = np.concatenate([np.tile(x0, (1, ph)), np.zeros((B, ch * nu + 1))], axis=1)
    FLT_INF = float(np.float32(np.inf))
    lb = np.full(nz, -FLT_INF); ub = np.full(nz, FLT_INF); lb[-1] = ub[-1] = 0.0
    out = L.nlmpc_solve(L.SYS_VANDERPOL, ph, ch, z0, x0, np.array([0.1]), lb, ub)
    assert (out["status"] == 0).all() and (out["viol"] < 1e-8).all()
    ok = g["success"].astype(bool)
    assert ok.sum() >= 7
    cmd = out["z"][:, ph * nx:ph * nx + nu]
    assert np.abs(cmd[ok] - g["cmd"][ok]).max() < 1e-5 * max(1.0, np.abs(g["cmd"][ok]).max())
    assert np.abs(out["cost"][ok] - g["cost"][ok]).max() < 1e-7 * max(1.0, np.abs(g["cost"][ok]).max())
    assert np.abs(out["z"][ok] - g["z"][ok]).max() < 1e-4
