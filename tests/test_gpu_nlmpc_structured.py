"""GPU: the stage-structured NLMPC solve kernel (libmpc_b200/csrc/nlmpc_structured.cuh) through the C ABI.

  * same optimum as the dense kernel and as the SLSQP oracle (stand-in for NLopt's LD_SLSQP, NLOptimizer.hpp:519 -- PARITY
    UNPINNED upstream): first command 1e-5 relative, cost 1e-7 relative;
  * major-iteration counts equal to the host emulation of the same source (tests/test_nlmpc_structured_host.py) up to the
    finite-difference noise tail -- the GPU runs what the CPU tests checked;
  * BASELINE configs[2] (unicycle nx3 nu2 Tph30, batch 64) against the committed SLSQP fixture.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import nlmpc_slsqp as S
from oracle.nlmpc_formulation import oscnet_formulation, ugv_formulation, vanderpol_formulation
from test_nlmpc_structured_host import host_solve

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cmd(f, z):
    return z[..., f.ph * f.nx:f.ph * f.nx + f.nu]


@pytest.fixture(autouse=True)
def _restore_solver():
    import libmpc_b200 as L
    yield
    L.nlmpc_set_solver(L.NL_SOLVER_AUTO)


@pytest.mark.parametrize("nt", ["32", "64", "128"])
def test_structured_equals_dense_and_host_emulation(nt, monkeypatch):
    import libmpc_b200 as L
    monkeypatch.setenv("B200MPC_NLS_THREADS", nt)
    rng = np.random.default_rng(17)
    cases = [(L.SYS_VANDERPOL, 0, vanderpol_formulation(), np.array([0.1]), True, np.vstack([[0.0, 1.0], rng.uniform(-1.5, 1.5, (5, 2))])),
             (L.SYS_UGV, 3, ugv_formulation(10, 10, v_pref=(0.6, 0.8)), None, False, np.array([[0.0, 0, 0, 0], [0.4, 0.5, 0.6, 0.8]])),
             (L.SYS_OSCNET4, 1, oscnet_formulation(4, 15, 8), np.array([0.1, 1.0, 0.1]), True, rng.uniform(-1, 1, (4, 8)))]
    for sid, hid, f, params, hard, x0 in cases:
        params = f.params if params is None else params
        lb, ub = S.default_bounds(f, hard)
        if not hard:
            lb[-1] = 0.0
        z0 = np.stack([S.initial_guess(f, x, np.zeros(f.nu), lb=lb, ub=ub) for x in x0])
        L.nlmpc_set_solver(L.NL_SOLVER_STRUCTURED)
        st = L.nlmpc_solve(sid, f.ph, f.ch, z0, x0, params, lb, ub)
        L.nlmpc_set_solver(L.NL_SOLVER_DENSE)
        de = L.nlmpc_solve(sid, f.ph, f.ch, z0, x0, params, lb, ub)
        assert (st["status"] == 0).all() and (st["viol"] < 1e-8).all()
        assert np.abs(st["cost"] - de["cost"]).max() < 1e-7 * max(1.0, np.abs(de["cost"]).max())
        assert np.abs(_cmd(f, st["z"]) - _cmd(f, de["z"])).max() < 1e-5 * max(1.0, np.abs(_cmd(f, de["z"])).max())
        for b in range(min(2, len(x0))):
            h = host_solve(hid, f.ph, f.ch, z0[b], x0[b], params, lb, ub)
            assert abs(int(st["iters"][b]) - h["nit"]) <= 3, (st["iters"][b], h["nit"])
            assert np.abs(st["z"][b] - h["z"]).max() < 2e-5
            ref = S.solve(f, x0[b], z0[b], lb, ub)
            if ref["success"]:
                assert abs(st["cost"][b] - ref["cost"]) < 1e-7 * max(1.0, abs(ref["cost"]))
                assert np.abs(_cmd(f, st["z"][b]) - ref["cmd"]).max() < 1e-5 * max(1.0, np.abs(ref["cmd"]).max())


def test_structured_unicycle_cfg2_batch64_vs_slsqp_fixture():
    """BASELINE configs[2] (cold start, 30 stages of curved dynamics, two obstacles) is a problem with several local optima and
    long iteration paths (median 90 major iterations), so two correct local solvers need not land in the same basin: where the
    structured kernel reaches the basin of the SLSQP fixture it must agree at command 1e-5 / cost 1e-7; where it does not, its
    answer must be a local optimum for the oracle too (SLSQP restarted from it stays put)."""
    import libmpc_b200 as L
    from libmpc_b200 import workloads as W
    from user_systems import unicycle_formulation
    sid = L.register_system(W.UNICYCLE_SRC, W.UNICYCLE_TYPE)
    g = np.load(os.path.join(GOLD, "nlmpc_unicycle.npz"))
    B, ph, ch, nx, nu = 64, 30, 30, 3, 2
    x0, params = W.unicycle_inputs(0, B)
    z0 = W.cold_start(x0, np.zeros(nu), ph, ch)
    lb, ub = W.soft_bounds(ph * nx + ch * nu + 1)
    L.nlmpc_set_solver(L.NL_SOLVER_STRUCTURED)
    out = L.nlmpc_solve(sid, ph, ch, z0, x0, params, lb, ub, max_sqp=300)
    conv = out["status"] == 0
    assert conv.sum() >= 60, np.unique(out["status"], return_counts=True)       # the rest: iteration limit on the longest paths
    assert (out["viol"][conv] < 1e-6).all()
    cmd = out["z"][:8, ph * nx:ph * nx + nu]
    rel_cost = np.abs(out["cost"][:8] - g["cost"]) / np.maximum(1.0, np.abs(g["cost"]))
    rel_cmd = np.abs(cmd - g["cmd"]).max(axis=1) / np.maximum(1.0, np.abs(g["cmd"]).max(axis=1))
    same = rel_cost < 1e-3
    assert same.sum() >= 6, rel_cost
    assert rel_cost[same].max() < 1e-7 and rel_cmd[same].max() < 1e-5, (rel_cost, rel_cmd)
    for b in np.nonzero(~same)[0]:
        assert conv[b]
        f = unicycle_formulation(params=params[b])
        ref = S.solve(f, x0[b], out["z"][b], lb, ub, maxiter=100)
        assert ref["cost"] > out["cost"][b] * (1 - 1e-4), (b, ref["cost"], out["cost"][b])     # nothing better nearby
    assert np.isfinite(out["cost"]).all() and (out["cost"] > 0).all() and (out["cost"][conv] < 400).sum() >= 48


def test_structured_rejects_what_it_cannot_solve():
    """A system with user equality constraints has dense rows: forcing the structured kernel is a loud error, the automatic
    choice falls back to the dense kernel."""
    import libmpc_b200 as L
    from user_systems import PENDULUM_SRC
    sid = L.register_system(PENDULUM_SRC, "UserPendulum")
    d = L.nlmpc_system_dims(sid, 8)
    nz = 8 * d["nx"] + 8 * d["nu"] + 1
    lb = np.full(nz, -L.FLT_INF); ub = np.full(nz, L.FLT_INF); lb[-1] = ub[-1] = 0.0
    z0 = np.zeros((1, nz)); x0 = np.zeros((1, d["nx"]))
    L.nlmpc_set_solver(L.NL_SOLVER_STRUCTURED)
    with pytest.raises(RuntimeError, match="stage-structured"):
        L.nlmpc_solve(sid, 8, 8, z0, x0, np.zeros(d["nparam"]), lb, ub)
