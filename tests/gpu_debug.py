"""Manual GPU diagnostic (not a pytest file): staged comparison of the CUDA path against the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import libmpc_b200 as L
from oracle.lmpc_formulation import quadrotor_formulation, quadrotor_model
from oracle.osqp_restated import Settings, lmpc_optimize

def mk(ph, B):
    f = quadrotor_formulation(ph)
    c = L.LMPC(12, 4, 4, 12, ph, ph, batch=B)
    Ad, Bd = quadrotor_model()
    c.setStateSpaceModel(Ad, Bd, np.eye(12))
    c.setObjectiveWeights(f.wOutput[:, 1], f.wU[:, 1], f.wDeltaU[:, 0], (0, ph))
    c.setStateBounds(f.minX[:, 1], f.maxX[:, 1], (0, ph))
    c.setInputBounds(f.minU[:, 0], f.maxU[:, 0], (0, ph))
    c.setReferences(f.yRef[:, 0], np.zeros(4), np.zeros(4), (0, ph))
    return f, c

ph = 10
f, c = mk(ph, 1)
for (mi, pol, ada) in ((1, False, False), (25, False, False), (25, False, True), (50, False, True), (250, False, True), (250, True, True)):
    p = L.LParameters(maximum_iteration=mi, polish=pol, adaptive_rho=ada)
    c.setOptimizerParameters(p)
    res = c.optimize(np.zeros(12), np.zeros(4))
    wx, wy = c.getSolverWarmStartPrimal()[0], c.getSolverWarmStartDual()[0]
    r = lmpc_optimize(f, np.zeros(12), np.zeros(4), Settings(max_iter=mi, polish=pol, adaptive_rho=ada))
    print(f"mi={mi} pol={pol} ada={ada}: gpu st={res.solver_status[0]} it={res.iterations[0]} ru={res.rho_updates[0]} ps={res.status_polish[0]} cost={res.cost[0]:.9g}"
          f" | ora st={r['solver_status']} it={r['iter']} ru={r['rho_updates']} ps={r['status_polish']} cost={r['cost']:.9g}"
          f" | dx={np.abs(wx - r['x']).max():.3e} dy={np.abs(wy - r['y']).max():.3e} cmd={res.cmd[0]}")
# throughput probe
for ph, B in ((10, 1024), (20, 4096)):
    f, c = mk(ph, B)
    c.setOptimizerParameters(L.LParameters(maximum_iteration=250))
    rng = np.random.default_rng(1)
    x0 = rng.uniform(-1, 1, (B, 12)) * np.array([0.2, 0.2, 0.5, 0.5, 0.5, 0.5] + [0.3] * 6)
    c.profile()
    for rep in range(3):
        t = time.time(); res = c.optimize(x0, np.zeros((B, 4))); dt = time.time() - t
        print(f"ph={ph} B={B}: {dt*1e3:.1f} ms -> {B/dt:.0f} solves/s; iters mean {res.iterations.mean():.1f} max {res.iterations.max()} status {np.unique(res.solver_status, return_counts=True)} info {c.info()}")
    pr = c.profile(fetch=True).astype(float)
    names = ['setup', 'factor', 'sweeps', 'info', 'polprep', 'polfac', 'polsolve', 'unpack']
    tot = pr[:, :8].sum(1).mean()
    sn = ['f_prologue', 'f_acq', 'f_B||Q1', 'f_C||Q2', 'b_setup', 'b_acq', 'b_A||R1', 'b_B||R2']
    print('  sweep split (cycles/instance):', {n: int(v) for n, v in zip(sn, pr[:, 8:].mean(0))})
    print('  phase cycles/instance:', {n: int(v) for n, v in zip(names, pr[:, :8].mean(0))}, 'total', int(tot), ' per-iter sweep', int(pr[:,2].mean()/res.iterations.mean()))
