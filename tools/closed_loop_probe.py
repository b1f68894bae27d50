"""Closed loop on the device vs step by step through optimize(), cold and with OSQP warm start, with per-phase cycle counters
(quadrotor ph=20, batch 4096).  usage: python tools/closed_loop_probe.py"""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import libmpc_b200 as L
from libmpc_b200 import workloads as W
PH, B, K = 20, 4096, 8
NAMES = ["setup", "factorize", "admm", "info", "polish_prep", "polish_factor", "polish_solve", "unpack"]
for warm in (False, True):
    c = W.build_quadrotor_controller(L, PH, B, 250)
    c.setOptimizerParameters(L.LParameters(maximum_iteration=250, enable_warm_start=warm))
    x0, r = W.quadrotor_inputs(0, B)
    yref = np.zeros((B, 12, PH)); yref[:, 2, :] = r[:, None]
    c.setReferences(yref, np.zeros((4, PH)), np.zeros((4, PH)))
    c.closed_loop(x0, np.zeros((B, 4)), 2)
    for rep in range(2):
        t = time.perf_counter(); out = c.closed_loop(x0, np.zeros((B, 4)), K); dt = time.perf_counter() - t
        print(json.dumps(dict(warm=warm, rep=rep, ms_per_step=1e3 * dt / K, solves_per_s=B * K / dt)), flush=True)
    # step-by-step through the public API with timing of each step
    x = x0.copy(); u = np.zeros((B, 4)); Ad, Bd = W.quadrotor_model()
    c.profile()
    ts = []
    for k in range(6):
        t = time.perf_counter(); o = c.optimize(x, u); ts.append(1e3 * (time.perf_counter() - t))
        pf = c.profile(fetch=True).astype(float)[:, :8]
        print(json.dumps(dict(warm=warm, step=k, ms=round(ts[-1], 2), iters=float(o.iterations.mean()), itmax=int(o.iterations.max()), rho=float(o.rho_updates.mean()),
                              phases={n: round(float(v) / 1e6, 2) for n, v in zip(NAMES, pf.mean(axis=0))})), flush=True)
        x = x @ Ad.T + o.cmd @ Bd.T; u = o.cmd.copy()
    del c
