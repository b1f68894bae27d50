"""Closed-loop throughput on the device (b200mpc_lmpc_closed_loop): quadrotor ph=20, batch 4096, K control steps with and
without OSQP warm start; one JSON line each.  The whole loop (K solves + K plant steps) is one host call with one final sync."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import libmpc_b200 as L
from libmpc_b200 import workloads as W

PH, B, K = 20, 4096, 8
for warm in (False, True):
    c = W.build_quadrotor_controller(L, PH, B, 250)
    c.setOptimizerParameters(L.LParameters(maximum_iteration=250, enable_warm_start=warm))
    x0, r = W.quadrotor_inputs(0, B)
    yref = np.zeros((B, 12, PH)); yref[:, 2, :] = r[:, None]
    c.setReferences(yref, np.zeros((4, PH)), np.zeros((4, PH)))
    c.closed_loop(x0, np.zeros((B, 4)), 2)                       # warm-up
    t = time.perf_counter()
    out = c.closed_loop(x0, np.zeros((B, 4)), K)
    dt = time.perf_counter() - t
    print(json.dumps(dict(workload="quadrotor LMPC ph=20 closed loop on the device", batch=B, control_steps=K, warm_start=warm,
                          solves_per_s=B * K / dt, ms_per_control_step=1e3 * dt / K,
                          iterations_mean_per_step=[float(v) for v in out["iterations"].mean(axis=1)],
                          success=float((out["status"] == 0).mean()))), flush=True)
    del c
