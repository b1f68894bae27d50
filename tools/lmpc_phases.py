"""Per-phase cycle shares of the LMPC solve kernel (b200mpc_lmpc_profile counters) and throughput against the batch size
on the bench workload (quadrotor ph=20).  usage: python tools/lmpc_phases.py [batch ...]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import libmpc_b200 as L
from libmpc_b200 import workloads as W

PH, MAXIT = 20, 250
names = ["setup", "factorize", "admm sweeps", "info/termination", "polish prep", "polish factor", "polish solve", "unpack"]
batches = [int(v) for v in sys.argv[1:]] or [2048, 4096, 8192, 16384, 32768]
for B in batches:
    c = W.build_quadrotor_controller(L, PH, B, MAXIT)
    x0, r = W.quadrotor_inputs(0, B)
    yref = np.zeros((B, 12, PH)); yref[:, 2, :] = r[:, None]
    c.setReferences(yref, np.zeros((4, PH)), np.zeros((4, PH)))
    u0 = np.zeros((B, 4))
    c.optimize(x0, u0)
    ts = []
    for _ in range(3):
        t = time.perf_counter(); res = c.optimize(x0, u0); ts.append(time.perf_counter() - t)
    line = dict(batch=B, solves_per_s=B / min(ts), ms=1e3 * min(ts), iters_mean=float(res.iterations.mean()),
                iters_hist={int(k): int(v) for k, v in zip(*np.unique(res.iterations, return_counts=True))})
    if B == batches[0]:
        c.profile()
        c.optimize(x0, u0)
        p = c.profile(fetch=True)[:, :8].astype(float)
        tot = p.sum()
        line["phase_share"] = {n: round(float(p[:, k].sum() / tot), 4) for k, n in enumerate(names)}
        line["cycles_per_solve_mean"] = float(p.sum(axis=1).mean())
    print(json.dumps(line), flush=True)
    del c
