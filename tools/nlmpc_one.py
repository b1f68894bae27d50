"""One batched NLMPC solve of a named workload (for ncu captures): python tools/nlmpc_one.py <vdp|ugv10|ugv30|osc4> <batch>."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import libmpc_b200 as L
from oracle import nlmpc_slsqp as S
from oracle.nlmpc_formulation import oscnet_formulation, ugv_formulation, vanderpol_formulation

name, batch = sys.argv[1], int(sys.argv[2])
if name == "vdp":
    f = vanderpol_formulation(); f.params = np.array([0.1]); system, ph, ch, hard, lo, hi = L.SYS_VANDERPOL, 10, 5, True, -1.5, 1.5
elif name == "ugv10":
    f = ugv_formulation(10, 10, v_pref=(0.6, 0.8)); system, ph, ch, hard, lo, hi = L.SYS_UGV, 10, 10, False, -0.3, 0.6
elif name == "ugv30":
    f = ugv_formulation(30, 30, v_pref=(0.6, 0.8)); system, ph, ch, hard, lo, hi = L.SYS_UGV, 30, 30, False, -0.3, 0.6
else:
    f = oscnet_formulation(4, 15, 8); f.params = np.array([0.1, 1.0, 0.1]); system, ph, ch, hard, lo, hi = L.SYS_OSCNET4, 15, 8, True, -1.0, 1.0
lb, ub = S.default_bounds(f, hard)
if not hard:
    lb[-1] = 0.0
x0 = np.random.default_rng(0).uniform(lo, hi, (batch, f.nx))
z0 = np.concatenate([np.tile(x0, (1, ph)), np.zeros((batch, ch * f.nu + 1))], axis=1)
t = time.perf_counter()
out = L.nlmpc_solve(system, ph, ch, z0, x0, f.params, lb, ub, max_sqp=200)
print(name, batch, "%.1f ms" % (1e3 * (time.perf_counter() - t)), "converged", float((out["status"] == 0).mean()), "sqp", out["iters"].mean(), "admm", out["qp_iters"].mean())
