import sys, time
import numpy as np
sys.path.insert(0, ".")
import libmpc_b200 as L
from oracle import nlmpc_slsqp as S
from oracle.nlmpc_formulation import vanderpol_formulation
f = vanderpol_formulation(); f.params = np.array([0.1])
lb, ub = S.default_bounds(f, True)
x0 = np.random.default_rng(0).uniform(-1.5, 1.5, (8192, 2))
z0 = np.concatenate([np.tile(x0, (1, 10)), np.zeros((8192, 6))], axis=1)
for k in range(6):
    t = time.perf_counter(); out = L.nlmpc_solve(L.SYS_VANDERPOL, 10, 5, z0, x0, f.params, lb, ub, max_sqp=200); dt = time.perf_counter() - t
    print(k, "%.1f ms" % (1e3 * dt), "%.0f solves/s" % (8192 / dt), out["iters"].mean(), out["qp_iters"].mean(), float((out["status"] == 0).mean()))
