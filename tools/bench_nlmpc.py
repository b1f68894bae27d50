"""Throughput of the batched NLMPC solve (K6/K7) on the BASELINE.json NLMPC shapes.  Secondary to bench.py (whose line is
the quadrotor LMPC metric): prints one JSON line per workload with solves/s measured through the C ABI with host buffers
(H2D/D2H and the workspace allocation inside the timed region), the mean SQP / ADMM iteration counts and the fraction of
instances that converged."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import libmpc_b200 as L
from oracle import nlmpc_slsqp as S
from oracle.nlmpc_formulation import oscnet_formulation, ugv_formulation, vanderpol_formulation


def run(name, system, f, ph, ch, batch, hard, x0_lo, x0_hi, reps=3):
    lb, ub = S.default_bounds(f, hard)
    if not hard:
        lb[-1] = 0.0
    rng = np.random.default_rng(0)
    x0 = rng.uniform(x0_lo, x0_hi, (batch, f.nx))
    z0 = np.concatenate([np.tile(x0, (1, ph)), np.zeros((batch, ch * f.nu + 1))], axis=1)
    L.nlmpc_solve(system, ph, ch, z0[:64], x0[:64], f.params, lb, ub, max_sqp=200)      # warm-up
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        out = L.nlmpc_solve(system, ph, ch, z0, x0, f.params, lb, ub, max_sqp=200)
        ts.append(time.perf_counter() - t)
    t = min(ts)
    print(json.dumps(dict(workload=name, nz=int(f.nz), batch=batch, solves_per_s=batch / t, ms_per_batch=1e3 * t,
                          smem_resident=bool(L.load_library().b200mpc_nlmpc_solve_smem_bytes(system, ph, ch) <= 227 * 1024),
                          converged=float((out["status"] == 0).mean()), sqp_iters=float(out["iters"].mean()),
                          admm_iters=float(out["qp_iters"].mean()), max_viol=float(out["viol"].max()))), flush=True)


def run_unicycle(batch=1024):
    """BASELINE configs[2] at its stated shape (user-defined system, NVRTC), cold start."""
    from libmpc_b200 import workloads as W
    sid = L.register_system(W.UNICYCLE_SRC, W.UNICYCLE_TYPE)
    x0, params = W.unicycle_inputs(0, batch)
    lb, ub = W.soft_bounds(151)
    z0 = W.cold_start(x0, np.zeros(2), 30, 30)
    L.nlmpc_solve(sid, 30, 30, z0[:64], x0[:64], params[:64], lb, ub, max_sqp=300)
    t = time.perf_counter()
    out = L.nlmpc_solve(sid, 30, 30, z0, x0, params, lb, ub, max_sqp=300)
    t = time.perf_counter() - t
    print(json.dumps(dict(workload="unicycle nx3 nu2 ph30 ch30 (configs[2])", nz=151, batch=batch, solves_per_s=batch / t, ms_per_batch=1e3 * t,
                          converged=float((out["status"] == 0).mean()), sqp_iters=float(out["iters"].mean()),
                          admm_iters=float(out["qp_iters"].mean()), max_viol=float(out["viol"].max()), max_cost=float(out["cost"].max()))), flush=True)


if __name__ == "__main__":
    import os
    print(json.dumps(dict(solver=os.environ.get("B200MPC_NLMPC_SOLVER", "auto"), threads=os.environ.get("B200MPC_NLS_THREADS", "default"))), flush=True)
    run_unicycle()
    f = vanderpol_formulation(); f.params = np.array([0.1])
    run("vanderpol_ex nx2 nu1 ph10 ch5", L.SYS_VANDERPOL, f, 10, 5, 8192, True, -1.5, 1.5)
    f = ugv_formulation(10, 10, v_pref=(0.6, 0.8))
    run("ugv_ex nx4 nu2 ph10 ch10", L.SYS_UGV, f, 10, 10, 1024, False, -0.3, 0.6)
    f = ugv_formulation(30, 30, v_pref=(0.6, 0.8))
    run("ugv_ex nx4 nu2 ph30 ch30", L.SYS_UGV, f, 30, 30, 1024, False, -0.3, 0.6, reps=1)
    f = oscnet_formulation(4, 15, 8); f.params = np.array([0.1, 1.0, 0.1])
    run("networked_oscillators_ex nx8 nu4 ph15 ch8", L.SYS_OSCNET4, f, 15, 8, 8192, True, -1.0, 1.0, reps=1)
