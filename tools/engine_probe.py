"""Engine 1 (warp per controller) vs engine 2 (CTA per controller) on the quadrotor workload: agreement, throughput, phases.
usage: python tools/engine_probe.py [batch ...]   (env: PH, CTA_THREADS)"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, ".")
import libmpc_b200 as L
from libmpc_b200 import workloads as W
PH, MAXIT = int(os.environ.get("PH", 20)), 250
NAMES = ["setup", "factorize", "admm", "info", "polish_prep", "polish_factor", "polish_solve", "unpack"]
for B in [int(v) for v in sys.argv[1:]] or [1, 148, 4096]:
    x0, r = W.quadrotor_inputs(0, B)
    yref = np.zeros((B, 12, PH)); yref[:, 2, :] = r[:, None]
    u0 = np.zeros((B, 4))
    res = {}
    for eng in (2, 1):
        c = W.build_quadrotor_controller(L, PH, B, MAXIT)
        c.set_engine(eng, int(os.environ.get("CTA_THREADS", 0)) if eng == 2 else 0)
        c.setReferences(yref, np.zeros((4, PH)), np.zeros((4, PH)))
        c.optimize(x0, u0)
        c.profile()
        ts = []
        for _ in range(3):
            t = time.perf_counter(); out = c.optimize(x0, u0); ts.append(time.perf_counter() - t)
        pf = c.profile(fetch=True).astype(float)
        p = pf[:, :8]
        sub = pf[:, [8, 9, 10, 11, 12, 13, 14, 15]] / np.maximum(1, out.iterations[:, None])
        res[eng] = out
        print(json.dumps(dict(batch=B, ph=PH, engine=c.get_engine(), ms=[round(1e3 * t, 3) for t in ts], solves_per_s=round(B / min(ts)),
                              iters_mean=float(out.iterations.mean()), cycles_per_solve=float(p.sum(axis=1).mean()),
                              phases={n: round(float(v), 0) for n, v in zip(NAMES, p.mean(axis=0))},
                              admm_subphase_cycles_per_iter=dict(zip(["chain_fwd", "chain_bwd"], [round(float(v)) for v in sub.mean(axis=0)[:2]])) if eng == 2 else None)), flush=True)
        del c
    a, b = res[2], res[1]
    print(json.dumps(dict(batch=B, agree=dict(iters=bool((a.iterations == b.iterations).all()), status=bool((a.solver_status == b.solver_status).all()),
                                             polish=bool((a.status_polish == b.status_polish).all()), rho=bool((a.rho_updates == b.rho_updates).all()),
                                             cmd_maxdiff=float(np.abs(a.cmd - b.cmd).max()), cost_maxrel=float((np.abs(a.cost - b.cost) / np.maximum(1, np.abs(b.cost))).max())))), flush=True)
