"""Generates the committed golden fixtures under tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

Two kinds of fixture:

* reference_kats.json -- the known-answer values the REFERENCE's own tests hold for this path, transcribed with the
  file:line they come from (libmpc++ is C++ and cannot be built in this image -- Eigen/OSQP/NLopt are absent -- so they
  are transcribed, not regenerated).  tests/test_golden.py checks the oracle against them; tests/test_gpu_golden.py
  checks the CUDA path against them.
* lmpc_quadrotor.npz / nlmpc_eval.npz / nlmpc_solve.npz -- seeded inputs and the outputs of the pinned oracle
  (oracle/osqp_restated.py, oracle/nlmpc_formulation.py, oracle/nlmpc_slsqp.py) on them.  They freeze the oracle (a
  later edit that changes an iterate path shows up as a fixture mismatch) and let the GPU tests run without executing
  the slow numpy oracle at the larger sizes.

Nothing here reads /root/reference at run time.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import nlmpc_slsqp as S                                                     # noqa: E402
from oracle.lmpc_formulation import quadrotor_formulation, quadrotor_model                               # noqa: E402
from oracle.nlmpc_formulation import oscnet_formulation, ugv_formulation, vanderpol_formulation  # noqa: E402
from oracle.osqp_restated import Settings, lmpc_optimize                                # noqa: E402

X0_SCALE = np.array([0.2, 0.2, 0.5, 0.5, 0.5, 0.5] + [0.3] * 6)

REFERENCE_KATS = {
    "_note": "values transcribed from the reference's own tests; paths relative to the libmpc++ repository",
    "lmpc_quadrotor_first_command": {
        "source": "test/LMPC/test_common.cpp:89-237",
        "setup": "quadrotor_ex model, nx12 nu4 ndu4 ny12 ph=ch=10, x0=0 u0=0, yRef=e3, maximum_iteration=250",
        "cmd": [-0.9916, 1.74839, -0.9916, 1.74839], "rel_tol": 1e-4},
    "nlmpc_objective_value": {
        "source": "test/NLMPC/test_objective.cpp:9-63",
        "setup": "nx5 nu3 ph=ch=7, z=0..56, x0=0, cost=sum x^2 + sum u^2 (duplicated last U row included)",
        "value": 65730.0},
    "nlmpc_vanderpol_dynamics_constraint": {
        "source": "test/NLMPC/test_constraints.cpp:60-142",
        "setup": "Van der Pol continuous Ts=0.01, nx2 nu1 ph=ch=2, z=0..6, x0=0",
        "c": [0.035, -1.0, -2.05, -1.99],
        "J": [[-1, -0.005, 0, 0, 0.01, 0, 0], [0.005, -1, 0, 0, 0, 0, 0],
              [1, -0.005, -1.04, -0.065, 0, 0.01, 0], [0.005, 1, 0.005, -1, 0, 0, 0]],
        "abs_tol": 1e-3},
    "nlmpc_unwrap_layout": {
        "source": "test/NLMPC/test_common.cpp:46-106",
        "setup": "z=0..nz-1, x0=-1..-nx; X row0 = x0, rows 1..ph = z blocks; U row i = block i for i<ch, the last block repeated afterwards; slack = z[nz-1]",
        "cases_nx_nu_ph_ch": [[1, 1, 1, 1], [5, 1, 1, 1], [5, 3, 1, 1], [5, 3, 7, 1], [5, 3, 7, 4], [5, 3, 7, 7]]},
}


def lmpc_fixture():
    out = {}
    for ph, B in ((10, 16), (20, 8)):
        f = quadrotor_formulation(ph)
        rng = np.random.default_rng(1000 + ph)
        x0 = rng.uniform(-1, 1, (B, 12)) * X0_SCALE
        x0[:, 0:2] = np.clip(x0[:, 0:2], -np.pi / 6, np.pi / 6)
        x0[0] = 0.0
        r = rng.uniform(0.5, 1.5, B)
        r[0] = 1.0
        cmd = np.zeros((B, 4)); cost = np.zeros(B)
        meta = np.zeros((B, 5), dtype=np.int64)           # solver_status, status, iter, rho_updates, status_polish
        seq_state = np.zeros((B, ph + 1, 12)); seq_input = np.zeros((B, ph + 1, 4))
        for b in range(B):
            yr = np.zeros(12); yr[2] = r[b]
            f.set_references(yr, np.zeros(4), np.zeros(4))
            res = lmpc_optimize(f, x0[b], np.zeros(4), Settings(max_iter=250))
            cmd[b] = res["cmd"]; cost[b] = res["cost"]
            meta[b] = [res["solver_status"], res["status"], res["iter"], res["rho_updates"], res["status_polish"]]
            seq_state[b] = res["state"]; seq_input[b] = res["input"]
        out.update({f"ph{ph}_x0": x0, f"ph{ph}_r": r, f"ph{ph}_cmd": cmd, f"ph{ph}_cost": cost, f"ph{ph}_meta": meta,
                    f"ph{ph}_state": seq_state, f"ph{ph}_input": seq_input})
    out["Ad"], out["Bd"] = quadrotor_model()                      # examples/quadrotor_ex.cpp:19-45
    np.savez_compressed(os.path.join(HERE, "lmpc_quadrotor.npz"), **out)


def nlmpc_cases():
    fv = vanderpol_formulation(); fv.params = np.array([0.1])
    fo = oscnet_formulation(4, 15, 8); fo.params = np.array([0.1, 1.0, 0.1])
    fu = ugv_formulation(10, 10, v_pref=(0.6, 0.8))
    return (("vanderpol", fv, 0, 10, 5, 1.0), ("oscnet4", fo, 1, 15, 8, 1.0), ("ugv", fu, 3, 10, 10, 0.5))


def nlmpc_eval_fixture():
    out = {}
    for name, f, system, ph, ch, x0s in nlmpc_cases():
        B = 4
        rng = np.random.default_rng(2000 + system)
        z = rng.standard_normal((B, f.nz)) * 0.7
        z[:, -1] = np.abs(z[:, -1]) * 0.1
        x0 = rng.uniform(-1, 1, (B, f.nx)) * x0s
        fval = np.zeros(B); grad = np.zeros((B, f.nz))
        ceq = np.zeros((B, ph * f.nx)); Jeq = np.zeros((B, ph * f.nx, f.nz))
        cin = np.zeros((B, f.nineq)); Jin = np.zeros((B, f.nineq, f.nz))
        for b in range(B):
            fval[b], grad[b] = f.objective(z[b], x0[b])
            ceq[b], Jeq[b] = f.state_eq(z[b], x0[b])
            cin[b], Jin[b] = f.ineq_con(z[b], x0[b])
        out.update({f"{name}_z": z, f"{name}_x0": x0, f"{name}_params": f.params, f"{name}_f": fval, f"{name}_grad": grad,
                    f"{name}_ceq": ceq, f"{name}_Jeq": Jeq, f"{name}_cin": cin, f"{name}_Jin": Jin,
                    f"{name}_dims": np.array([system, ph, ch])})
    np.savez_compressed(os.path.join(HERE, "nlmpc_eval.npz"), **out)


def nlmpc_solve_fixture():
    """Van der Pol example (examples/vanderpol_ex.cpp) from 8 seeded initial states: SLSQP-oracle optimum."""
    f = vanderpol_formulation()
    lb, ub = S.default_bounds(f, True)
    rng = np.random.default_rng(3000)
    x0 = rng.uniform(-1.0, 1.0, (8, 2))
    x0[0] = [0.0, 1.0]                                                       # the example's initial state
    z = np.zeros((8, f.nz)); cmd = np.zeros((8, 1)); cost = np.zeros(8); ok = np.zeros(8, dtype=np.int64)
    for b in range(8):
        z0 = S.initial_guess(f, x0[b], np.zeros(1))
        r = S.solve(f, x0[b], z0, lb, ub, maxiter=400)
        z[b] = r["z"]; cmd[b] = r["cmd"]; cost[b] = r["cost"]; ok[b] = int(r["success"])
    np.savez_compressed(os.path.join(HERE, "nlmpc_solve.npz"), x0=x0, z=z, cmd=cmd, cost=cost, success=ok)


def nlmpc_unicycle_fixture(B=8):
    """BASELINE.json configs[2] at its stated shape (unicycle nx3 nu2 Tph=Tch=30, 2 obstacles): the first B
    instances of the bench workload (libmpc_b200.workloads.unicycle_inputs, seed 30), cold start, SLSQP-oracle optimum."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from user_systems import unicycle_formulation
    from libmpc_b200.workloads import unicycle_inputs, cold_start, soft_bounds
    x0, params = unicycle_inputs(0, B)
    f0 = unicycle_formulation(params=params[0])
    lb, ub = soft_bounds(f0.nz)
    z0 = cold_start(x0, np.zeros(2), 30, 30)
    z = np.zeros((B, f0.nz)); cmd = np.zeros((B, 2)); cost = np.zeros(B); ok = np.zeros(B, dtype=np.int64); nit = np.zeros(B, dtype=np.int64)
    for b in range(B):
        f = unicycle_formulation(params=params[b])
        r = S.solve(f, x0[b], z0[b], lb, ub, maxiter=600)
        z[b] = r["z"]; cmd[b] = r["cmd"]; cost[b] = r["cost"]; ok[b] = int(r["success"]); nit[b] = r["nit"]
        print("unicycle", b, r["success"], r["nit"], r["cost"], flush=True)
    np.savez_compressed(os.path.join(HERE, "nlmpc_unicycle.npz"), x0=x0, params=params, z0=z0, z=z, cmd=cmd, cost=cost, success=ok, nit=nit)


def nlmpc_output_fixture():
    """A system with an output map (NLMPC::setOutputFunction): evaluation through Y and OptSequence::output of the optimum."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from user_systems import output_map_formulation
    f = output_map_formulation()
    rng = np.random.default_rng(5)
    B = 4
    z = rng.standard_normal((B, f.nz)) * 0.6; z[:, -1] = 0.0
    x0 = rng.uniform(-0.5, 0.5, (B, 2))
    out = dict(z=z, x0=x0, params=f.params)
    fv = np.zeros(B); g = np.zeros((B, f.nz)); ci = np.zeros((B, f.nineq)); Ji = np.zeros((B, f.nineq, f.nz)); Y = np.zeros((B, f.ph + 1, f.ny))
    for b in range(B):
        fv[b], g[b] = f.objective(z[b], x0[b]); ci[b], Ji[b] = f.ineq_con(z[b], x0[b])
        X, U, e = f.unwrap(z[b], x0[b]); Y[b] = f.output(X, U)
    out.update(f=fv, grad=g, cin=ci, Jin=Ji, Y=Y)
    lb, ub = S.default_bounds(f, True)
    zs = np.zeros((B, f.nz)); Ys = np.zeros((B, f.ph + 1, f.ny)); cost = np.zeros(B); ok = np.zeros(B, dtype=np.int64)
    for b in range(B):
        z0 = S.initial_guess(f, x0[b], np.zeros(1), lb=lb, ub=ub)
        r = S.solve(f, x0[b], z0, lb, ub, maxiter=400)
        zs[b] = r["z"]; cost[b] = r["cost"]; ok[b] = int(r["success"]); Ys[b] = f.output(r["state"], r["input"])
    out.update(sol_z=zs, sol_Y=Ys, sol_cost=cost, sol_success=ok)
    np.savez_compressed(os.path.join(HERE, "nlmpc_output_map.npz"), **out)


if __name__ == "__main__":
    if "--unicycle" in sys.argv:
        nlmpc_unicycle_fixture(); sys.exit(0)
    if "--output-map" in sys.argv:
        nlmpc_output_fixture(); sys.exit(0)
    with open(os.path.join(HERE, "reference_kats.json"), "w") as fh:
        json.dump(REFERENCE_KATS, fh, indent=1)
    lmpc_fixture()
    nlmpc_eval_fixture()
    nlmpc_solve_fixture()
    print("wrote", sorted(p for p in os.listdir(HERE) if p.endswith((".npz", ".json"))))
