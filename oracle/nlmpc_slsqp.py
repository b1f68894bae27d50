"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- the reference's NLMPC solve path restated with SciPy's SLSQP.

The reference hands its NLP to NLopt's LD_SLSQP (include/mpc/NLMPC/NLOptimizer.hpp:64,519); NLopt is a third-party
dependency that is NOT under /root/reference and is pinned to nothing (configure.sh:25-26 clones master).  NLopt's
slsqp.c and SciPy's `minimize(method="SLSQP")` are both translations of Dieter Kraft's SLSQP (DFVLR-FB 88-28, 1988), so
SciPy 1.18 (present in the image) is used as the stand-in for that algorithm; the callbacks it is given are the restated
Objective / Constraints of oracle/nlmpc_formulation.py, i.e. exactly what the reference gives NLopt, finite-difference
gradients included.  Decision-vector initialisation, bound repair and the one-stage warm-start shift follow
NLOptimizer::run (NLOptimizer.hpp:412-510,705-716).

PARITY UNPINNED: the reference has no test that calls NLMPC::optimize, its default tolerances are
all disabled (Types.hpp:127-136) so it stops on maxeval or on a roundoff exception, and NLopt's own modifications to
Kraft's line search / stopping rules are not reproduced by SciPy.  What this oracle pins is the LOCAL OPTIMUM the
reference's formulation converges to under tight tolerances; the GPU solver is compared with it at solution level.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import minimize

FLT_INF = float(np.float32(np.inf))


def initial_guess(f, x0, u0, prev=None, slack=0.0, lb=None, ub=None):
    """NLOptimizer::run :431-510 (cold tile or previous solution, fixOptimalSolution, one-stage shift)."""
    nx, nu, ph, ch = f.nx, f.nu, f.ph, f.ch
    if prev is None:
        z = np.concatenate([np.tile(np.asarray(x0, float), ph), np.tile(np.asarray(u0, float), ch), [0.0]])
    else:
        z = np.array(prev, float)
    if lb is not None:
        bad = (z < lb) | (z > ub)
        with np.errstate(invalid="ignore"):
            z = np.where(bad, (ub - lb) / 2.0, z)          # NB: (ub-lb)/2, not the midpoint (NLOptimizer.hpp:713)
    out = z.copy()
    X = z[:ph * nx].reshape(ph, nx)
    out[:ph * nx] = np.vstack([X[1:], X[-1:]]).ravel()
    umv = (f.Iz2u @ z[ph * nx:ph * nx + nu * ch]).reshape(ph, nu)
    umv = np.vstack([umv[1:], umv[-1:]]).ravel()
    out[ph * nx:ph * nx + nu * ch] = f.Iu2z @ umv
    out[-1] = slack
    return out


def solve(f, x0, z0, lb, ub, maxiter=200, ftol=1e-12):
    x0 = np.asarray(x0, float)
    cons = [{"type": "eq", "fun": lambda z: f.state_eq(z, x0, want_jac=False)[0], "jac": lambda z: f.state_eq(z, x0)[1]}]
    if f.ineq is not None:     # NLopt convention c(z) <= 0; SciPy wants >= 0
        cons.append({"type": "ineq", "fun": lambda z: -f.ineq_con(z, x0)[0], "jac": lambda z: -f.ineq_con(z, x0)[1]})
    if getattr(f, "eq", None) is not None:     # user equality constraints (NLOptimizer::bindUserEq, NLOptimizer.hpp:314)
        cons.append({"type": "eq", "fun": lambda z: f.eq_con(z, x0)[0], "jac": lambda z: f.eq_con(z, x0)[1]})
    bounds = [(None if not np.isfinite(l) else l, None if not np.isfinite(u) else u) for l, u in zip(lb, ub)]
    res = minimize(lambda z: f.objective(z, x0, want_grad=False)[0], z0, jac=lambda z: f.objective(z, x0)[1],
                   method="SLSQP", bounds=bounds, constraints=cons, options=dict(maxiter=maxiter, ftol=ftol))
    X, U, e = f.unwrap(res.x, x0)
    return dict(z=res.x, cmd=U[0].copy(), cost=float(res.fun), nit=int(res.nit), success=bool(res.success), state=X, input=U, slack=e)


def default_bounds(f, hard_constraints=True):
    """lb/ub default to +-float infinity (NLOptimizer.hpp:70-73); hard constraints pin the slack to 0 (:160-164)."""
    lb = np.full(f.nz, -FLT_INF); ub = np.full(f.nz, FLT_INF)
    if hard_constraints:
        lb[-1] = ub[-1] = 0.0
    return lb, ub
