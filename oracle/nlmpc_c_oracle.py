"""ctypes front-end of oracle/nlmpc_oracle.c + a compiled-callback SLSQP solve (CPU oracle: TEST INFRASTRUCTURE / CPU
BASELINE ONLY; only tests/ and the CPU-baseline tooling may import this).

`solve` is the reference's NLMPC solve path as far as it can be rebuilt here: SciPy's SLSQP core (the same Kraft algorithm
NLopt's LD_SLSQP translates; NLopt itself is absent, see oracle/nlmpc_slsqp.py) driving the restated Objective /
Constraints evaluated in C, finite differences included -- i.e. the work the reference does per step without the Python
callback overhead of oracle/nlmpc_slsqp.py.  PARITY UNPINNED upstream (no reference test calls NLMPC::optimize)."""
import ctypes as C
import os
import time

import numpy as np
from scipy.optimize import minimize

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libnlmpc_oracle.so")
_lib = None
_DIMS = {0: (2, 1, lambda ph: ph + 1), 1: (8, 4, lambda ph: (ph + 1) * 4), 2: (12, 6, lambda ph: (ph + 1) * 6), 3: (4, 2, lambda ph: (ph + 1) * 2)}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise RuntimeError(f"{_LIB} is missing: run `make -C oracle`")
        _lib = C.CDLL(_LIB)
        _lib.nlmpc_oracle_eval.argtypes = [C.c_int] * 3 + [C.c_void_p] * 9
    return _lib


def evaluate(system, ph, ch, z, x0, params, want=("f", "grad", "ceq", "Jeq", "cin", "Jin")):
    nx, nu, nif = _DIMS[system]
    nz, ni = ph * nx + ch * nu + 1, nif(ph)
    z = np.ascontiguousarray(z, dtype=np.float64); x0 = np.ascontiguousarray(x0, dtype=np.float64)
    params = np.ascontiguousarray(params, dtype=np.float64)
    shapes = dict(f=(1,), grad=(nz,), ceq=(ph * nx,), Jeq=(ph * nx, nz), cin=(ni,), Jin=(ni, nz))
    out = {k: np.empty(s) for k, s in shapes.items() if k in want}
    ptr = lambda k: out[k].ctypes.data_as(C.c_void_p) if k in out else None
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    if lib().nlmpc_oracle_eval(system, ph, ch, vp(z), vp(x0), vp(params), ptr("f"), ptr("grad"), ptr("ceq"), ptr("Jeq"), ptr("cin"),
                               ptr("Jin")):
        raise RuntimeError("nlmpc_oracle_eval failed")
    if "f" in out:
        out["f"] = float(out["f"][0])
    return out


def solve(system, ph, ch, x0, z0, params, lb, ub, maxiter=200, ftol=1e-12):
    """NLOptimizer::run's optimize call (NLOptimizer.hpp:519) on the C-evaluated problem."""
    ev = lambda z, w: evaluate(system, ph, ch, z, x0, params, want=w)
    cons = [{"type": "eq", "fun": lambda z: ev(z, ("ceq",))["ceq"], "jac": lambda z: ev(z, ("ceq", "Jeq"))["Jeq"]},
            {"type": "ineq", "fun": lambda z: -ev(z, ("cin",))["cin"], "jac": lambda z: -ev(z, ("cin", "Jin"))["Jin"]}]
    bounds = [(None if not np.isfinite(l) else l, None if not np.isfinite(u) else u) for l, u in zip(lb, ub)]
    res = minimize(lambda z: ev(z, ("f",))["f"], z0, jac=lambda z: ev(z, ("f", "grad"))["grad"], method="SLSQP", bounds=bounds,
                   constraints=cons, options=dict(maxiter=maxiter, ftol=ftol))
    nx, nu, _ = _DIMS[system]
    return dict(z=res.x, cmd=res.x[ph * nx:ph * nx + nu].copy(), cost=float(res.fun), nit=int(res.nit), success=bool(res.success))


def _worker(args):
    system, ph, ch, x0s, z0s, params, lb, ub = args
    t = time.perf_counter()
    ok = 0
    for x0, z0 in zip(x0s, z0s):
        ok += int(solve(system, ph, ch, x0, z0, params, lb, ub)["success"])
    return time.perf_counter() - t, ok


def time_batch(system, ph, ch, x0, z0, params, lb, ub, cores=1):
    """solves/s of `len(x0)` solves spread over `cores` processes (one solve stream per core, as the reference would run)."""
    import multiprocessing as mp
    chunks = [(system, ph, ch, x0[k::cores], z0[k::cores], params, lb, ub) for k in range(cores)]
    t = time.perf_counter()
    if cores > 1:
        with mp.get_context("fork").Pool(cores) as pool:
            r = pool.map(_worker, chunks)
    else:
        r = [_worker(chunks[0])]
    dt = time.perf_counter() - t
    return dict(solves_per_s=len(x0) / dt, seconds=dt, cores=cores, converged=sum(k for _, k in r) / len(x0))
