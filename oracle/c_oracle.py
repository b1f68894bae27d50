"""ctypes front-end of oracle/lmpc_oracle.c (CPU oracle: TEST INFRASTRUCTURE / CPU BASELINE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this."""
import ctypes as C
import os
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liblmpc_oracle.so")


class Dims(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("nx", "nu", "ndu", "ny", "ph", "ch")]


class Params(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("max_iter", "adaptive_rho", "polish", "scaling", "check_termination",
                                       "adaptive_rho_interval", "polish_refine_iter", "warm_start")] + \
               [(k, C.c_double) for k in ("alpha", "rho", "sigma", "delta", "eps_abs", "eps_rel", "eps_prim_inf",
                                          "eps_dual_inf", "adaptive_rho_tolerance")]


_PROB_FIELDS = ("A", "B", "C", "Bd", "Dd", "OW", "UW", "DUW", "XMin", "XMax", "YMin", "YMax", "UMin", "UMax", "SMin",
                "SMax", "SX", "SU", "yRef", "uRef", "duRef", "uMeas")


class Prob(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in _PROB_FIELDS]


class Result(C.Structure):
    _fields_ = [("cost", C.c_double)] + [(k, C.c_int) for k in ("status", "solver_status", "is_feasible", "iters",
                                                                   "rho_updates", "status_polish")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise RuntimeError("oracle/liblmpc_oracle.so missing: run `make -C oracle` (or __graft_entry__.build())")
        _lib = C.CDLL(_LIB)
        _lib.lmpc_oracle_solve_batch.restype = C.c_int
        _lib.lmpc_oracle_max_threads.restype = C.c_int
    return _lib


def default_params(**kw):
    p = Params(max_iter=100, adaptive_rho=1, polish=1, scaling=10, check_termination=25, adaptive_rho_interval=25,
               polish_refine_iter=3, warm_start=0, alpha=1.6, rho=1e-6, sigma=1e-6, delta=1e-6, eps_abs=1e-4,
               eps_rel=1e-4, eps_prim_inf=1e-3, eps_dual_inf=1e-3, adaptive_rho_tolerance=5.0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def prob_from_formulation(f, yref_batch=None):
    """Build the C problem description from an oracle.lmpc_formulation.LMPCFormulation (API-level arrays)."""
    nx, nu, ndu, ny, ph = f.nx, f.nu, f.ndu, f.ny, f.ph
    a = {}
    a["A"] = f.ssA[:nx, :nx]; a["B"] = f.ssB[:nx, :]; a["C"] = f.ssC[:ny, :nx]
    a["Bd"] = f.ssBv[:nx, :]; a["Dd"] = f.ssDv[:ny, :]
    a["OW"] = f.wOutput[:, 1:].T; a["UW"] = f.wU[:, 1:].T; a["DUW"] = f.wDeltaU.T
    a["XMin"] = f.minX[:, 1:].T; a["XMax"] = f.maxX[:, 1:].T; a["YMin"] = f.minY[:, 1:].T; a["YMax"] = f.maxY[:, 1:].T
    a["UMin"] = f.minU.T; a["UMax"] = f.maxU.T
    a["SMin"] = f.sMin[1:]; a["SMax"] = f.sMax[1:]
    a["SX"] = f.sMultiplier[0, :nx]; a["SU"] = f.sMultiplier[0, nx:nx + nu]
    a["yRef"] = f.yRef.T if yref_batch is None else yref_batch
    a["uRef"] = f.uRef.T; a["duRef"] = f.duRef.T; a["uMeas"] = f.uMeas.T
    keep = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in a.items()}
    for k, v in keep.items():
        if v.size == 0:
            keep[k] = np.zeros(1)
    pr = Prob(**{k: keep[k].ctypes.data_as(C.c_void_p) for k in _PROB_FIELDS})
    return pr, keep


def solve_batch(f, x0, u0, params=None, yref_batch=None, warm=None, nthreads=1, want_xy=True):
    """x0 [B,nx], u0 [B,nu]; yref_batch optional [B, ph, ny] (stage-major).  Returns dict of arrays."""
    L = lib()
    params = params or default_params()
    x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
    B = x0.shape[0]
    u0 = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(u0), (B, f.nu)), dtype=np.float64)
    d = Dims(f.nx, f.nu, f.ndu, f.ny, f.ph, f.ch)
    if yref_batch is not None:
        yref_batch = np.ascontiguousarray(yref_batch, dtype=np.float64)
    pr, keep = prob_from_formulation(f, yref_batch)
    cmd = np.zeros((B, f.nu))
    res = (Result * B)()
    sx = np.zeros((B, f.n)) if want_xy else None
    sy = np.zeros((B, f.m)) if want_xy else None
    wx = wy = None
    if warm is not None:
        wx = np.ascontiguousarray(warm[0], dtype=np.float64); wy = np.ascontiguousarray(warm[1], dtype=np.float64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    t = time.perf_counter()
    L.lmpc_oracle_solve_batch(C.byref(d), C.byref(params), C.byref(pr), B, vp(x0), vp(u0), 1 if yref_batch is not None else 0,
                              vp(wx), vp(wy), vp(cmd), res, vp(sx), vp(sy), int(nthreads))
    dt = time.perf_counter() - t
    out = dict(cmd=cmd, x=sx, y=sy, seconds=dt)
    for k in ("cost", "status", "solver_status", "is_feasible", "iters", "rho_updates", "status_polish"):
        out[k] = np.array([getattr(r, k) for r in res])
    return out


def time_batch(ph, x0, r, max_iter, cores):
    """bench.py cpu_baseline: quadrotor workload, `cores` OpenMP threads, returns the cpu_baseline dict."""
    from oracle.lmpc_formulation import quadrotor_formulation
    f = quadrotor_formulation(ph)
    B = x0.shape[0]
    yref = np.zeros((B, ph, f.ny))
    yref[:, :, 2] = np.asarray(r)[:, None]
    out = solve_batch(f, x0, np.zeros((B, f.nu)), default_params(max_iter=max_iter), yref_batch=yref, nthreads=cores, want_xy=False)
    return {"value": B / out["seconds"], "unit": "solves/s", "cores": int(cores), "kind": "port",
            "sample": f"{B} solves of the same workload through oracle/lmpc_oracle.c (gcc -O3 -march=native, {cores} pthread(s)): "
                      "dense P/A resident, per-step q/l/u + dense->CSC scan + Ruiz scaling + sparse LDL' + ADMM + polish",
            "iters_mean": float(out["iters"].mean())}
