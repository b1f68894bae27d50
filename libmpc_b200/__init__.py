"""libmpc_b200 -- B200-native batched MPC solve engine behind the libmpc++ (mpc::LMPC<>) API.

Python host-side mirror of the reference's controller interface over the C ABI in include/b200mpc.h
(libmpc_b200/libb200mpc.so, hand-written sm_100a CUDA).  The C++20 mirror of the same interface is
include/mpc_b200/LMPC.hpp.  There is NO CPU fallback: constructing a controller without the CUDA extension or
without a GPU raises.

Reference anchors (paths relative to the libmpc++ repository):
  mpc::LMPC<Tnx,Tnu,Tndu,Tny,Tph,Tch>   include/mpc/LMPC.hpp:23-749
  mpc::IMPC::optimize                    include/mpc/IMPC.hpp:149-166
  mpc::LParameters / Result / OptSequence include/mpc/Types.hpp:142-199
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200MPC_LIB") or os.path.join(_HERE, "libb200mpc.so")      # env override: kernel build experiments

inf = float("inf")

# mpc::ResultStatus (Types.hpp:84-91)
SUCCESS, MAX_ITERATION, INFEASIBLE, ERROR, UNKNOWN = range(5)


class _Dims(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("nx", "nu", "ndu", "ny", "ph", "ch")]


class _Params(C.Structure):
    _fields_ = [("maximum_iteration", C.c_int), ("enable_warm_start", C.c_int), ("alpha", C.c_double),
                ("rho", C.c_double), ("eps_rel", C.c_double), ("eps_abs", C.c_double), ("eps_prim_inf", C.c_double),
                ("eps_dual_inf", C.c_double), ("adaptive_rho", C.c_int), ("polish", C.c_int), ("sigma", C.c_double),
                ("delta", C.c_double), ("adaptive_rho_tolerance", C.c_double), ("scaling", C.c_int),
                ("check_termination", C.c_int), ("adaptive_rho_interval", C.c_int), ("polish_refine_iter", C.c_int),
                ("time_limit", C.c_double)]


_lib = None


class _NLParams(C.Structure):
    _fields_ = [("max_sqp", C.c_int32), ("max_qp", C.c_int32), ("tol", C.c_double), ("ftol", C.c_double), ("qp_eps", C.c_double),
                ("rho", C.c_double)]


def load_library():
    """dlopen the CUDA extension; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(libmpc_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int32)
    H = C.c_void_p
    lib.b200mpc_last_error.restype = C.c_char_p
    lib.b200mpc_device_count.restype = C.c_int
    lib.b200mpc_lmpc_default_params.argtypes = [C.POINTER(_Params)]
    lib.b200mpc_lmpc_create.argtypes = [C.POINTER(_Dims), C.c_int, C.c_int, C.POINTER(H)]
    lib.b200mpc_lmpc_destroy.argtypes = [H]
    lib.b200mpc_lmpc_set_stream.argtypes = [H, C.c_void_p]
    lib.b200mpc_lmpc_set_params.argtypes = [H, C.POINTER(_Params)]
    for name, nptr in (("set_model", 3), ("set_disturbances", 2), ("set_weights", 3), ("set_state_bounds", 2),
                       ("set_input_bounds", 2), ("set_input_bounds_full", 2), ("set_output_bounds", 2), ("set_scalar_constraint", 4),
                       ("set_references", 3), ("set_exogenous_inputs", 1)):
        getattr(lib, "b200mpc_lmpc_" + name).argtypes = [H] + [C.c_void_p] * nptr + [C.c_int, C.c_int]
    lib.b200mpc_lmpc_set_warm_start.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int]
    lib.b200mpc_lmpc_get_warm_start.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int]
    lib.b200mpc_lmpc_solve.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int]
    lib.b200mpc_lmpc_closed_loop.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int]
    lib.b200mpc_lmpc_get_result.argtypes = [H] + [C.c_void_p] * 8 + [C.c_int]
    lib.b200mpc_lmpc_get_sequence.argtypes = [H] + [C.c_void_p] * 3 + [C.c_int]
    lib.b200mpc_lmpc_cmd_device_ptr.argtypes = [H, C.POINTER(C.c_void_p)]
    lib.b200mpc_lmpc_info.argtypes = [H, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_longlong)]
    lib.b200mpc_lmpc_set_launch.argtypes = [H, C.c_int, C.c_int]
    lib.b200mpc_lmpc_advance.argtypes = [H, C.c_void_p, C.c_void_p]
    lib.b200mpc_lmpc_set_engine.argtypes = [H, C.c_int, C.c_int]
    lib.b200mpc_lmpc_get_engine.argtypes = [H] + [C.POINTER(C.c_int)] * 3
    lib.b200mpc_lmpc_set_schedule.argtypes = [H, C.c_int]
    lib.b200mpc_lmpc_set_history_order.argtypes = [H, C.c_int]
    lib.b200mpc_lmpc_profile.argtypes = [H, C.c_void_p]
    lib.b200mpc_sync.argtypes = [H]
    lib.b200mpc_comm_unique_id.argtypes = [C.c_void_p]
    lib.b200mpc_comm_init_rank.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(H)]
    lib.b200mpc_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(H)]
    lib.b200mpc_comm_destroy.argtypes = [H]
    lib.b200mpc_comm_size.argtypes = [H, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.b200mpc_lmpc_allgather_cmd.argtypes = [H, H, C.c_void_p]
    lib.b200mpc_c2d.argtypes = [C.c_int] * 3 + [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.b200mpc_nlmpc_system_dims.argtypes = [C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 4
    lib.b200mpc_nlmpc_eval.argtypes = [C.c_int] * 4 + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p]
    lib.b200mpc_nlmpc_system_neq.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.b200mpc_nlmpc_system_ny.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.b200mpc_nlmpc_output.argtypes = [C.c_int] * 4 + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 2 + [C.c_int, C.c_void_p]
    lib.b200mpc_nlmpc_register_system.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int)]
    lib.b200mpc_nlmpc_compile_check.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_size_t)]
    lib.b200mpc_nlmpc_eval_ex.argtypes = ([C.c_int] * 4 + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] + [C.c_void_p] * 8 +
                                          [C.c_int, C.c_void_p])
    lib.b200mpc_nlmpc_solve_ex.argtypes = ([C.c_int] * 4 + [C.POINTER(_NLParams)] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] +
                                           [C.c_void_p] * 8 + [C.c_int, C.c_void_p])
    lib.b200mpc_nlmpc_default_params.argtypes = [C.POINTER(_NLParams)]
    lib.b200mpc_nlmpc_rk4.argtypes = [C.c_int] * 3 + [C.c_void_p] * 3 + [C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.b200mpc_nlmpc_closed_loop.argtypes = ([C.c_int] * 4 + [C.POINTER(_NLParams)] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 +
                                              [C.c_int] * 4 + [C.c_double] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p])
    lib.b200mpc_nlmpc_solve_smem_bytes.argtypes = [C.c_int] * 3
    lib.b200mpc_nlmpc_solve_smem_bytes.restype = C.c_longlong
    lib.b200mpc_nlmpc_solve.argtypes = [C.c_int] * 4 + [C.POINTER(_NLParams)] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 8 + [C.c_int, C.c_void_p]
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "b200mpc_last_error", "b200mpc_device_count", "b200mpc_lmpc_default_params", "b200mpc_lmpc_create",
    "b200mpc_lmpc_destroy", "b200mpc_lmpc_set_stream", "b200mpc_lmpc_set_params", "b200mpc_lmpc_set_model",
    "b200mpc_lmpc_set_disturbances", "b200mpc_lmpc_set_weights", "b200mpc_lmpc_set_state_bounds",
    "b200mpc_lmpc_set_input_bounds", "b200mpc_lmpc_set_output_bounds", "b200mpc_lmpc_set_scalar_constraint",
    "b200mpc_lmpc_set_references", "b200mpc_lmpc_set_exogenous_inputs", "b200mpc_lmpc_set_warm_start",
    "b200mpc_lmpc_get_warm_start", "b200mpc_lmpc_solve", "b200mpc_lmpc_closed_loop", "b200mpc_lmpc_get_result", "b200mpc_lmpc_get_sequence",
    "b200mpc_lmpc_cmd_device_ptr", "b200mpc_lmpc_info", "b200mpc_lmpc_set_launch", "b200mpc_lmpc_set_schedule", "b200mpc_lmpc_set_history_order", "b200mpc_lmpc_profile", "b200mpc_sync", "b200mpc_c2d", "b200mpc_nlmpc_system_dims", "b200mpc_nlmpc_eval",
    "b200mpc_nlmpc_default_params", "b200mpc_nlmpc_solve_smem_bytes", "b200mpc_nlmpc_solve",
    "b200mpc_nlmpc_system_neq", "b200mpc_nlmpc_register_system", "b200mpc_nlmpc_compile_check", "b200mpc_nlmpc_eval_ex",
    "b200mpc_nlmpc_solve_ex", "b200mpc_nlmpc_system_ny", "b200mpc_nlmpc_output", "b200mpc_lmpc_set_input_bounds_full", "b200mpc_lmpc_set_engine", "b200mpc_lmpc_get_engine", "b200mpc_lmpc_advance",
    "b200mpc_comm_unique_id", "b200mpc_comm_init_rank", "b200mpc_comm_init", "b200mpc_comm_destroy", "b200mpc_comm_size",
    "b200mpc_lmpc_allgather_cmd", "b200mpc_nlmpc_rk4", "b200mpc_nlmpc_closed_loop", "b200mpc_nlmpc_set_solver",
]


def _check(rc):
    if rc != 0:
        raise RuntimeError(f"b200mpc error {rc}: {load_library().b200mpc_last_error().decode()}")


@dataclass
class LParameters:
    """mpc::LParameters (Types.hpp:99-160)."""
    maximum_iteration: int = 100
    time_limit: float = 0.0
    enable_warm_start: bool = False
    alpha: float = 1.6
    rho: float = 1e-6
    eps_rel: float = 1e-4
    eps_abs: float = 1e-4
    eps_prim_inf: float = 1e-3
    eps_dual_inf: float = 1e-3
    verbose: bool = False
    adaptive_rho: bool = True
    polish: bool = True


@dataclass
class HorizonSlice:
    """mpc::HorizonSlice (Types.hpp:57-79): [start, end), all() == {-1,-1}."""
    start: int = -1
    end: int = -1

    @staticmethod
    def all():
        return HorizonSlice(-1, -1)


class Result:
    """mpc::Result<nu> for a batch (Types.hpp:168-182): arrays with a leading batch axis."""

    def __init__(self, cmd, cost, status, solver_status, is_feasible, iterations, rho_updates, status_polish):
        self.cmd, self.cost, self.status, self.solver_status = cmd, cost, status, solver_status
        self.is_feasible, self.iterations, self.rho_updates, self.status_polish = is_feasible, iterations, rho_updates, status_polish


class OptSequence:
    def __init__(self, state, input, output):
        self.state, self.input, self.output = state, input, output


def _slice(s):
    if s is None:
        return HorizonSlice.all()
    if isinstance(s, HorizonSlice):
        return s
    a, b = s
    return HorizonSlice(int(a), int(b))


class LMPC:
    """Batched mpc::LMPC<Tnx,Tnu,Tndu,Tny,Tph,Tch>.  With batch=1 it is a drop-in for one reference controller; every
    setter accepts either the reference's shape (shared by the batch) or the same with a leading batch axis."""

    def __init__(self, nx, nu, ndu, ny, ph, ch, batch=1, device=0):
        self.lib = load_library()
        self.nx, self.nu, self.ndu, self.ny, self.ph, self.ch, self.batch = nx, nu, ndu, ny, ph, ch, batch
        self.n = (ph + 1) * (nx + nu) + ph * nu
        self.m = 2 * (ph + 1) * (nx + nu) + (ph + 1) * ny + ph * nu + (ph + 1)
        self._h = C.c_void_p()
        dims = _Dims(nx, nu, ndu, ny, ph, ch)
        _check(self.lib.b200mpc_lmpc_create(C.byref(dims), batch, device, C.byref(self._h)))
        # host mirror of the builder state (API-level matrices, [.., dim, ph])
        z = np.zeros
        self._st = dict(
            OW=z((ny, ph)), UW=z((nu, ph)), DUW=z((nu, ph)),
            XMin=np.full((nx, ph), -inf), XMax=np.full((nx, ph), inf),
            YMin=np.full((ny, ph), -inf), YMax=np.full((ny, ph), inf),
            UMin=np.full((nu, ph), -inf), UMax=np.full((nu, ph), inf),   # internal minU/maxU: nu x ph
            SMin=np.full((ph,), -inf), SMax=np.full((ph,), inf), SX=z((nx,)), SU=z((nu,)),
            yRef=z((ny, ph)), uRef=z((nu, ph)), duRef=z((nu, ph)), uMeas=z((ndu, ph)))
        self._keep = []

    def __del__(self):
        try:
            if self._h:
                self.lib.b200mpc_lmpc_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ---- helpers ------------------------------------------------------------------------------
    def _hz(self, a, tail):
        """[.., dim, ph] -> contiguous stage-major [.., ph, dim]; returns (array, per_instance)."""
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == len(tail):
            pi = 0
        elif a.ndim == len(tail) + 1 and a.shape[0] == self.batch:
            pi = 1
        else:
            raise ValueError(f"expected shape {tail} or {(self.batch,) + tuple(tail)}, got {a.shape}")
        if tuple(a.shape[-len(tail):]) != tuple(tail):
            raise ValueError(f"expected trailing shape {tail}, got {a.shape}")
        if len(tail) == 2:
            a = np.swapaxes(a, -1, -2)
        return np.ascontiguousarray(a), pi

    def _mat(self, a, shape):
        a = np.asarray(a, dtype=np.float64)
        if a.shape == tuple(shape):
            return np.ascontiguousarray(a), 0
        if a.shape == (self.batch,) + tuple(shape):
            return np.ascontiguousarray(a), 1
        raise ValueError(f"expected shape {shape} or {(self.batch,) + tuple(shape)}, got {a.shape}")

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def _bcast_pi(self, arrs_pi):
        """The C setters take one per_instance flag for the whole group: broadcast shared members if needed."""
        pi = max(p for _, p in arrs_pi)
        out = []
        for a, p in arrs_pi:
            if pi and not p:
                a = np.ascontiguousarray(np.broadcast_to(a, (self.batch,) + a.shape))
            out.append(a)
        return out, pi

    # ---- unsupported in LMPC, as in the reference (LMPC.hpp:68-100) -----------------------------
    def setDiscretizationSamplingTime(self, ts):
        raise RuntimeError("Linear MPC supports only discrete time systems")

    def setInputScale(self, scaling):
        raise RuntimeError("Linear MPC does not support input scaling")

    def setStateScale(self, scaling):
        raise RuntimeError("Linear MPC does not support state scaling")

    # ---- setters ---------------------------------------------------------------------------------
    def setOptimizerParameters(self, p: LParameters):
        q = _Params()
        self.lib.b200mpc_lmpc_default_params(C.byref(q))
        q.maximum_iteration = int(p.maximum_iteration)
        q.enable_warm_start = int(bool(p.enable_warm_start))
        q.alpha, q.rho, q.eps_rel, q.eps_abs = p.alpha, p.rho, p.eps_rel, p.eps_abs
        q.eps_prim_inf, q.eps_dual_inf = p.eps_prim_inf, p.eps_dual_inf
        q.adaptive_rho, q.polish = int(bool(p.adaptive_rho)), int(bool(p.polish))
        q.time_limit = float(p.time_limit)                        # LOptimizer.hpp:256
        _check(self.lib.b200mpc_lmpc_set_params(self._h, C.byref(q)))

    def setStateSpaceModel(self, A, B, Cm):
        (A, B, Cm), pi = self._bcast_pi([self._mat(A, (self.nx, self.nx)), self._mat(B, (self.nx, self.nu)),
                                         self._mat(Cm, (self.ny, self.nx))])
        _check(self.lib.b200mpc_lmpc_set_model(self._h, self._p(A), self._p(B), self._p(Cm), pi, 0))
        _check(self.lib.b200mpc_sync(self._h))
        return True

    def setDisturbances(self, Bd, Dd):
        (Bd, Dd), pi = self._bcast_pi([self._mat(Bd, (self.nx, self.ndu)), self._mat(Dd, (self.ny, self.ndu))])
        _check(self.lib.b200mpc_lmpc_set_disturbances(self._h, self._p(Bd), self._p(Dd), pi, 0))
        _check(self.lib.b200mpc_sync(self._h))
        return True

    def _group(self, names, values, slice_, dims, horizon, validate_ctrl=False):
        """Shared implementation of the matrix / vector+slice setter pairs of LMPC.hpp."""
        st = self._st
        sl = _slice(slice_)
        vals = [np.asarray(v, dtype=np.float64) for v in values]
        is_matrix = vals[0].ndim >= 2 and vals[0].shape[-1] == horizon and vals[0].shape[-2] == dims[0] and slice_ is None
        if is_matrix:
            for nme, v in zip(names, vals):
                st[nme] = v.copy()
            return True
        # vector form
        lim = self.ch if validate_ctrl else self.ph
        if sl.start == -1 and sl.end == -1:
            rng = range(horizon)
        else:
            if sl.start >= sl.end or sl.start > lim or sl.end > lim:   # IMPC.hpp:252-272
                return False
            rng = range(sl.start, sl.end)
        for nme, v, dim in zip(names, vals, dims):
            cur = st[nme]
            if v.ndim == 2 and cur.ndim == 2:     # per-instance vectors: promote the host mirror
                cur = np.broadcast_to(cur, (self.batch,) + cur.shape).copy()
            if v.shape[-1] != dim:
                raise ValueError(f"{nme}: expected vector of length {dim}")
            for i in rng:
                cur[..., :, i] = v
            st[nme] = cur
        return True

    def setObjectiveWeights(self, OWeight, UWeight, DeltaUWeight, slice=None):
        ok = self._group(("OW", "UW", "DUW"), (OWeight, UWeight, DeltaUWeight), slice, (self.ny, self.nu, self.nu), self.ph)
        if ok:
            self._push_weights()
        return ok

    def setStateBounds(self, XMin, XMax, slice=None):
        ok = self._group(("XMin", "XMax"), (XMin, XMax), slice, (self.nx, self.nx), self.ph)
        if ok:
            self._push2("b200mpc_lmpc_set_state_bounds", "XMin", "XMax", (self.nx, self.ph))
        return ok

    def setOutputBounds(self, YMin, YMax, slice=None):
        ok = self._group(("YMin", "YMax"), (YMin, YMax), slice, (self.ny, self.ny), self.ph)
        if ok:
            self._push2("b200mpc_lmpc_set_output_bounds", "YMin", "YMax", (self.ny, self.ph))
        return ok

    def setInputBounds(self, UMin, UMax, slice=None):
        """Matrix form is [nu x ch] with the tail replicated (ProblemBuilder.hpp:397-413); vector+slice form indexes
        the control horizon (LMPC.hpp:197-241) and touches only the named columns (ProblemBuilder.hpp:469-477)."""
        a = np.asarray(UMin, dtype=np.float64)
        if slice is None and a.ndim >= 2 and a.shape[-1] == self.ch and a.shape[-2] == self.nu:
            for nme, v in (("UMin", UMin), ("UMax", UMax)):
                v = np.asarray(v, dtype=np.float64)
                full = np.concatenate([v, np.repeat(v[..., :, -1:], self.ph - self.ch, axis=-1)], axis=-1)
                self._st[nme] = full
        else:
            sl = _slice(slice)
            if sl.start == -1 and sl.end == -1:
                # LMPC.hpp:200-216 builds a [nu x ch] matrix and calls the matrix setter
                v0, v1 = np.asarray(UMin, float), np.asarray(UMax, float)
                return self.setInputBounds(np.repeat(v0[..., :, None], self.ch, axis=-1),
                                           np.repeat(v1[..., :, None], self.ch, axis=-1))
            if not self._group(("UMin", "UMax"), (UMin, UMax), sl, (self.nu, self.nu), self.ph, validate_ctrl=True):
                return False
        self._push_input_bounds()
        return True

    def setScalarConstraint(self, smin, smax, X, U, slice=None):
        """LMPC.hpp:355-422.  The multiplier [X;U] applies to every stage (ProblemBuilder.hpp:329-332)."""
        st = self._st
        sl = _slice(slice)
        if sl.start == -1 and sl.end == -1:
            for key, v in (("SMin", smin), ("SMax", smax)):
                v = np.asarray(v, dtype=np.float64)
                if v.ndim == 0:
                    v = np.full((self.ph,), float(v))
                elif v.shape not in ((self.ph,), (self.batch, self.ph)):
                    raise ValueError("scalar-constraint bounds: expected scalar, [ph] or [batch, ph]")
                st[key] = v.copy()
        else:
            if sl.start >= sl.end or sl.start > self.ph or sl.end > self.ph:
                return False
            for i in range(sl.start, sl.end):
                st["SMin"][..., i] = smin
                st["SMax"][..., i] = smax
        st["SX"], st["SU"] = np.asarray(X, float).copy(), np.asarray(U, float).copy()
        arrs = [self._hz(st["SMin"], (self.ph,)), self._hz(st["SMax"], (self.ph,)), self._hz(st["SX"], (self.nx,)),
                self._hz(st["SU"], (self.nu,))]
        (a, b, c, d), pi = self._bcast_pi(arrs)
        _check(self.lib.b200mpc_lmpc_set_scalar_constraint(self._h, self._p(a), self._p(b), self._p(c), self._p(d), pi, 0))
        _check(self.lib.b200mpc_sync(self._h))
        return True

    def setReferences(self, outRef, cmdRef, deltaCmdRef, slice=None):
        ok = self._group(("yRef", "uRef", "duRef"), (outRef, cmdRef, deltaCmdRef), slice, (self.ny, self.nu, self.nu), self.ph)
        if ok:
            arrs = [self._hz(self._st[k], s) for k, s in (("yRef", (self.ny, self.ph)), ("uRef", (self.nu, self.ph)),
                                                           ("duRef", (self.nu, self.ph)))]
            (a, b, c), pi = self._bcast_pi(arrs)
            _check(self.lib.b200mpc_lmpc_set_references(self._h, self._p(a), self._p(b), self._p(c), pi, 0))
            _check(self.lib.b200mpc_sync(self._h))
        return ok

    def setExogenousInputs(self, uMeas, slice=None):
        ok = self._group(("uMeas",), (uMeas,), slice, (self.ndu,), self.ph, validate_ctrl=slice is not None)
        if ok:
            a, pi = self._hz(self._st["uMeas"], (self.ndu, self.ph))
            _check(self.lib.b200mpc_lmpc_set_exogenous_inputs(self._h, self._p(a), pi, 0))
            _check(self.lib.b200mpc_sync(self._h))
        return ok

    def _push_weights(self):
        arrs = [self._hz(self._st[k], s) for k, s in (("OW", (self.ny, self.ph)), ("UW", (self.nu, self.ph)),
                                                       ("DUW", (self.nu, self.ph)))]
        (a, b, c), pi = self._bcast_pi(arrs)
        _check(self.lib.b200mpc_lmpc_set_weights(self._h, self._p(a), self._p(b), self._p(c), pi, 0))
        _check(self.lib.b200mpc_sync(self._h))

    def _push2(self, fn, k0, k1, shape):
        (a, b), pi = self._bcast_pi([self._hz(self._st[k0], shape), self._hz(self._st[k1], shape)])
        _check(getattr(self.lib, fn)(self._h, self._p(a), self._p(b), pi, 0))
        _check(self.lib.b200mpc_sync(self._h))

    def _push_input_bounds(self):
        # The host mirror holds the reference's internal [nu x ph] matrices.  The per-index setter writes ONE column and never
        # re-replicates the tail (ProblemBuilder.hpp:469-477), so after a slice set at ch-1 the tail columns ch..ph-1 keep what
        # the last matrix call put there: all ph columns are pushed verbatim.
        self._push2("b200mpc_lmpc_set_input_bounds_full", "UMin", "UMax", (self.nu, self.ph))

    # ---- warm start accessors (LMPC.hpp:677-722) -------------------------------------------------
    def getSolverWarmStartPrimal(self):
        x = np.empty((self.batch, self.n))
        _check(self.lib.b200mpc_lmpc_get_warm_start(self._h, self._p(x), None, 0))
        return x

    def getSolverWarmStartDual(self):
        y = np.empty((self.batch, self.m))
        _check(self.lib.b200mpc_lmpc_get_warm_start(self._h, None, self._p(y), 0))
        return y

    def setSolverWarmStart(self, primal, dual):
        x = np.ascontiguousarray(np.broadcast_to(np.asarray(primal, float), (self.batch, self.n)))
        y = np.ascontiguousarray(np.broadcast_to(np.asarray(dual, float), (self.batch, self.m)))
        _check(self.lib.b200mpc_lmpc_set_warm_start(self._h, self._p(x), self._p(y), 0))
        _check(self.lib.b200mpc_sync(self._h))

    # ---- solve -------------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        _check(self.lib.b200mpc_lmpc_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_launch(self, warps_per_cta=0, ctas_per_sm=0):
        _check(self.lib.b200mpc_lmpc_set_launch(self._h, warps_per_cta, ctas_per_sm))

    def set_engine(self, engine=0, cta_threads=0):
        """0 auto, 1 warp per controller (TMA-streamed state), 2 CTA per controller (state in shared memory); include/b200mpc.h."""
        _check(self.lib.b200mpc_lmpc_set_engine(self._h, int(engine), int(cta_threads)))

    def get_engine(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        _check(self.lib.b200mpc_lmpc_get_engine(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(engine=a.value, threads_per_cta=b.value, factor_in_shared_memory=bool(c.value))

    def set_schedule(self, gang=True):
        """Gang (default) or free scheduling of the persistent warps (include/b200mpc.h); results are identical."""
        _check(self.lib.b200mpc_lmpc_set_schedule(self._h, 1 if gang else 0))

    def set_history_order(self, enable=True):
        """Draw instances in the order of their previous solve's iteration counts, longest first (scheduling only)."""
        _check(self.lib.b200mpc_lmpc_set_history_order(self._h, 1 if enable else 0))

    def info(self):
        a, b, c = C.c_int(), C.c_size_t(), C.c_longlong()
        _check(self.lib.b200mpc_lmpc_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(warp_slots=a.value, workspace_bytes_per_slot=b.value, launches=c.value)

    def profile(self, fetch=False):
        """Enable (first call) or fetch the per-instance cycle counters [batch, 16]: columns 0..7 the phases [setup, factorize, admm sweeps, info,
        polish prep, polish factor, polish solve, unpack], 8..15 engine-specific sub-phase counters (engine 2: 8 / 9 = cycles in the
        forward / backward recurrence of all KKT solves)."""
        if not fetch:
            _check(self.lib.b200mpc_lmpc_profile(self._h, None))
            return None
        out = np.zeros((self.batch, 16), dtype=np.int64)
        _check(self.lib.b200mpc_lmpc_profile(self._h, self._p(out)))
        return out

    def solve_async(self, x0, u0, dev=False):
        """Enqueue one batched IOptimizer::run.  x0/u0: host arrays, or raw device pointers (ints) when dev=True."""
        if dev:
            _check(self.lib.b200mpc_lmpc_solve(self._h, C.c_void_p(int(x0)), C.c_void_p(int(u0)), 1))
            return
        x0 = np.ascontiguousarray(np.broadcast_to(np.asarray(x0, float), (self.batch, self.nx)))
        u0 = np.ascontiguousarray(np.broadcast_to(np.asarray(u0, float), (self.batch, self.nu)))
        self._keep = [x0, u0]
        _check(self.lib.b200mpc_lmpc_solve(self._h, self._p(x0), self._p(u0), 0))

    def advance(self, x_dev_ptr, u_out_dev_ptr):
        """x <- A x + B cmd on the device (raw device pointers), asynchronous: the plant step between two optimize() calls."""
        _check(self.lib.b200mpc_lmpc_advance(self._h, C.c_void_p(int(x_dev_ptr)), C.c_void_p(int(u_out_dev_ptr))))

    def fetch_result(self):
        B = self.batch
        cmd, cost = np.empty((B, self.nu)), np.empty(B)
        ints = [np.empty(B, dtype=np.int32) for _ in range(6)]
        _check(self.lib.b200mpc_lmpc_get_result(self._h, self._p(cmd), self._p(cost), *[self._p(a) for a in ints], 0))
        return Result(cmd, cost, ints[0], ints[1], ints[2].astype(bool), ints[3], ints[4], ints[5])

    def optimize(self, x0, lastU):
        """IMPC::optimize(x0,lastU) (IMPC.hpp:149-166) for the whole batch."""
        self.solve_async(x0, lastU)
        self._last = self.fetch_result()
        return self._last

    step = optimize   # pre-0.5.0 name of the same call (CHANGELOG.md:64-65)

    def closed_loop(self, x0, u0, steps, plant=None):
        """`steps` control steps on the device: optimize -> apply cmd -> x+ = Ap x + Bp u (the loop of the reference's
        examples).  plant = (Ap, Bp), shared [nx,nx],[nx,nu] or per instance [B,...]; None = the controller's model.
        Returns dict(x [steps+1,B,nx], u [steps,B,nu], status [steps,B], iterations [steps,B])."""
        B = self.batch
        x0 = np.ascontiguousarray(np.broadcast_to(np.asarray(x0, float), (B, self.nx)))
        u0 = np.ascontiguousarray(np.broadcast_to(np.asarray(u0, float), (B, self.nu)))
        Ap = Bp = None
        ppi = 0
        if plant is not None:
            Ap = np.ascontiguousarray(plant[0], dtype=np.float64); Bp = np.ascontiguousarray(plant[1], dtype=np.float64)
            ppi = 1 if Ap.ndim == 3 else 0
            if Ap.shape[-2:] != (self.nx, self.nx) or Bp.shape[-2:] != (self.nx, self.nu):
                raise ValueError("plant matrices have the wrong shape")
        out = dict(x=np.empty((steps + 1, B, self.nx)), u=np.empty((steps, B, self.nu)), status=np.empty((steps, B), np.int32),
                   iterations=np.empty((steps, B), np.int32))
        _check(self.lib.b200mpc_lmpc_closed_loop(self._h, self._p(x0), self._p(u0), int(steps), None if Ap is None else self._p(Ap),
                                                 None if Bp is None else self._p(Bp), ppi, self._p(out["x"]), self._p(out["u"]),
                                                 self._p(out["status"]), self._p(out["iterations"]), 0))
        return out

    def getLastResult(self):
        return self._last

    def getOptimalSequence(self):
        B, ph = self.batch, self.ph
        s, i, o = np.empty((B, ph + 1, self.nx)), np.empty((B, ph + 1, self.nu)), np.empty((B, ph + 1, self.ny))
        _check(self.lib.b200mpc_lmpc_get_sequence(self._h, self._p(s), self._p(i), self._p(o), 0))
        return OptSequence(s, i, o)

    def cmd_device_ptr(self):
        p = C.c_void_p()
        _check(self.lib.b200mpc_lmpc_cmd_device_ptr(self._h, C.byref(p)))
        return p.value

    def allgather_cmd(self, comm, cmd_all_dev_ptr):
        """The one exchange step of the sharded batch: cmd[batch,nu] of every rank -> cmd_all[nranks*batch,nu] (raw device
        pointer) on every rank; NCCL all-gather enqueued on this controller's stream right behind the solve."""
        _check(self.lib.b200mpc_lmpc_allgather_cmd(self._h, comm._c, C.c_void_p(int(cmd_all_dev_ptr))))

    def get_result_into(self, cmd_ptr=None, status_ptr=None, iters_ptr=None):
        """Device-to-device copy of results into caller-owned device buffers (async on the handle's stream)."""
        v = lambda p: C.c_void_p(int(p)) if p else None
        _check(self.lib.b200mpc_lmpc_get_result(self._h, v(cmd_ptr), None, v(status_ptr), None, None, v(iters_ptr), None, None, 1))


class Comm:
    """The multi-GPU exchange context of the engine (include/b200mpc.h): one NCCL communicator over the ranks that
    share a sharded batch.  `Comm.unique_id()` on rank 0 -> broadcast the 128 bytes by any means (torch.distributed, MPI, a file)
    -> `Comm(nranks, rank, uid, device)` on every rank.  The only collective of the path is LMPC.allgather_cmd."""

    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        _check(load_library().b200mpc_comm_unique_id(buf))
        return bytes(buf)

    def __init__(self, nranks, rank, uid, device=0):
        self.lib = load_library()
        self._c = C.c_void_p()
        if len(uid) != 128:
            raise ValueError("unique id must be 128 bytes")
        buf = (C.c_char * 128).from_buffer_copy(uid)
        _check(self.lib.b200mpc_comm_init_rank(int(nranks), int(rank), buf, int(device), C.byref(self._c)))
        self.nranks, self.rank = int(nranks), int(rank)

    def close(self):
        if getattr(self, "_c", None):
            self.lib.b200mpc_comm_destroy(self._c)
            self._c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def discretization(A, B, Ts):
    """mpc::discretization (include/mpc/Utils.hpp:23-47) on the device: A [nx,nx] or [batch,nx,nx], B likewise, Ts scalar or
    [batch].  Returns (Ad, Bd) with the batch axis of the inputs (none when everything is a single system)."""
    lib = load_library()
    A = np.ascontiguousarray(A, dtype=np.float64); B = np.ascontiguousarray(B, dtype=np.float64)
    Ts = np.ascontiguousarray(Ts, dtype=np.float64)
    nx, nu = A.shape[-1], B.shape[-1]
    mpi, tpi = A.ndim == 3, Ts.ndim == 1
    batch = A.shape[0] if mpi else (Ts.shape[0] if tpi else 1)
    if A.shape[-2:] != (nx, nx) or B.shape[-2:] != (nx, nu) or (mpi and B.shape[0] != batch) or (mpi and tpi and Ts.shape[0] != batch):
        raise ValueError("discretization: inconsistent shapes")
    Ad, Bd = np.empty((batch, nx, nx)), np.empty((batch, nx, nu))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    _check(lib.b200mpc_c2d(nx, nu, batch, vp(A), vp(B), int(mpi), vp(Ts.reshape(-1)), int(tpi), vp(Ad), vp(Bd), 0, None))
    return (Ad, Bd) if (mpi or tpi) else (Ad[0], Bd[0])


# ---- NLMPC problem evaluation ------------------------------------------------------------------------
SYS_VANDERPOL, SYS_OSCNET4, SYS_OSCNET6, SYS_UGV = 0, 1, 2, 3


def nlmpc_system_dims(system, ph):
    lib = load_library()
    v = [C.c_int() for _ in range(5)]
    _check(lib.b200mpc_nlmpc_system_dims(system, ph, *[C.byref(x) for x in v[:4]]))
    _check(lib.b200mpc_nlmpc_system_neq(system, ph, C.byref(v[4])))
    ny, ho = C.c_int(), C.c_int()
    _check(lib.b200mpc_nlmpc_system_ny(system, ph, C.byref(ny), C.byref(ho)))
    return dict(nx=v[0].value, nu=v[1].value, nparam=v[2].value, nineq=v[3].value, neq=v[4].value, ny=ny.value,
                has_output_map=bool(ho.value))


def register_system(cuda_source, type_name):
    """NLMPC::setStateSpaceFunction / setObjectiveFunction / setIneqConFunction / setEqConFunction for the batched engine:
    the model, cost and constraints as CUDA source (contract in include/b200mpc.h), compiled with NVRTC into the solver
    kernels.  Returns a system id usable wherever SYS_VANDERPOL etc. are."""
    lib = load_library()
    sid = C.c_int()
    _check(lib.b200mpc_nlmpc_register_system(cuda_source.encode(), type_name.encode(), C.byref(sid)))
    return sid.value


def compile_check(cuda_source, type_name, kernel=0):
    """NVRTC-compile a user system into one engine kernel (0 = evaluation, 1..4 = solve variants) without a GPU.
    Returns the cubin size; raises with the compiler log on error."""
    lib = load_library()
    n = C.c_size_t()
    _check(lib.b200mpc_nlmpc_compile_check(cuda_source.encode(), type_name.encode(), int(kernel), C.byref(n)))
    return n.value


class _NLScaling(C.Structure):
    _fields_ = [("state_scale", C.c_void_p), ("input_scale", C.c_void_p)]


def _scaling_arg(d, state_scale, input_scale):
    """-> (ctypes pointer or None, keep-alive tuple)."""
    if state_scale is None and input_scale is None:
        return None, ()
    keep = []
    def one(a, n):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape != (n,):
            raise ValueError("scaling vector has the wrong length")
        keep.append(a)
        return a.ctypes.data_as(C.c_void_p)
    sc = _NLScaling(one(state_scale, d["nx"]), one(input_scale, d["nu"]))
    keep.append(sc)
    return C.byref(sc), tuple(keep)


def nlmpc_eval(system, ph, ch, z, x0, params, want=("f", "grad", "ceq", "Jeq", "cin", "Jin"), state_scale=None, input_scale=None):
    """Batched Objective / Constraints evaluation with finite-difference derivatives on the GPU.
    z [B, nz], x0 [B, nx], params [nparam] (shared) or [B, nparam].  Returns a dict of numpy arrays ("cue"/"Jue": the
    user equality constraints of systems that define them)."""
    lib = load_library()
    d = nlmpc_system_dims(system, ph)
    z = np.ascontiguousarray(np.atleast_2d(z), dtype=np.float64)
    B, nz = z.shape
    if nz != ph * d["nx"] + ch * d["nu"] + 1:
        raise ValueError("z has the wrong length")
    x0 = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(x0), (B, d["nx"])), dtype=np.float64)
    params = np.ascontiguousarray(params, dtype=np.float64)
    ppi = 1 if params.ndim == 2 else 0
    if params.shape[-1] != d["nparam"]:
        raise ValueError(f"params: expected {d['nparam']} values")
    out = {}
    shapes = dict(f=(B,), grad=(B, nz), ceq=(B, ph * d["nx"]), Jeq=(B, ph * d["nx"], nz), cin=(B, d["nineq"]), Jin=(B, d["nineq"], nz),
                  cue=(B, d["neq"]), Jue=(B, d["neq"], nz))
    ptr = {}
    for k, shp in shapes.items():
        if k in want:
            out[k] = np.zeros(shp)
            ptr[k] = out[k].ctypes.data_as(C.c_void_p)
        else:
            ptr[k] = None
    sc, keep = _scaling_arg(d, state_scale, input_scale)
    _check(lib.b200mpc_nlmpc_eval_ex(system, ph, ch, B, z.ctypes.data_as(C.c_void_p), x0.ctypes.data_as(C.c_void_p),
                                     params.ctypes.data_as(C.c_void_p), ppi, sc, ptr["f"], ptr["grad"], ptr["ceq"], ptr["Jeq"],
                                     ptr["cin"], ptr["Jin"], ptr["cue"], ptr["Jue"], 0, None))
    del keep
    return out


def nlmpc_output(system, ph, ch, z, x0, params, state_scale=None, input_scale=None):
    """OptSequence::output of NLOptimizer::run (NLOptimizer.hpp:596-611): Model::getOutput (Model.hpp:72-96) of the unwrapped
    sequences of z on the GPU -> [B, ph+1, ny]; zeros when the system has no output map (as the reference)."""
    lib = load_library()
    d = nlmpc_system_dims(system, ph)
    z = np.ascontiguousarray(np.atleast_2d(z), dtype=np.float64)
    B, nz = z.shape
    if nz != ph * d["nx"] + ch * d["nu"] + 1:
        raise ValueError("z has the wrong length")
    x0 = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(x0), (B, d["nx"])), dtype=np.float64)
    params = np.ascontiguousarray(params, dtype=np.float64)
    ppi = 1 if params.ndim == 2 else 0
    y = np.zeros((B, ph + 1, d["ny"]))
    if d["ny"] == 0:
        return y
    sc, keep = _scaling_arg(d, state_scale, input_scale)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    _check(lib.b200mpc_nlmpc_output(system, ph, ch, B, vp(z), vp(x0), vp(params), ppi, sc, vp(y), 0, None))
    del keep
    return y


# ---- NLMPC solve -----------------------------------------------------------------------------------
FLT_INF = float(np.float32(np.inf))


@dataclass
class NLParameters:
    """mpc::NLParameters (Types.hpp:121-140).  The reference's four NLopt tolerances default to "disabled" (-1), which makes
    NLopt run to maximum_iteration; here positive xtol / ftol values become the SQP's step / |g'd| tolerances (defaults 1e-7 /
    1e-12, i.e. convergence to the finite-difference noise floor)."""
    maximum_iteration: int = 100
    time_limit: float = 0.0
    enable_warm_start: bool = False
    relative_ftol: float = -1.0
    relative_xtol: float = -1.0
    absolute_ftol: float = -1.0
    absolute_xtol: float = -1.0
    hard_constraints: bool = True
    verbose: bool = False


def nlmpc_solve(system, ph, ch, z0, x0, params, lb, ub, max_sqp=100, max_qp=200, tol=1e-7, ftol=1e-12, qp_eps=1e-5, rho=0.1,
                state_scale=None, input_scale=None):
    """Batched NLOptimizer::run core (NLOptimizer.hpp:519): z0 [B,nz] -> dict(z, cost, viol, status, iters, qp_iters)."""
    lib = load_library()
    d = nlmpc_system_dims(system, ph)
    z0 = np.ascontiguousarray(np.atleast_2d(z0), dtype=np.float64)
    B, nz = z0.shape
    if nz != ph * d["nx"] + ch * d["nu"] + 1:
        raise ValueError("z0 has the wrong length")
    x0 = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(x0), (B, d["nx"])), dtype=np.float64)
    params = np.ascontiguousarray(params, dtype=np.float64)
    ppi = 1 if params.ndim == 2 else 0
    if params.shape[-1] != d["nparam"] or (ppi and params.shape[0] != B):
        raise ValueError(f"params: expected [{d['nparam']}] or [batch, {d['nparam']}]")
    lb = np.ascontiguousarray(lb, dtype=np.float64); ub = np.ascontiguousarray(ub, dtype=np.float64)
    if lb.shape != (nz,) or ub.shape != (nz,):
        raise ValueError("lb/ub must have nz entries")
    q = _NLParams(int(max_sqp), int(max_qp), float(tol), float(ftol), float(qp_eps), float(rho))
    out = dict(z=np.zeros((B, nz)), cost=np.zeros(B), viol=np.zeros(B), status=np.zeros(B, np.int32), iters=np.zeros(B, np.int32),
               qp_iters=np.zeros(B, np.int32))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    sc, keep = _scaling_arg(d, state_scale, input_scale)
    _check(lib.b200mpc_nlmpc_solve_ex(system, ph, ch, B, C.byref(q), vp(z0), vp(x0), vp(params), ppi, sc, vp(lb), vp(ub), vp(out["z"]),
                                      vp(out["cost"]), vp(out["viol"]), vp(out["status"]), vp(out["iters"]), vp(out["qp_iters"]), 0, None))
    del keep
    return out


NL_SOLVER_AUTO, NL_SOLVER_DENSE, NL_SOLVER_STRUCTURED = 0, 1, 2


def nlmpc_set_solver(solver):
    """Select the NLMPC solve kernel (include/b200mpc.h: b200mpc_nlmpc_set_solver): automatic / dense / stage-structured."""
    _check(load_library().b200mpc_nlmpc_set_solver(int(solver)))


def nlmpc_rk4(system, x, u, params, h, integration_steps=1, stage=0):
    """mpc::RK4<N>::run (include/mpc/Integrator.hpp:38-56) for a batch of states on the GPU: the system's model with the input held,
    `integration_steps` classical Runge-Kutta steps of size h.  x [B,nx], u [B,nu] -> [B,nx]."""
    lib = load_library()
    d = nlmpc_system_dims(system, 1)
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
    B = x.shape[0]
    u = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(u), (B, d["nu"])), dtype=np.float64)
    params = np.ascontiguousarray(params, dtype=np.float64)
    ppi = 1 if params.ndim == 2 else 0
    if x.shape[1] != d["nx"] or params.shape[-1] != d["nparam"]:
        raise ValueError("nlmpc_rk4: wrong shapes")
    out = np.empty_like(x)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    _check(lib.b200mpc_nlmpc_rk4(system, B, int(stage), vp(x), vp(u), vp(params), ppi, float(h), int(integration_steps), vp(out), 0, None))
    return out


PLANT_DISCRETE, PLANT_EULER, PLANT_RK4 = 0, 1, 2


def nlmpc_closed_loop(system, ph, ch, x0, u0, params, lb, ub, steps, warm_start=True, plant_mode=PLANT_DISCRETE, plant_substeps=1,
                      plant_h=0.0, max_sqp=100, max_qp=200, tol=1e-7, ftol=1e-12, qp_eps=1e-5, rho=0.1, state_scale=None, input_scale=None):
    """`steps` control steps of the NLMPC examples' loop on the device (include/b200mpc.h: b200mpc_nlmpc_closed_loop): guess /
    repair / shift (NLOptimizer.hpp:425-510) -> solve -> apply cmd -> plant step, no host round trip.
    Returns dict(x [steps+1,B,nx], u [steps,B,nu], status, iterations [steps,B], cost [steps,B])."""
    lib = load_library()
    d = nlmpc_system_dims(system, ph)
    x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
    B = x0.shape[0]
    u0 = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(u0), (B, d["nu"])), dtype=np.float64)
    params = np.ascontiguousarray(params, dtype=np.float64)
    ppi = 1 if params.ndim == 2 else 0
    nz = ph * d["nx"] + ch * d["nu"] + 1
    lb = np.ascontiguousarray(lb, dtype=np.float64); ub = np.ascontiguousarray(ub, dtype=np.float64)
    if lb.shape != (nz,) or ub.shape != (nz,) or x0.shape[1] != d["nx"]:
        raise ValueError("nlmpc_closed_loop: wrong shapes")
    q = _NLParams(int(max_sqp), int(max_qp), float(tol), float(ftol), float(qp_eps), float(rho))
    out = dict(x=np.empty((steps + 1, B, d["nx"])), u=np.empty((steps, B, d["nu"])), status=np.empty((steps, B), np.int32),
               iterations=np.empty((steps, B), np.int32), cost=np.empty((steps, B)))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    sc, keep = _scaling_arg(d, state_scale, input_scale)
    _check(lib.b200mpc_nlmpc_closed_loop(system, ph, ch, B, C.byref(q), vp(x0), vp(u0), vp(params), ppi, sc, vp(lb), vp(ub), int(steps),
                                         int(bool(warm_start)), int(plant_mode), int(plant_substeps), float(plant_h), vp(out["x"]),
                                         vp(out["u"]), vp(out["status"]), vp(out["iterations"]), vp(out["cost"]), 0, None))
    del keep
    return out


class NLMPC:
    """Batched mpc::NLMPC<Tnx,Tnu,Tny,Tph,Tch,Tineq,Teq> (include/mpc/NLMPC.hpp) for the built-in device systems.

    The reference takes the model / objective / constraints as std::function callbacks (NLMPC.hpp:139-281), which cannot
    run on the device; here `system` selects a device functor (SYS_VANDERPOL, SYS_UGV, ..., or the id `register_system`
    returns for the user's own CUDA source) and `setSystemParameters` supplies its numbers (shared or per controller).
    Bounds, scaling, parameters, warm start and results follow the reference."""

    def __init__(self, system, ph, ch, batch=1):
        self.lib = load_library()
        d = nlmpc_system_dims(system, ph)
        self.system, self.ph, self.ch, self.batch = system, ph, ch, batch
        self.nx, self.nu, self.nparam, self.nineq, self.neq = d["nx"], d["nu"], d["nparam"], d["nineq"], d["neq"]
        self.state_scale = self.input_scale = None
        self.eq_tolerance = 1e-10                                                      # NLMPC.hpp:262
        self.nz = ph * self.nx + ch * self.nu + 1
        self.lb = np.full(self.nz, -FLT_INF); self.ub = np.full(self.nz, FLT_INF)     # NLOptimizer.hpp:69-73
        self.params = None
        self.p = NLParameters()
        self.ineq_tolerance = 1e-10                                                    # NLMPC.hpp:229
        self._apply_slack_bound()
        self.opt_vector = np.zeros((batch, self.nz))
        self.current_slack = np.zeros(batch)
        self.is_first_iteration = True
        self.sequence = None
        self.result = None

    def _apply_slack_bound(self):
        # NLOptimizer::setParameters :160-190: hard constraints pin the slack to zero, otherwise it is free in [0, inf)
        if self.p.hard_constraints:
            self.lb[-1] = self.ub[-1] = 0.0
        else:
            self.lb[-1] = 0.0; self.ub[-1] = FLT_INF

    def setOptimizerParameters(self, p: NLParameters):
        self.p = p
        self._apply_slack_bound()

    def setSystemParameters(self, params):
        params = np.ascontiguousarray(params, dtype=np.float64)
        if params.shape not in ((self.nparam,), (self.batch, self.nparam)):
            raise ValueError(f"params: expected [{self.nparam}] or [{self.batch}, {self.nparam}]")
        self.params = params
        return True

    def setIneqTolerance(self, tol):
        self.ineq_tolerance = float(tol)

    def setEqTolerance(self, tol):
        self.eq_tolerance = float(tol)

    def setStateScale(self, scaling):
        """NLMPC::setStateScale (NLMPC.hpp:122-130)."""
        self.state_scale = np.ascontiguousarray(scaling, dtype=np.float64).reshape(self.nx)

    def setInputScale(self, scaling):
        """NLMPC::setInputScale (NLMPC.hpp:108-116)."""
        self.input_scale = np.ascontiguousarray(scaling, dtype=np.float64).reshape(self.nu)

    def _bounds(self, lo, hi, dim, horizon, offset, slice_):
        lo = np.asarray(lo, float); hi = np.asarray(hi, float)
        if lo.ndim == 2:                                   # mat<dim, horizon>  (NLMPC.hpp:292-330)
            if lo.shape != (dim, horizon) or hi.shape != (dim, horizon):
                raise ValueError("bounds matrix has the wrong shape")
            rng = range(horizon); col = lambda a, i: a[:, i]
        else:                                              # vector + HorizonSlice (NLMPC.hpp:362-400)
            if lo.shape != (dim,) or hi.shape != (dim,):
                raise ValueError("bounds vector has the wrong shape")
            s = _slice(slice_)
            if s.start == -1 and s.end == -1:
                rng = range(horizon)
            else:
                a = 0 if s.start == -1 else s.start
                b = horizon if s.end == -1 else s.end
                if not (0 <= a < b <= horizon):
                    return False
                rng = range(a, b)
            col = lambda a, i: a
        for i in rng:
            self.lb[offset + i * dim: offset + (i + 1) * dim] = col(lo, i)
            self.ub[offset + i * dim: offset + (i + 1) * dim] = col(hi, i)
        return True

    def setStateBounds(self, XMin, XMax, slice=None):
        return self._bounds(XMin, XMax, self.nx, self.ph, 0, slice)

    def setInputBounds(self, UMin, UMax, slice=None):
        return self._bounds(UMin, UMax, self.nu, self.ch, self.ph * self.nx, slice)

    def setOutputBounds(self, YMin, YMax, slice=None):
        return False                                       # the reference ignores them too (NLMPC.hpp:342-349,410-417)

    def _initial_guess(self, x0, u0):
        """NLOptimizer::run :431-510 -- cold tile or previous optimum, fixOptimalSolution (:705-716), one-stage shift."""
        B, ph, ch, nx, nu = self.batch, self.ph, self.ch, self.nx, self.nu
        z = self.opt_vector
        if self.is_first_iteration or not self.p.enable_warm_start:
            z[:, :ph * nx] = np.tile(x0, (1, ph))
            z[:, ph * nx:ph * nx + ch * nu] = np.tile(u0, (1, ch))
        bad = (z < self.lb) | (z > self.ub)
        with np.errstate(invalid="ignore"):
            z = np.where(bad, (self.ub - self.lb) / 2.0, z)
        out = z.copy()
        X = z[:, :ph * nx].reshape(B, ph, nx)
        out[:, :ph * nx] = np.concatenate([X[:, 1:], X[:, -1:]], axis=1).reshape(B, -1)
        Uz = z[:, ph * nx:ph * nx + ch * nu].reshape(B, ch, nu)
        blk = np.minimum(np.arange(ph), ch - 1)            # Iz2u: stage -> control block (Mapping.hpp:100-140)
        Umv = Uz[:, blk]                                   # [B, ph, nu]
        Umv = np.concatenate([Umv[:, 1:], Umv[:, -1:]], axis=1)
        res = Umv[:, :ch]                                  # Iu2z picks the first stage of every block (Mapping.hpp:245-256)
        out[:, ph * nx:ph * nx + ch * nu] = res.reshape(B, -1)
        out[:, -1] = self.current_slack
        return out

    def optimize(self, x0, lastU):
        if self.params is None:
            raise RuntimeError("setSystemParameters has not been called")
        B = self.batch
        x0 = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(np.asarray(x0, float)), (B, self.nx)))
        u0 = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(np.asarray(lastU, float)), (B, self.nu)))
        z0 = self._initial_guess(x0, u0)
        xt = [t for t in (self.p.relative_xtol, self.p.absolute_xtol) if t > 0]
        ft = [t for t in (self.p.relative_ftol, self.p.absolute_ftol) if t > 0]
        r = nlmpc_solve(self.system, self.ph, self.ch, z0, x0, self.params, self.lb, self.ub, max_sqp=self.p.maximum_iteration,
                        tol=min(xt) if xt else 1e-7, ftol=min(ft) if ft else 1e-12, state_scale=self.state_scale,
                        input_scale=self.input_scale)
        z = r["z"]
        self.opt_vector = z.copy()
        self.is_first_iteration = False
        self.current_slack = z[:, -1].copy()
        ph, ch, nx, nu = self.ph, self.ch, self.nx, self.nu
        X = np.concatenate([x0[:, None, :], z[:, :ph * nx].reshape(B, ph, nx)], axis=1)
        blk = np.minimum(np.minimum(np.arange(ph + 1), ph - 1), ch - 1)
        U = z[:, ph * nx:ph * nx + ch * nu].reshape(B, ch, nu)[:, blk]
        if self.state_scale is not None:
            X = X / self.state_scale                              # Mapping::unwrapVector (row 0 = x0 included)
        if self.input_scale is not None:
            U = U * self.input_scale
        feas = np.ones(B, bool)
        if self.nineq or self.neq:
            ev = nlmpc_eval(self.system, ph, ch, z, x0, self.params, want=("cin", "cue"), state_scale=self.state_scale,
                            input_scale=self.input_scale)
            if self.nineq:
                feas &= ~(ev["cin"] > self.ineq_tolerance).any(axis=1)       # Constraints::isFeasible (Constraints.hpp:157-201)
            if self.neq:
                feas &= ~(np.abs(ev["cue"]) > self.eq_tolerance).any(axis=1)
        status = np.where(r["status"] == 0, 0, 1).astype(np.int32)            # SUCCESS / MAX_ITERATION (Types.hpp:87-94)
        solver_status = np.where(r["status"] == 0, 4, 5).astype(np.int32)     # nlopt::XTOL_REACHED / MAXEVAL_REACHED
        self.result = Result(U[:, 0].copy(), r["cost"], status, solver_status, feas, r["iters"], r["qp_iters"], np.zeros(B, np.int32))
        self.result.viol = r["viol"]
        Y = nlmpc_output(self.system, ph, ch, z, x0, self.params, state_scale=self.state_scale, input_scale=self.input_scale)
        self.sequence = OptSequence(X, U, Y)                      # sequence.output = model->getOutput(Xmat, Umat) (NLOptimizer.hpp:611)
        return self.result

    step = optimize

    def getLastResult(self):
        return self.result

    def getOptimalSequence(self):
        return self.sequence
