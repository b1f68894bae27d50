"""`pympcxx`-compatible front end over the B200 engine.

The reference ships a pybind11 module `pympcxx` (python/pybind_export.cpp:13-213) exposing ONE controller per object:
`LMPC(nx, nu, ndu, ny, ph, ch)`, `NLMPC(nx, nu, ny, ph, ch, ineq, eq)`, `LParameters`, `NLParameters`, `HorizonSlice`,
`Result`, `OptSequence`, `SolutionStats`, `ResultStatus`, `LoggerLevel`.  This module has the same names, constructor
arguments, method names and single-controller shapes (`res.cmd` is a vector of length nu, `seq.state` is [(ph+1), nx]), so
`import libmpc_b200.pympcxx as mpc` runs python/examples/example.py unchanged.  Each object is a batch == 1 handle of the
batched engine; use `libmpc_b200.LMPC(..., batch=B)` directly for batches.

The one thing that cannot carry over is the Python callbacks of NLMPC (`setStateSpaceFunction(f)` etc.,
pybind_export.cpp:80-84): a Python function cannot run inside a CUDA kernel.  The callback setters raise with that
explanation; `NLMPC.setSystemSource(cuda_source, type_name, params)` takes the same model / cost / constraints as CUDA
source instead (contract in include/b200mpc.h), `NLMPC.setSystem(system_id, params)` selects a built-in example system.
"""
from __future__ import annotations

import datetime
import enum
import statistics
import time

import numpy as np

import libmpc_b200 as _L
from libmpc_b200 import HorizonSlice, LParameters, NLParameters      # noqa: F401  (same fields as the reference's)

__all__ = ["LMPC", "NLMPC", "Parameters", "LParameters", "NLParameters", "HorizonSlice", "Result", "OptSequence", "SolutionStats",
           "ResultStatus", "LoggerLevel", "SUCCESS", "MAX_ITERATION", "INFEASIBLE", "ERROR", "UNKNOWN"]


class ResultStatus(enum.IntEnum):
    """mpc::ResultStatus (Types.hpp:84-91)."""
    SUCCESS = 0
    MAX_ITERATION = 1
    INFEASIBLE = 2
    ERROR = 3
    UNKNOWN = 4


SUCCESS, MAX_ITERATION, INFEASIBLE, ERROR, UNKNOWN = ResultStatus     # py::enum_::export_values()


class LoggerLevel(enum.Enum):
    """mpc::Logger::LogLevel (Logger.hpp:30-36)."""
    DEEP = 0
    NORMAL = 1
    ALERT = 2
    NONE = 3


class Parameters:
    """mpc::Parameters (Types.hpp:99-119): the common base of LParameters / NLParameters."""
    maximum_iteration = 100
    time_limit = 0.0
    enable_warm_start = False


_OSQP_MSG = {1: "solved", 2: "solved inaccurate", 3: "primal infeasible inaccurate", 4: "dual infeasible inaccurate",
             -2: "maximum iterations reached", -3: "primal infeasible", -4: "dual infeasible", -7: "problem non convex",
             -10: "unsolved", -1: "setup error"}


class Result:
    """mpc::Result<nu> of one controller (Types.hpp:168-182, pybind_export.cpp:173-178)."""

    def __init__(self, solver_status, solver_status_msg, cost, cmd, status, is_feasible=True):
        self.solver_status, self.solver_status_msg, self.cost, self.cmd = int(solver_status), solver_status_msg, float(cost), cmd
        self.status, self.is_feasible = ResultStatus(int(status)), bool(is_feasible)


class OptSequence:
    """mpc::OptSequence (Types.hpp:184-199): state [(ph+1), nx], input [(ph+1), nu], output [(ph+1), ny]."""

    def __init__(self, state, input, output):
        self.state, self.input, self.output = state, input, output


class SolutionStats:
    """mpc::SolutionStats (Types.hpp:201-219), filled by the host-side timer around optimize() (Profiler.hpp:30-110)."""

    def __init__(self, times, states):
        td = lambda s: datetime.timedelta(seconds=float(s))
        self.numberOfSolutions = len(times)
        self.totalSolutionTime = td(sum(times))
        self.minSolutionTime = td(min(times)) if times else td(0)
        self.maxSolutionTime = td(max(times)) if times else td(0)
        self.averageSolutionTime = td(sum(times) / len(times)) if times else td(0)
        self.standardDeviation = td(statistics.pstdev(times)) if len(times) > 1 else td(0)
        self.solutionsStates = dict(states)


class _Common:
    def _init_common(self):
        self._times, self._states = [], {}
        self._log_level, self._log_prefix = LoggerLevel.NORMAL, ""

    def setLoggerLevel(self, level):
        self._log_level = level

    def setLoggerPrefix(self, prefix):
        self._log_prefix = str(prefix)

    def getExecutionStats(self):
        return SolutionStats(self._times, self._states)

    def resetStats(self):
        self._times, self._states = [], {}

    def _timed(self, fn):
        t = time.perf_counter()
        r = fn()
        self._times.append(time.perf_counter() - t)
        return r

    def _record(self, res):
        self._states[res.status] = self._states.get(res.status, 0) + 1
        if self._log_level == LoggerLevel.DEEP:
            print(f"[MPC++{' ' + self._log_prefix if self._log_prefix else ''}] Optimization step: status={res.status.name} cost={res.cost}")
        return res

    def getOptimalSequence(self):
        s = self._c.getOptimalSequence()
        return OptSequence(s.state[0], s.input[0], s.output[0])

    def getLastResult(self):
        return self._last


class LMPC(_Common):
    """mpc::LMPC<Dynamic...> as exported at pybind_export.cpp:93-123."""

    def __init__(self, nx, nu, ndu, ny, ph, ch):
        self._c = _L.LMPC(int(nx), int(nu), int(ndu), int(ny), int(ph), int(ch), batch=1)
        self._last = None
        self._init_common()

    def setOptimizerParameters(self, p):
        self._c.setOptimizerParameters(p)

    # every setter has the reference's overloads: (matrices) or (vectors, HorizonSlice)
    def setStateSpaceModel(self, A, B, C):
        return self._c.setStateSpaceModel(A, B, C)

    def setDisturbances(self, Bd, Dd):
        return self._c.setDisturbances(Bd, Dd)

    def setObjectiveWeights(self, OWeight, UWeight, DeltaUWeight, slice=None):
        return self._c.setObjectiveWeights(OWeight, UWeight, DeltaUWeight, slice)

    def setStateBounds(self, XMin, XMax, slice=None):
        return self._c.setStateBounds(XMin, XMax, slice)

    def setInputBounds(self, UMin, UMax, slice=None):
        return self._c.setInputBounds(UMin, UMax, slice)

    def setOutputBounds(self, YMin, YMax, slice=None):
        return self._c.setOutputBounds(YMin, YMax, slice)

    def setScalarConstraint(self, *args):
        """(index, min, max, X, U)  or  (min, max, X, U, HorizonSlice)   (LMPC.hpp:355,409)."""
        if len(args) != 5:
            raise TypeError("setScalarConstraint takes (index, min, max, X, U) or (min, max, X, U, slice)")
        if isinstance(args[4], HorizonSlice) or isinstance(args[4], tuple):
            smin, smax, X, U, sl = args
            return self._c.setScalarConstraint(smin, smax, X, U, sl)
        index, smin, smax, X, U = args
        return self._c.setScalarConstraint(smin, smax, X, U, HorizonSlice(int(index), int(index) + 1))

    def setExogenousInputs(self, uMeas, slice=None):
        return self._c.setExogenousInputs(uMeas, slice)

    def setReferences(self, outRef, cmdRef, deltaCmdRef, slice=None):
        return self._c.setReferences(outRef, cmdRef, deltaCmdRef, slice)

    def getSolverWarmStartPrimal(self):
        return self._c.getSolverWarmStartPrimal()[0]

    def getSolverWarmStartDual(self):
        return self._c.getSolverWarmStartDual()[0]

    def setSolverWarmStart(self, primal, dual):
        return self._c.setSolverWarmStart(primal, dual)

    def optimize(self, x0, lastU):
        r = self._timed(lambda: self._c.optimize(np.asarray(x0, float), np.asarray(lastU, float)))
        st = int(r.solver_status[0])
        self._last = Result(st, _OSQP_MSG.get(st, "unknown"), r.cost[0], r.cmd[0].copy(), r.status[0], r.is_feasible[0])
        return self._record(self._last)

    step = optimize


class NLMPC(_Common):
    """mpc::NLMPC<Dynamic...> as exported at pybind_export.cpp:59-84."""

    def __init__(self, nx, nu, ny, ph, ch, ineq, eq):
        self._dims = dict(nx=int(nx), nu=int(nu), ny=int(ny), ph=int(ph), ch=int(ch), ineq=int(ineq), eq=int(eq))
        self._c = None
        self._pending = []          # setter calls made before the system is known are replayed on the engine object
        self._last = None
        self._ts = None
        self._init_common()

    # ---- the system: CUDA source or a built-in id instead of Python callbacks
    def setSystem(self, system_id, params):
        d = self._dims
        c = _L.NLMPC(int(system_id), d["ph"], d["ch"], batch=1)
        if (c.nx, c.nu) != (d["nx"], d["nu"]) or c.nineq != d["ineq"] or c.neq != d["eq"]:
            raise ValueError(f"system has nx={c.nx} nu={c.nu} Tineq={c.nineq} Teq={c.neq}, the controller was built for {d}")
        params = np.array(params, dtype=np.float64)
        if self._ts is not None and params.size:
            params[..., 0] = self._ts          # continuous systems keep their sampling time in slot 0
        c.setSystemParameters(params)
        self._c = c
        for name, args in self._pending:
            getattr(c, name)(*args)
        self._pending = []
        return True

    def setSystemSource(self, cuda_source, type_name, params):
        return self.setSystem(_L.register_system(cuda_source, type_name), params)

    def _no_callback(self, name):
        raise RuntimeError(f"{name} takes a Python callback, which cannot run inside a CUDA kernel: give the model / cost / "
                           "constraints as CUDA source with setSystemSource(cuda_source, type_name, params) (contract in "
                           "include/b200mpc.h) or select a built-in system with setSystem(system_id, params)")

    def setStateSpaceFunction(self, handle, eq_tol=1e-10):
        self._no_callback("setStateSpaceFunction")

    def setObjectiveFunction(self, handle):
        self._no_callback("setObjectiveFunction")

    def setOutputFunction(self, handle):
        self._no_callback("setOutputFunction")

    def setIneqConFunction(self, handle, tol=1e-10):
        self._no_callback("setIneqConFunction")

    def setEqConFunction(self, handle, tol=1e-10):
        self._no_callback("setEqConFunction")

    # ---- everything else as in the reference
    def _fwd(self, name, *args):
        if self._c is None:
            self._pending.append((name, args))
            return True
        return getattr(self._c, name)(*args)

    def setDiscretizationSamplingTime(self, ts):
        self._ts = float(ts)
        if self._c is not None and self._c.params is not None and self._c.params.size:
            p = self._c.params.copy(); p[..., 0] = self._ts
            self._c.setSystemParameters(p)
        return True

    def setInputScale(self, scaling):
        return self._fwd("setInputScale", scaling)

    def setStateScale(self, scaling):
        return self._fwd("setStateScale", scaling)

    def setOptimizerParameters(self, p):
        return self._fwd("setOptimizerParameters", p)

    def setStateBounds(self, XMin, XMax, slice=None):
        return self._fwd("setStateBounds", XMin, XMax, slice)

    def setInputBounds(self, UMin, UMax, slice=None):
        return self._fwd("setInputBounds", UMin, UMax, slice)

    def setOutputBounds(self, YMin, YMax, slice=None):
        return False                                        # ignored by the reference too (NLMPC.hpp:342-349)

    def optimize(self, x0, lastU):
        if self._c is None:
            raise RuntimeError("NLMPC: no system set (setSystemSource / setSystem)")
        r = self._timed(lambda: self._c.optimize(np.asarray(x0, float), np.asarray(lastU, float)))
        st = int(r.solver_status[0])
        msg = {4: "XTOL_REACHED", 5: "MAXEVAL_REACHED"}.get(st, "unknown")         # nlopt::result names
        self._last = Result(st, msg, r.cost[0], r.cmd[0].copy(), r.status[0], r.is_feasible[0])
        return self._record(self._last)

    step = optimize
