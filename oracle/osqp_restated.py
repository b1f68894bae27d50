"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- numpy restatement of the OSQP v0.6.3 algorithm.

OSQP is a third-party dependency of the reference that is NOT under /root/reference
(pinned to tag v0.6.3 by /root/reference/configure.sh:36-38 and Dockerfile:81; call sites
include/mpc/LMPC/LOptimizer.hpp:244-284).  Its algorithm is restated here from the OSQP paper
(Stellato et al., "OSQP: an operator splitting solver for quadratic programs", 2020) and from
recollection of the v0.6.3 sources (osqp.c, auxil.c, scaling.c, polish.c) -- NOT read from a
copy of those sources.  It is pinned by the reference's own golden vector
(test/LMPC/test_common.cpp:230-236) and by a solver-independent KKT optimality check, both in
tests/test_oracle_lmpc.py.

The one non-deterministic rule of v0.6.3 -- adaptive_rho_interval==0 picks the rho-update interval from
wall-clock time (0.4 x setup time, rounded to a multiple of check_termination, at least
check_termination) -- is pinned to its smallest possible outcome, 25 iterations, which is what small
problems like these produce (setup costs more than 0.4 x 12 ADMM iterations).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

# constants.h (v0.6.3)
RHO_MIN = 1e-6
RHO_MAX = 1e6
RHO_EQ_OVER_RHO_INEQ = 1e3
RHO_TOL = 1e-4
OSQP_INFTY = 1e30
MIN_SCALING = 1e-4
MAX_SCALING = 1e4
OSQP_DIVISION_TOL = 1.0 / OSQP_INFTY

# status values (constants.h)
OSQP_DUAL_INFEASIBLE_INACCURATE = 4
OSQP_PRIMAL_INFEASIBLE_INACCURATE = 3
OSQP_SOLVED_INACCURATE = 2
OSQP_SOLVED = 1
OSQP_MAX_ITER_REACHED = -2
OSQP_PRIMAL_INFEASIBLE = -3
OSQP_DUAL_INFEASIBLE = -4
OSQP_SIGINT = -5
OSQP_TIME_LIMIT_REACHED = -6
OSQP_NON_CVX = -7
OSQP_UNSOLVED = -10


class Settings:
    """osqp_set_default_settings + the libmpc overrides (LOptimizer.hpp:244-257, Types.hpp:142-160)."""

    def __init__(self, **kw):
        self.rho = 1e-6          # LParameters::rho
        self.sigma = 1e-6
        self.scaling = 10
        self.adaptive_rho = True
        self.adaptive_rho_interval = 25   # pinned, see module docstring
        self.adaptive_rho_tolerance = 5.0
        self.max_iter = 100      # Parameters::maximum_iteration
        self.eps_abs = 1e-4
        self.eps_rel = 1e-4
        self.eps_prim_inf = 1e-3
        self.eps_dual_inf = 1e-3
        self.alpha = 1.6
        self.delta = 1e-6
        self.polish = True
        self.polish_refine_iter = 3
        self.check_termination = 25
        self.warm_start = False
        for k, v in kw.items():
            if not hasattr(self, k):
                raise KeyError(k)
            setattr(self, k, v)


def _limit_scaling(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.where(v > MAX_SCALING, MAX_SCALING, v)


class Info:
    pass


class OSQPRestated:
    """Dense-matrix OSQP v0.6.3.  P full symmetric, A dense; bounds may hold IEEE infinities (libmpc passes
    mpc::inf straight through, include/mpc/Types.hpp:228)."""

    def __init__(self, P, q, A, l, u, settings: Settings):
        self.s = settings
        self.n = P.shape[0]
        self.m = A.shape[0]
        self.P = np.array(P, float)
        self.q = np.array(q, float)
        self.A = np.array(A, float)
        self.l = np.array(l, float)
        self.u = np.array(u, float)
        self.x = np.zeros(self.n)
        self.z = np.zeros(self.m)
        self.y = np.zeros(self.m)
        self.info = Info()
        self.info.rho_updates = 0
        self.info.status_polish = 0
        self.info.iter = 0
        self._scale_data()
        self._set_rho_vec()
        self._factor()

    # ---- scaling.c: scale_data ---------------------------------------------------------------
    def _scale_data(self):
        n, m = self.n, self.m
        self.D = np.ones(n)
        self.E = np.ones(m)
        self.c = 1.0
        with np.errstate(invalid="ignore"):
            for _ in range(self.s.scaling):
                Dt = np.maximum(np.abs(self.P).max(axis=0, initial=0.0), np.abs(self.A).max(axis=0, initial=0.0))
                Et = np.abs(self.A).max(axis=1, initial=0.0)
                Dt = 1.0 / np.sqrt(_limit_scaling(Dt))
                Et = 1.0 / np.sqrt(_limit_scaling(Et))
                self.P = Dt[:, None] * self.P * Dt[None, :]
                self.A = Et[:, None] * self.A * Dt[None, :]
                self.q = Dt * self.q
                self.D *= Dt
                self.E *= Et
                c_temp = np.abs(self.P).max(axis=0, initial=0.0).mean()
                inf_norm_q = float(_limit_scaling(np.array([np.abs(self.q).max(initial=0.0)]))[0])
                c_temp = max(c_temp, inf_norm_q)
                c_temp = 1.0 / float(_limit_scaling(np.array([c_temp]))[0])
                self.P = self.P * c_temp
                self.q = self.q * c_temp
                self.c *= c_temp
        self.Dinv = 1.0 / self.D
        self.Einv = 1.0 / self.E
        self.cinv = 1.0 / self.c
        self.l = self.E * self.l
        self.u = self.E * self.u

    # ---- auxil.c: set_rho_vec ----------------------------------------------------------------
    def _set_rho_vec(self):
        s = self.s
        s.rho = min(max(s.rho, RHO_MIN), RHO_MAX)
        loose = (self.l < -OSQP_INFTY * MIN_SCALING) & (self.u > OSQP_INFTY * MIN_SCALING)
        with np.errstate(invalid="ignore"):
            eq = (~loose) & ((self.u - self.l) < RHO_TOL)
        self.constr_type = np.where(loose, -1, np.where(eq, 1, 0))
        self._update_rho_vec()

    def _update_rho_vec(self):
        rho = self.s.rho
        self.rho_vec = np.where(self.constr_type == -1, RHO_MIN,
                                np.where(self.constr_type == 1, RHO_EQ_OVER_RHO_INEQ * rho, rho))
        self.rho_inv_vec = 1.0 / self.rho_vec

    def _factor(self):
        n, m = self.n, self.m
        K = np.zeros((n + m, n + m))
        K[:n, :n] = self.P + self.s.sigma * np.eye(n)
        K[:n, n:] = self.A.T
        K[n:, :n] = self.A
        K[n:, n:] = -np.diag(self.rho_inv_vec)
        self._lu = sla.lu_factor(K)

    # ---- osqp.c: osqp_warm_start ---------------------------------------------------------------
    def warm_start(self, x, y):
        """x,y in user (unscaled) units: x <- Dinv x, y <- c Einv y, z <- A x."""
        self.s.warm_start = True
        self.x = self.Dinv * np.asarray(x, float)
        self.y = self.c * self.Einv * np.asarray(y, float)
        self.z = self.A @ self.x

    # ---- auxil.c: update_info / residuals ----------------------------------------------------
    def _update_info(self, x, z, y, polish=False):
        self.Ax = self.A @ x
        self._pri_vec = self.Ax - z           # (z_prev used as temp in C)
        pri_res = np.abs(self.Einv * self._pri_vec).max(initial=0.0)
        self.Px = self.P @ x
        self.Aty = self.A.T @ y
        self._dua_vec = self.q + self.Px + self.Aty   # (x_prev used as temp in C)
        dua_res = self.cinv * np.abs(self.Dinv * self._dua_vec).max(initial=0.0)
        if polish:
            self.pol_obj = (0.5 * x @ self.Px + self.q @ x) * self.cinv
            self.pol_pri_res, self.pol_dua_res = pri_res, dua_res
        else:
            self.info.pri_res, self.info.dua_res = pri_res, dua_res

    def _compute_rho_estimate(self):
        pri_res = np.abs(self._pri_vec).max(initial=0.0)
        dua_res = np.abs(self._dua_vec).max(initial=0.0)
        pri_norm = max(np.abs(self.z).max(initial=0.0), np.abs(self.Ax).max(initial=0.0))
        pri_res /= (pri_norm + 1e-10)
        dua_norm = max(np.abs(self.q).max(initial=0.0), np.abs(self.Aty).max(initial=0.0),
                       np.abs(self.Px).max(initial=0.0))
        dua_res /= (dua_norm + 1e-10)
        est = self.s.rho * np.sqrt(pri_res / (dua_res + 1e-10))
        return min(max(est, RHO_MIN), RHO_MAX)

    def _is_primal_infeasible(self, eps):
        l, u = self.l, self.u
        dy = self.delta_y
        up_inf = u > OSQP_INFTY * MIN_SCALING
        lo_inf = l < -OSQP_INFTY * MIN_SCALING
        dy = np.where(up_inf & lo_inf, 0.0, np.where(up_inf, np.minimum(dy, 0.0),
                                                     np.where(lo_inf, np.maximum(dy, 0.0), dy)))
        self.delta_y = dy   # projected in place, as in C
        norm_dy = np.abs(self.E * dy).max(initial=0.0)
        if norm_dy > eps:
            with np.errstate(invalid="ignore"):
                # IEEE semantics on purpose: inf * 0 = NaN, as happens in the reference when bounds are mpc::inf
                lhs = float(np.sum(u * np.maximum(dy, 0.0) + l * np.minimum(dy, 0.0)))
            if lhs < -eps * norm_dy:
                Atdy = self.Dinv * (self.A.T @ dy)
                return np.abs(Atdy).max(initial=0.0) < eps * norm_dy
        return False

    def _is_dual_infeasible(self, eps):
        dx = self.delta_x
        norm_dx = np.abs(self.D * dx).max(initial=0.0)
        cs = self.c
        if norm_dx > eps:
            if self.q @ dx < -cs * eps * norm_dx:
                Pdx = self.Dinv * (self.P @ dx)
                if np.abs(Pdx).max(initial=0.0) < cs * eps * norm_dx:
                    Adx = self.Einv * (self.A @ dx)
                    bad = ((self.u < OSQP_INFTY * MIN_SCALING) & (Adx > eps * norm_dx)) | \
                          ((self.l > -OSQP_INFTY * MIN_SCALING) & (Adx < -eps * norm_dx))
                    return not bool(bad.any())
        return False

    def _check_termination(self, approximate):
        s, info = self.s, self.info
        eps_abs, eps_rel = s.eps_abs, s.eps_rel
        eps_pi, eps_di = s.eps_prim_inf, s.eps_dual_inf
        if info.pri_res > OSQP_INFTY or info.dua_res > OSQP_INFTY:
            info.status_val = OSQP_NON_CVX
            info.obj_val = np.nan
            return True
        if approximate:
            eps_abs *= 10; eps_rel *= 10; eps_pi *= 10; eps_di *= 10
        prim_ok = dual_ok = prim_inf = dual_inf = False
        if self.m == 0:
            prim_ok = True
        else:
            eps_prim = eps_abs + eps_rel * max(np.abs(self.Einv * self.z).max(initial=0.0),
                                               np.abs(self.Einv * self.Ax).max(initial=0.0))
            if info.pri_res < eps_prim:
                prim_ok = True
            else:
                prim_inf = self._is_primal_infeasible(eps_pi)
        eps_dual = eps_abs + eps_rel * self.cinv * max(np.abs(self.Dinv * self.q).max(initial=0.0),
                                                       np.abs(self.Dinv * self.Aty).max(initial=0.0),
                                                       np.abs(self.Dinv * self.Px).max(initial=0.0))
        if info.dua_res < eps_dual:
            dual_ok = True
        else:
            dual_inf = self._is_dual_infeasible(eps_di)
        if prim_ok and dual_ok:
            info.status_val = OSQP_SOLVED_INACCURATE if approximate else OSQP_SOLVED
            return True
        if prim_inf:
            info.status_val = OSQP_PRIMAL_INFEASIBLE_INACCURATE if approximate else OSQP_PRIMAL_INFEASIBLE
            info.obj_val = OSQP_INFTY
            return True
        if dual_inf:
            info.status_val = OSQP_DUAL_INFEASIBLE_INACCURATE if approximate else OSQP_DUAL_INFEASIBLE
            info.obj_val = -OSQP_INFTY
            return True
        return False

    # ---- osqp.c: osqp_solve -------------------------------------------------------------------
    def solve(self):
        s, info, n, m = self.s, self.info, self.n, self.m
        info.status_val = OSQP_UNSOLVED
        info.obj_val = None
        if not s.warm_start:
            self.x[:] = 0; self.z[:] = 0; self.y[:] = 0
        can_check = False
        it = 0
        for it in range(1, s.max_iter + 1):
            x_prev, z_prev = self.x, self.z
            rhs = np.concatenate([s.sigma * x_prev - self.q, z_prev - self.rho_inv_vec * self.y])
            sol = sla.lu_solve(self._lu, rhs)
            xt = sol[:n]
            zt = rhs[n:] + self.rho_inv_vec * sol[n:]
            self.x = s.alpha * xt + (1 - s.alpha) * x_prev
            self.delta_x = self.x - x_prev
            zr = s.alpha * zt + (1 - s.alpha) * z_prev
            self.z = np.minimum(np.maximum(zr + self.rho_inv_vec * self.y, self.l), self.u)
            self.delta_y = self.rho_vec * (zr - self.z)
            self.y = self.y + self.delta_y
            can_check = s.check_termination and (it % s.check_termination == 0)
            if can_check:
                self._update_info(self.x, self.z, self.y)
                if self._check_termination(False):
                    break
            if s.adaptive_rho and s.adaptive_rho_interval and it % s.adaptive_rho_interval == 0:
                if not can_check:
                    self._update_info(self.x, self.z, self.y)
                rho_new = self._compute_rho_estimate()
                if rho_new > s.rho * s.adaptive_rho_tolerance or rho_new < s.rho / s.adaptive_rho_tolerance:
                    s.rho = min(max(rho_new, RHO_MIN), RHO_MAX)
                    self._update_rho_vec()
                    self._factor()
                    info.rho_updates += 1
        else:
            it = s.max_iter + 1
        if not can_check:
            self._update_info(self.x, self.z, self.y)
            self._check_termination(False)
        info.iter = it - 1 if it == s.max_iter + 1 else it
        has_solution = info.status_val not in (OSQP_PRIMAL_INFEASIBLE, OSQP_PRIMAL_INFEASIBLE_INACCURATE,
                                               OSQP_DUAL_INFEASIBLE, OSQP_DUAL_INFEASIBLE_INACCURATE,
                                               OSQP_NON_CVX)
        if has_solution:
            info.obj_val = (0.5 * self.x @ (self.P @ self.x) + self.q @ self.x) * self.cinv
        if info.status_val == OSQP_UNSOLVED:
            if not self._check_termination(True):
                info.status_val = OSQP_MAX_ITER_REACHED
        if s.polish and info.status_val == OSQP_SOLVED:
            self._polish()
        # store_solution
        if has_solution:
            self.sol_x = self.D * self.x
            self.sol_y = self.cinv * self.E * self.y
        else:
            self.sol_x = np.full(n, np.nan)
            self.sol_y = np.full(m, np.nan)
        return self.sol_x, self.sol_y, info

    # ---- polish.c ------------------------------------------------------------------------------
    def _polish(self):
        s, info, n, m = self.s, self.info, self.n, self.m
        with np.errstate(invalid="ignore"):
            low = (self.z - self.l) < -self.y
            upp = (self.u - self.z) < self.y
        ind_low = np.nonzero(low)[0]
        ind_upp = np.nonzero(upp)[0]
        self.pol_ind_low, self.pol_ind_upp = ind_low, ind_upp
        Ared = np.vstack([self.A[ind_low], self.A[ind_upp]])
        mred = Ared.shape[0]
        K = np.zeros((n + mred, n + mred))
        K[:n, :n] = self.P + s.delta * np.eye(n)
        K[:n, n:] = Ared.T
        K[n:, :n] = Ared
        K[n:, n:] = -s.delta * np.eye(mred)
        lu = sla.lu_factor(K)
        rhs = np.concatenate([-self.q, self.l[ind_low], self.u[ind_upp]])
        sol = sla.lu_solve(lu, rhs)
        for _ in range(s.polish_refine_iter):
            r = rhs.copy()
            r[:n] -= self.P @ sol[:n] + Ared.T @ sol[n:]
            r[n:] -= Ared @ sol[:n]
            sol = sol + sla.lu_solve(lu, r)
        px = sol[:n]
        pz = self.A @ px
        py = np.zeros(m)
        py[ind_low] = sol[n:n + len(ind_low)]
        py[ind_upp] = sol[n + len(ind_low):]
        # project_normalcone
        t = pz + py
        pz = np.minimum(np.maximum(t, self.l), self.u)
        py = t - pz
        self._update_info(px, pz, py, polish=True)
        ok = ((self.pol_pri_res < info.pri_res and self.pol_dua_res < info.dua_res) or
              (self.pol_pri_res < info.pri_res and info.dua_res < 1e-10) or
              (self.pol_dua_res < info.dua_res and info.pri_res < 1e-10))
        if ok:
            info.obj_val = self.pol_obj
            info.pri_res, info.dua_res = self.pol_pri_res, self.pol_dua_res
            info.status_polish = 1
            self.x, self.z, self.y = px, pz, py
        else:
            info.status_polish = -1


# ---- libmpc glue: LOptimizer::run (include/mpc/LMPC/LOptimizer.hpp:189-368) --------------------
# mpc::ResultStatus (include/mpc/Types.hpp:84-91)
SUCCESS, MAX_ITERATION, INFEASIBLE, ERROR, UNKNOWN = range(5)


def convert_status(st):
    """LOptimizer::convertToResultStatus (LOptimizer.hpp:386-415)."""
    return {OSQP_SOLVED: SUCCESS, OSQP_MAX_ITER_REACHED: MAX_ITERATION, OSQP_PRIMAL_INFEASIBLE: INFEASIBLE,
            OSQP_DUAL_INFEASIBLE: INFEASIBLE, OSQP_SOLVED_INACCURATE: SUCCESS,
            OSQP_PRIMAL_INFEASIBLE_INACCURATE: SUCCESS, OSQP_DUAL_INFEASIBLE_INACCURATE: SUCCESS,
            OSQP_SIGINT: ERROR, OSQP_TIME_LIMIT_REACHED: UNKNOWN, OSQP_NON_CVX: ERROR,
            OSQP_UNSOLVED: UNKNOWN}.get(st, UNKNOWN)


def lmpc_optimize(form, x0, u0, settings: Settings | None = None, warm=None):
    """One LMPC::optimize(x0,u0) through the oracle.  Returns a dict mirroring mpc::Result + sequences."""
    settings = settings or Settings()
    P, q, A, l, u = form.build(x0, u0)
    solver = OSQPRestated(P, q, A, l, u, settings)
    if settings.warm_start and warm is not None:
        solver.warm_start(warm[0], warm[1])
    else:
        settings.warm_start = False   # osqp_update_warm_start(work, 0)  (LOptimizer.hpp:280)
    x, y, info = solver.solve()
    state, inp, out = form.unpack(x)
    st = info.status_val
    return dict(cmd=inp[0].copy(), solver_status=st, status=convert_status(st), cost=info.obj_val,
                is_feasible=st in (OSQP_SOLVED, OSQP_SOLVED_INACCURATE, OSQP_MAX_ITER_REACHED),
                iter=info.iter, rho_updates=info.rho_updates, status_polish=info.status_polish,
                state=state, input=inp, output=out, x=x, y=y, rho=settings.rho,
                pri_res=info.pri_res, dua_res=info.dua_res, solver=solver)
