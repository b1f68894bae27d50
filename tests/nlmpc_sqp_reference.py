"""Host-side executable specification of the batched SQP the CUDA NLMPC kernel implements (test infrastructure).

Algorithm (per instance): damped-BFGS SQP on the reference's multiple-shooting NLP
    min f(z)  s.t.  c_eq(z) = 0,  c_in(z) <= 0,  lb <= z <= ub
with f, gradients and Jacobians evaluated exactly like the reference does (oracle/nlmpc_formulation.py: forward /
central finite differences with the reference's step rules).  Each QP subproblem
    min 1/2 d'Bd + g'd   s.t.  J_eq d = -c_eq,  J_in d <= -c_in,  lb - z <= d <= ub - z
is solved by a dense OSQP-style ADMM (Ruiz equilibration, rho by row class, over-relaxation 1.6, adaptive rho) run to a
moderate accuracy and then polished (active-set guess + regularised KKT solve with iterative refinement, OSQP polish.c),
warm started from the previous SQP iteration's multipliers; the step is globalised by an L1 merit function (one penalty
per constraint row, Powell's update as in Kraft's SLSQP) with backtracking; a failed line search restarts B = I once (as SLSQP does) and stops at the second failure in a row.
"""
import numpy as np


class QPADMM:
    """Dense OSQP-style ADMM: Jacobi equilibration, rho_eq = 1e3 rho, alpha = 1.6, residual test + adaptive rho (OSQP's
    estimate, refactor when it moves by more than 5x) every `check` iterations."""

    def __init__(self, rho=0.1, sigma=1e-6, alpha=1.6, max_iter=200, eps=1e-5, check=25, polish=True, delta=1e-6, refine=5,
                 spd_factor=None, carry_rho=False):
        # carry_rho: start every QP from the penalty the previous one ended with (as OSQP does between re-solves) -- what the
        # stage-structured kernel does (nlmpc_structured.cuh); the dense kernel restarts from `rho` every time
        self.carry_rho, self.rho_init = carry_rho, rho
        self.rho, self.sigma, self.alpha, self.max_iter, self.eps, self.check = rho, sigma, alpha, max_iter, eps, check
        self.polish, self.delta, self.refine = polish, delta, refine
        # H -> (rhs -> H^-1 rhs).  Default: dense Cholesky (what the CUDA kernel does today); the stage-structured kernel
        # plugs in the bordered block-tridiagonal factorisation of tests/nlmpc_structured_kkt_reference.py.
        self.spd_factor = spd_factor or self._dense_factor

    @staticmethod
    def _dense_factor(H):
        L = np.linalg.cholesky(H)
        return lambda rhs: np.linalg.solve(L.T, np.linalg.solve(L, rhs))

    def solve(self, B, g, A, l, u, x=None, y=None):
        n, m = B.shape[0], A.shape[0]
        # Ruiz equilibration of [B A'; A 0] with cost normalisation (OSQP scaling.c), 10 passes
        D = np.ones(n); E = np.ones(m); c = 1.0
        Bs = B.copy(); As = A.copy(); gs = g.copy()
        lim = lambda v: np.minimum(np.where(v < 1e-4, 1.0, v), 1e4)
        for _ in range(10):
            dn = np.maximum(np.abs(Bs).max(0), np.abs(As).max(0) if m else 0.0)
            en = np.abs(As).max(1) if m else np.zeros(0)
            dt = 1.0 / np.sqrt(lim(dn)); et = 1.0 / np.sqrt(lim(en))
            Bs = dt[:, None] * Bs * dt[None, :]
            As = et[:, None] * As * dt[None, :]
            gs = dt * gs
            D *= dt; E *= et
            ct = max(np.abs(Bs).max(0).mean(), float(lim(np.array([np.abs(gs).max()]))[0]))
            ct = 1.0 / float(lim(np.array([ct]))[0])
            Bs *= ct; gs *= ct; c *= ct
        ls, us = E * l, E * u
        eq = (us - ls) < 1e-9
        loose = (l < -1e20) & (u > 1e20)          # OSQP: rows without bounds get RHO_MIN
        rho0 = self.rho
        def factor(rho0):
            rho = np.where(loose, 1e-6, np.where(eq, 1e3 * rho0, rho0))
            H = Bs + self.sigma * np.eye(n) + As.T @ (rho[:, None] * As)
            return rho, self.spd_factor(H)
        rho, kkt_solve = factor(rho0)
        xs = np.zeros(n) if x is None else x / D
        ys = np.zeros(m) if y is None else c * y / E
        zs = np.clip(As @ xs, ls, us)
        it = 0
        for it in range(1, self.max_iter + 1):
            rhs = self.sigma * xs - gs + As.T @ (rho * zs - ys)
            xt = kkt_solve(rhs)
            zt = As @ xt
            xn = self.alpha * xt + (1 - self.alpha) * xs
            zr = self.alpha * zt + (1 - self.alpha) * zs
            zn = np.clip(zr + ys / rho, ls, us)
            ys = ys + rho * (zr - zn)
            xs, zs = xn, zn
            if it % self.check == 0:
                Ax = As @ xs
                Px = Bs @ xs
                Aty = As.T @ ys
                pri = np.abs(Ax - zs).max() if m else 0.0
                dua = np.abs(Px + gs + Aty).max()
                if pri < self.eps and dua < self.eps:
                    break
                pn = pri / (max(np.abs(zs).max(initial=0.0), np.abs(Ax).max(initial=0.0)) + 1e-10)
                dn = dua / (max(np.abs(gs).max(), np.abs(Aty).max(), np.abs(Px).max()) + 1e-10)
                est = min(max(rho0 * np.sqrt(pn / (dn + 1e-10)), 1e-6), 1e6)
                if est > 5 * rho0 or est < rho0 / 5:
                    rho0 = est
                    rho, kkt_solve = factor(rho0)
        if self.carry_rho:
            self.rho = rho0
        if self.polish:
            # OSQP polish.c: guess the active set from (z, y), solve the equality-constrained QP on it through the same
            # reduced system (rho = 1/delta on active rows, sigma = delta) with iterative refinement, keep it if both
            # residuals improve.
            lo = (zs - ls) < -ys
            up = (us - zs) < ys
            act = lo | up
            b = np.where(lo, ls, us)[act]
            Aa = As[act]
            dl = self.delta
            Hp = Bs + dl * np.eye(n) + Aa.T @ Aa / dl
            polish_solve = self.spd_factor(Hp)
            def kkt(r1, r2):           # [Bs+dl I, Aa'; Aa, -dl I] [x; y] = [r1; r2]
                x = polish_solve(r1 + Aa.T @ r2 / dl)
                return x, (Aa @ x - r2) / dl
            xp, yp = kkt(-gs, b)
            for _ in range(self.refine):
                r1 = -gs - Bs @ xp - Aa.T @ yp
                r2 = b - Aa @ xp
                dx, dy = kkt(r1, r2)
                xp, yp = xp + dx, yp + dy
            yf = np.zeros(m); yf[act] = yp
            Axp = As @ xp
            pri_p = np.maximum(np.maximum(ls - Axp, Axp - us), 0).max() if m else 0.0
            dua_p = np.abs(Bs @ xp + gs + As.T @ yf).max()
            Ax = As @ xs
            pri_a = np.maximum(np.maximum(ls - Ax, Ax - us), 0).max() if m else 0.0
            dua_a = np.abs(Bs @ xs + gs + As.T @ ys).max()
            self.polished = bool(pri_p <= max(pri_a, 1e-10) and dua_p <= max(dua_a, 1e-10))
            if self.polished:
                xs, ys = xp, yf
        return D * xs, E * ys / c, it


def stage_groups(f):
    """Index groups of a per-stage block-diagonal quasi-Newton matrix: stage s holds X_{s+1} and, when s opens a control
    block, that block; the slack is its own group."""
    ph, ch, nx, nu = f.ph, f.ch, f.nx, f.nu
    G = []
    for s in range(ph):
        idx = list(range(s * nx, (s + 1) * nx))
        if s < ch:
            idx += list(range(ph * nx + s * nu, ph * nx + (s + 1) * nu))
        G.append(np.array(idx))
    G.append(np.array([f.nz - 1]))
    return G


def sqp_solve(f, x0, z0, lb, ub, max_sqp=100, tol=1e-7, ftol=1e-12, qp=None, verbose=False, bfgs_groups=None):
    """bfgs_groups=None: the dense damped BFGS the CUDA kernel implements.  bfgs_groups=stage_groups(f): the same update applied
    block by block (B stays block diagonal, so the QP's reduced KKT matrix is block tridiagonal over the stages) -- the
    specification of the stage-structured kernel (libmpc_b200/csrc/nlmpc_structured.cuh)."""
    qp = qp or QPADMM()
    qp_cap0 = qp.max_iter
    qp.rho = qp.rho_init
    x0 = np.asarray(x0, float)
    z = np.clip(np.array(z0, float), lb, ub)
    n = z.size
    B = np.eye(n)
    fval, g = f.objective(z, x0)
    ce, Je = f.state_eq(z, x0)
    # dense block = [user inequalities ; user equalities] (the kernel's J_in block, first mii rows are inequalities)
    def dense(zz, jac=True):
        ci_, Ji_ = f.ineq_con(zz, x0) if f.ineq is not None else (np.zeros(0), np.zeros((0, n)))
        if getattr(f, "eq", None) is not None:
            cu_, Ju_ = f.eq_con(zz, x0)
            ci_, Ji_ = np.concatenate([ci_, cu_]), np.vstack([Ji_, Ju_])
        return ci_, Ji_
    mii = f.nineq if f.ineq is not None else 0
    ci, Ji = dense(z)
    me, mi = ce.size, ci.size
    muv = None
    y_prev = None
    resets, just_reset = 0, False
    hist = []
    for k in range(max_sqp):
        A = np.vstack([Je, Ji, np.eye(n)])
        l = np.concatenate([-ce, np.full(mii, -np.inf), -ci[mii:], lb - z])
        u = np.concatenate([-ce, -ci, ub - z])
        d, y, qit = qp.solve(B, g, A, l, u, None, y_prev)
        y_prev = y
        lam_e, lam_i = y[:me], y[me:me + mi]
        viol = lambda ce_, ci_: np.abs(ce_).sum() + np.maximum(ci_[:mii], 0).sum() + np.abs(ci_[mii:]).sum()
        rowv = lambda ce_, ci_: np.concatenate([np.abs(ce_), np.maximum(ci_[:mii], 0), np.abs(ci_[mii:])])
        v0 = viol(ce, ci)
        # one penalty per constraint row, Powell's update as in Kraft's SLSQP: mu_r <- max(|lambda_r|, (mu_r + |lambda_r|) / 2)
        lam = np.abs(y[:me + mi])
        muv = lam.copy() if (k == 0 or just_reset) else np.maximum(lam, 0.5 * (muv + lam))
        wv = lambda ce_, ci_: float(muv @ rowv(ce_, ci_))
        phi0 = fval + wv(ce, ci)
        dphi = g @ d - wv(ce, ci)         # directional derivative bound of the L1 merit
        # Kraft's first stopping test (SLSQP: |g'd| and the violation below the accuracy): nothing left to gain
        if abs(g @ d) < ftol * max(1.0, abs(fval)) and v0 < 1e-8:
            hist.append((k, fval, v0, np.abs(d).max(), 0.0, qit))
            break
        t = 1.0
        ls_ok = False
        for _ in range(25):
            zt = z + t * d
            ft, _ = f.objective(zt, x0, want_grad=False)
            cet, _ = f.state_eq(zt, x0, want_jac=False)
            cit = dense(zt)[0]
            if ft + wv(cet, cit) <= phi0 + 1e-4 * t * dphi:
                ls_ok = True
                break
            t *= 0.5
        if not ls_ok:
            # no decrease of the merit along d at any step length.  Like SLSQP, restart the quasi-Newton matrix once
            # (a poor B or an inexact QP step); failing again straight after the restart is the finite-difference noise floor.
            hist.append((k, fval, v0, np.abs(d).max(), 0.0, qit))
            if qp.max_iter == qp_cap0:         # first suspect an inexact QP step: re-solve (and go on) with a 5x ADMM cap
                qp.max_iter = 5 * qp_cap0
                continue
            if just_reset or resets >= 5:
                break
            B = np.eye(n); resets += 1; just_reset = True
            continue
        just_reset = False
        s = t * d
        z_new = z + s
        f_new, g_new = f.objective(z_new, x0)
        ce_new, Je_new = f.state_eq(z_new, x0)
        ci_new, Ji_new = dense(z_new)
        # damped BFGS on the Lagrangian gradient
        gl_new = g_new + Je_new.T @ lam_e + Ji_new.T @ lam_i
        gl_old = g + Je.T @ lam_e + Ji.T @ lam_i
        yk = gl_new - gl_old
        def damped_bfgs(Bm, sv, yv):
            Bs = Bm @ sv
            sBs = sv @ Bs
            sy = sv @ yv
            if sBs > 1e-300:
                theta = 1.0 if sy >= 0.2 * sBs else 0.8 * sBs / (sBs - sy)
                r = theta * yv + (1 - theta) * Bs
                return Bm - np.outer(Bs, Bs) / sBs + np.outer(r, r) / (sv @ r)
            return Bm
        if bfgs_groups is None:
            B = damped_bfgs(B, s, yk)
        else:
            for G in bfgs_groups:
                B[np.ix_(G, G)] = damped_bfgs(B[np.ix_(G, G)], s[G], yk[G])
        step = np.abs(d).max()            # the full QP step: small only at a KKT point (t*d can be small far from one)
        hist.append((k, f_new, v0, step, t, qit))
        if verbose:
            print(hist[-1])
        z, fval, g, ce, Je, ci, Ji = z_new, f_new, g_new, ce_new, Je_new, ci_new, Ji_new
        if step < tol * max(1.0, np.abs(z).max()) and viol(ce, ci) < 1e-8:
            break
    qp.max_iter = qp_cap0
    X, U, e = f.unwrap(z, x0)
    return dict(z=z, cmd=U[0].copy(), cost=fval, nit=k + 1,
                viol=float(np.abs(ce).sum() + np.maximum(ci[:mii], 0).sum() + np.abs(ci[mii:]).sum()), hist=hist)
