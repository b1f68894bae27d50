"""Host-side executable specification of the batched SQP the CUDA NLMPC kernel implements (test infrastructure).

Algorithm (per instance): damped-BFGS SQP on the reference's multiple-shooting NLP
    min f(z)  s.t.  c_eq(z) = 0,  c_in(z) <= 0,  lb <= z <= ub
with f, gradients and Jacobians evaluated exactly like the reference does (oracle/nlmpc_formulation.py: forward /
central finite differences with the reference's step rules).  Each QP subproblem
    min 1/2 d'Bd + g'd   s.t.  J_eq d = -c_eq,  J_in d <= -c_in,  lb - z <= d <= ub - z
is solved by a dense OSQP-style ADMM (Jacobi row/column equilibration, rho_eq = 1e3 rho, over-relaxation 1.6, fixed
iteration budget with residual test), warm started from the previous SQP iteration; the step is globalised by an L1
merit function with backtracking.
"""
import numpy as np


class QPADMM:
    """Dense OSQP-style ADMM: Jacobi equilibration, rho_eq = 1e3 rho, alpha = 1.6, residual test + adaptive rho (OSQP's
    estimate, refactor when it moves by more than 5x) every `check` iterations."""

    def __init__(self, rho=0.1, sigma=1e-6, alpha=1.6, max_iter=1000, eps=1e-9, check=25):
        self.rho, self.sigma, self.alpha, self.max_iter, self.eps, self.check = rho, sigma, alpha, max_iter, eps, check

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
            return rho, np.linalg.cholesky(H)
        rho, L = factor(rho0)
        xs = np.zeros(n) if x is None else x / D
        ys = np.zeros(m) if y is None else c * y / E
        zs = np.clip(As @ xs, ls, us)
        it = 0
        for it in range(1, self.max_iter + 1):
            rhs = self.sigma * xs - gs + As.T @ (rho * zs - ys)
            xt = np.linalg.solve(L.T, np.linalg.solve(L, rhs))
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
                    rho, L = factor(rho0)
        return D * xs, E * ys / c, it


def sqp_solve(f, x0, z0, lb, ub, max_sqp=60, tol=1e-7, qp=None, verbose=False):
    qp = qp or QPADMM()
    x0 = np.asarray(x0, float)
    z = np.clip(np.array(z0, float), lb, ub)
    n = z.size
    B = np.eye(n)
    fval, g = f.objective(z, x0)
    ce, Je = f.state_eq(z, x0)
    ci, Ji = f.ineq_con(z, x0) if f.ineq is not None else (np.zeros(0), np.zeros((0, n)))
    me, mi = ce.size, ci.size
    mu = 1.0
    d_prev = y_prev = None
    hist = []
    for k in range(max_sqp):
        A = np.vstack([Je, Ji, np.eye(n)])
        l = np.concatenate([-ce, np.full(mi, -np.inf), lb - z])
        u = np.concatenate([-ce, -ci, ub - z])
        d, y, qit = qp.solve(B, g, A, l, u, d_prev, y_prev)
        lam_e, lam_i = y[:me], y[me:me + mi]
        viol = lambda ce_, ci_: np.abs(ce_).sum() + np.maximum(ci_, 0).sum()
        v0 = viol(ce, ci)
        mu = max(mu, 1.1 * (np.abs(y[:me + mi]).max() if me + mi else 0.0))
        phi0 = fval + mu * v0
        dphi = g @ d - mu * v0            # directional derivative bound of the L1 merit
        t = 1.0
        for _ in range(25):
            zt = z + t * d
            ft, _ = f.objective(zt, x0, want_grad=False)
            cet, _ = f.state_eq(zt, x0, want_jac=False)
            cit = f.ineq_con(zt, x0)[0] if f.ineq is not None else np.zeros(0)
            if ft + mu * viol(cet, cit) <= phi0 + 1e-4 * t * dphi:
                break
            t *= 0.5
        s = t * d
        z_new = z + s
        f_new, g_new = f.objective(z_new, x0)
        ce_new, Je_new = f.state_eq(z_new, x0)
        ci_new, Ji_new = f.ineq_con(z_new, x0) if f.ineq is not None else (np.zeros(0), np.zeros((0, n)))
        # damped BFGS on the Lagrangian gradient
        gl_new = g_new + Je_new.T @ lam_e + Ji_new.T @ lam_i
        gl_old = g + Je.T @ lam_e + Ji.T @ lam_i
        yk = gl_new - gl_old
        Bs = B @ s
        sBs = s @ Bs
        sy = s @ yk
        if sBs > 1e-300:
            theta = 1.0 if sy >= 0.2 * sBs else 0.8 * sBs / (sBs - sy)
            r = theta * yk + (1 - theta) * Bs
            B = B - np.outer(Bs, Bs) / sBs + np.outer(r, r) / (s @ r)
        step = np.abs(d).max()            # the full QP step: small only at a KKT point (t*d can be small far from one)
        hist.append((k, f_new, v0, step, t, qit))
        if verbose:
            print(hist[-1])
        z, fval, g, ce, Je, ci, Ji = z_new, f_new, g_new, ce_new, Je_new, ci_new, Ji_new
        d_prev, y_prev = None, y
        if step < tol * max(1.0, np.abs(z).max()) and viol(ce, ci) < 1e-8:
            break
    X, U, e = f.unwrap(z, x0)
    return dict(z=z, cmd=U[0].copy(), cost=fval, nit=k + 1, viol=float(np.abs(ce).sum() + np.maximum(ci, 0).sum()), hist=hist)
