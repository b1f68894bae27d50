"""Host-side executable specification of the STRUCTURED algorithm the CUDA kernels implement
(libmpc_b200/csrc/lmpc_kernels.cuh).  Test infrastructure: it lets the CPU-only suite check -- without a GPU --
that the stage-structured arithmetic (structured Ruiz scaling, block-tridiagonal reduced-KKT Cholesky, penalty-form
polish with iterative refinement) follows the same iterate path as the dense OSQP restatement in oracle/.

Never imported by the product path.
"""
import numpy as np

from oracle.osqp_restated import (MAX_SCALING, MIN_SCALING, OSQP_INFTY, RHO_EQ_OVER_RHO_INEQ, RHO_MAX, RHO_MIN,
                                  RHO_TOL)


def _lim(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.where(v > MAX_SCALING, MAX_SCALING, v)


class StructuredLMPC:
    """Everything is held per row-group / per stage, never as an (m x n) matrix.

    row groups: eq[(ph+1),ne]  box[(ph+1),ne]  out[(ph+1),ny]  du[ph,nu]  sc[(ph+1)]
    variables : e[(ph+1),ne]   du[ph,nu]
    """

    def __init__(self, form):
        f = form
        self.f = f
        self.nx, self.nu, self.ny, self.ph, self.ne = f.nx, f.nu, f.ny, f.ph, f.ne
        self.G = np.hstack([f.ssA, f.ssB])          # ne x (ne+nu)
        self.C = f.ssC[:f.ny, :f.nx]
        # the scalar multiplier is one row for every stage (ProblemBuilder.hpp:329-332,359-362)
        self.s = f.sMultiplier[0, :f.ne].copy()
        self.Pe = np.zeros((f.ph + 1, f.ne, f.ne))
        for i in range(f.ph + 1):
            self.Pe[i, :f.nx, :f.nx] = self.C.T @ np.diag(f.wOutput[:, i]) @ self.C
            self.Pe[i, f.nx:, f.nx:] = np.diag(f.wU[:, i])
        self.Pdu = f.wDeltaU.T.copy()               # ph x nu (diagonal)

    # -------- operators on the UNSCALED structure ------------------------------------------------
    def A_mul(self, e, du):
        ph, ne, nx = self.ph, self.ne, self.nx
        eq = -e.copy()
        w = np.hstack([e[:ph], du])                 # ph x (ne+nu)
        eq[1:] += w @ self.G.T
        box = e.copy()
        out = e[:, :nx] @ self.C.T
        sc = e @ self.s
        return eq, box, out, du.copy(), sc

    def At_mul(self, eq, box, out, dur, sc):
        ph, ne, nx = self.ph, self.ne, self.nx
        e = -eq + box
        e[:, :nx] += out @ self.C
        e += sc[:, None] * self.s[None, :]
        w = eq[1:] @ self.G                          # ph x (ne+nu)
        e[:ph] += w[:, :ne]
        du = w[:, ne:] + dur
        return e, du

    def P_mul(self, e, du):
        return np.einsum("ijk,ik->ij", self.Pe, e), self.Pdu * du

    # -------- structured Ruiz equilibration (scaling.c: scale_data) -------------------------------
    def scale(self, q_e, q_du, iters=10):
        ph, ne, nx, nu, ny = self.ph, self.ne, self.nx, self.nu, self.ny
        G, C, s = np.abs(self.G), np.abs(self.C), np.abs(self.s)
        absPe = np.abs(self.Pe)
        De, Ddu = np.ones((ph + 1, ne)), np.ones((ph, nu))
        Eeq, Ebox, Eout = np.ones((ph + 1, ne)), np.ones((ph + 1, ne)), np.ones((ph + 1, ny))
        Edu, Esc = np.ones((ph, nu)), np.ones(ph + 1)
        c = 1.0
        qe, qdu = q_e.copy(), q_du.copy()
        for _ in range(iters):
            Dw = np.hstack([De[:ph], Ddu])
            # column norms of [P;A]
            ce = np.maximum(c * (De[:, :, None] * absPe * De[:, None, :]).max(axis=1), Eeq * De)
            ce = np.maximum(ce, Ebox * De)
            ce[:, :nx] = np.maximum(ce[:, :nx], (Eout[:, :, None] * C[None] * De[:, None, :nx]).max(axis=1))
            ce = np.maximum(ce, Esc[:, None] * s[None, :] * De)
            colw = (Eeq[1:, :, None] * G[None] * Dw[:, None, :]).max(axis=1)      # ph x (ne+nu)
            ce[:ph] = np.maximum(ce[:ph], colw[:, :ne])
            cdu = np.maximum(c * np.abs(self.Pdu) * Ddu * Ddu, colw[:, ne:])
            cdu = np.maximum(cdu, Edu * Ddu)
            # row norms of A
            req = Eeq * De
            req[1:] = np.maximum(req[1:], (Eeq[1:, :, None] * G[None] * Dw[:, None, :]).max(axis=2))
            rbox = Ebox * De
            rout = (Eout[:, :, None] * C[None] * De[:, None, :nx]).max(axis=2) if nx > 0 else np.zeros_like(Eout)
            rdu = Edu * Ddu
            rsc = (Esc[:, None] * s[None, :] * De).max(axis=1)
            te, tdu = 1 / np.sqrt(_lim(ce)), 1 / np.sqrt(_lim(cdu))
            De, Ddu = De * te, Ddu * tdu
            qe, qdu = qe * te, qdu * tdu
            Eeq, Ebox, Eout = Eeq / np.sqrt(_lim(req)), Ebox / np.sqrt(_lim(rbox)), Eout / np.sqrt(_lim(rout))
            Edu, Esc = Edu / np.sqrt(_lim(rdu)), Esc / np.sqrt(_lim(rsc))
            # cost normalisation: mean of P column norms (over all n columns), inf-norm of q
            pe = c * (De[:, :, None] * absPe * De[:, None, :]).max(axis=1)
            pdu = c * np.abs(self.Pdu) * Ddu * Ddu
            n = pe.size + pdu.size
            c_temp = (pe.sum() + pdu.sum()) / n
            nq = max(np.abs(qe).max(initial=0.0), np.abs(qdu).max(initial=0.0))
            nq = float(_lim(np.array([nq]))[0])
            c_temp = 1.0 / float(_lim(np.array([max(c_temp, nq)]))[0])
            c *= c_temp
            qe, qdu = qe * c_temp, qdu * c_temp
        self.De, self.Ddu, self.c = De, Ddu, c
        self.E = (Eeq, Ebox, Eout, Edu, Esc)
        self.qe, self.qdu = qe, qdu                   # scaled q = c*D*q

    # -------- scaled operators --------------------------------------------------------------------
    def As(self, e, du):
        r = self.A_mul(self.De * e, self.Ddu * du)
        return tuple(E * v for E, v in zip(self.E, r))

    def Ats(self, rows):
        e, du = self.At_mul(*(E * v for E, v in zip(self.E, rows)))
        return self.De * e, self.Ddu * du

    def Ps(self, e, du):
        pe, pdu = self.P_mul(self.De * e, self.Ddu * du)
        return self.c * self.De * pe, self.c * self.Ddu * pdu

    # -------- block tridiagonal reduced KKT:  Hbar = Pbar + sigma I + Abar' R Abar -------------------
    def factor(self, rho_rows, sigma):
        """rho_rows: per-row weights (same grouping as rows).  Builds Linv_i (b x b) and Lc_i (ne x b)."""
        ph, ne, nx, nu = self.ph, self.ne, self.nx, self.nu
        rp = tuple(r * E * E for r, E in zip(rho_rows, self.E))      # rho' = rho E^2 on unscaled rows
        req, rbox, rout, rdu, rsc = rp
        self.Linv, self.Lc = [], []
        S_carry = None
        for i in range(ph + 1):
            b = ne + nu if i < ph else ne
            Hu = np.zeros((b, b))
            Hu[:ne, :ne] = self.c * self.Pe[i] + np.diag(req[i] + rbox[i]) + rsc[i] * np.outer(self.s, self.s)
            Hu[:nx, :nx] += self.C.T @ (rout[i][:, None] * self.C)
            if i < ph:
                Hu += self.G.T @ (req[i + 1][:, None] * self.G)
                Hu[ne:, ne:] += np.diag(self.c * self.Pdu[i] + rdu[i])
                Dw = np.concatenate([self.De[i], self.Ddu[i]])
            else:
                Dw = self.De[i]
            H = Dw[:, None] * Hu * Dw[None, :] + sigma * np.eye(b)
            if S_carry is not None:
                H[:ne, :ne] -= S_carry
            L = np.linalg.cholesky(H)
            Linv = np.linalg.inv(L)
            self.Linv.append(Linv)
            if i < ph:
                Hc = -(self.De[i + 1] * req[i + 1])[:, None] * self.G * Dw[None, :]      # ne x b
                Lc = Hc @ Linv.T
                self.Lc.append(Lc)
                S_carry = Lc @ Lc.T

    def solve(self, re, rdu):
        """Solve Hbar [e;du] = [re;rdu]."""
        ph, ne = self.ph, self.ne
        t = []
        for i in range(ph + 1):
            r = np.concatenate([re[i], rdu[i]]) if i < ph else re[i].copy()
            if i > 0:
                r[:ne] -= self.Lc[i - 1] @ t[i - 1]
            t.append(self.Linv[i] @ r)
        e = np.zeros_like(re)
        du = np.zeros_like(rdu)
        nxt = None
        for i in range(ph, -1, -1):
            r = t[i].copy()
            if i < ph:
                r -= self.Lc[i].T @ nxt
            w = self.Linv[i].T @ r
            e[i] = w[:ne]
            if i < ph:
                du[i] = w[ne:]
            nxt = w[:ne]
        return e, du


def rows_cat(rows):
    return np.concatenate([np.asarray(r).ravel() for r in rows])


def split_rows(v, ph, ne, ny, nu):
    o = 0
    out = []
    for shape in ((ph + 1, ne), (ph + 1, ne), (ph + 1, ny), (ph, nu), (ph + 1,)):
        k = int(np.prod(shape))
        out.append(v[o:o + k].reshape(shape))
        o += k
    return tuple(out)


def solve_structured(form, x0, u0, settings):
    """Full OSQP-semantics solve on the structured representation; returns the same dict keys as
    oracle.osqp_restated.lmpc_optimize that the tests compare."""
    from oracle import osqp_restated as O
    s = settings
    f = form
    ph, ne, nx, nu, ny = f.ph, f.ne, f.nx, f.nu, f.ny
    S = StructuredLMPC(f)
    _, _, lineq, uineq = f.build_PA()
    q, l, u = f.build_qlu(x0, u0, lineq, uineq)
    qe = q[:(ph + 1) * ne].reshape(ph + 1, ne)
    qdu = q[(ph + 1) * ne:].reshape(ph, nu)
    S.scale(qe, qdu, s.scaling)
    Ecat = rows_cat(S.E)
    ls, us = Ecat * l, Ecat * u
    loose = (ls < -OSQP_INFTY * MIN_SCALING) & (us > OSQP_INFTY * MIN_SCALING)
    with np.errstate(invalid="ignore"):
        eq = (~loose) & ((us - ls) < RHO_TOL)
    ctype = np.where(loose, -1, np.where(eq, 1, 0))
    rho = min(max(s.rho, RHO_MIN), RHO_MAX)

    def rho_vec_of(rho):
        return np.where(ctype == -1, RHO_MIN, np.where(ctype == 1, RHO_EQ_OVER_RHO_INEQ * rho, rho))

    rv = rho_vec_of(rho)
    sp = lambda v: split_rows(v, ph, ne, ny, nu)
    S.factor(sp(rv), s.sigma)
    xe, xdu = np.zeros((ph + 1, ne)), np.zeros((ph, nu))
    m = l.size
    z, y = np.zeros(m), np.zeros(m)
    Dcat = np.concatenate([S.De.ravel(), S.Ddu.ravel()])
    qs = np.concatenate([S.qe.ravel(), S.qdu.ravel()])
    cinv = 1.0 / S.c
    status = O.OSQP_UNSOLVED
    rho_updates = 0
    info = {}

    def update_info(xe, xdu, z, y):
        Ax = rows_cat(S.As(xe, xdu))
        pv = Ax - z
        pe, pdu = S.Ps(xe, xdu)
        Px = np.concatenate([pe.ravel(), pdu.ravel()])
        ae, adu = S.Ats(sp(y))
        Aty = np.concatenate([ae.ravel(), adu.ravel()])
        dv = qs + Px + Aty
        return dict(Ax=Ax, pv=pv, Px=Px, Aty=Aty, dv=dv, pri=np.abs(pv / Ecat).max(),
                    dua=cinv * np.abs(dv / Dcat).max())

    def check(ii, z, approx):
        k = 10.0 if approx else 1.0
        eps_prim = k * s.eps_abs + k * s.eps_rel * max(np.abs(z / Ecat).max(), np.abs(ii["Ax"] / Ecat).max())
        eps_dual = k * s.eps_abs + k * s.eps_rel * cinv * max(np.abs(qs / Dcat).max(), np.abs(ii["Aty"] / Dcat).max(),
                                                          np.abs(ii["Px"] / Dcat).max())
        return ii["pri"] < eps_prim and ii["dua"] < eps_dual

    it_done = 0
    ii = None
    can_check = False
    for it in range(1, s.max_iter + 1):
        ye, ydu = S.Ats(sp(rv * z - y))
        re = s.sigma * xe - S.qe + ye
        rdu = s.sigma * xdu - S.qdu + ydu
        te, tdu = S.solve(re, rdu)
        zt = rows_cat(S.As(te, tdu))
        xe_n = s.alpha * te + (1 - s.alpha) * xe
        xdu_n = s.alpha * tdu + (1 - s.alpha) * xdu
        zr = s.alpha * zt + (1 - s.alpha) * z
        z_n = np.minimum(np.maximum(zr + y / rv, ls), us)
        y = y + rv * (zr - z_n)
        xe, xdu, z = xe_n, xdu_n, z_n
        it_done = it
        can_check = it % s.check_termination == 0
        if can_check:
            ii = update_info(xe, xdu, z, y)
            if check(ii, z, False):
                status = O.OSQP_SOLVED
                break
        if s.adaptive_rho and it % s.adaptive_rho_interval == 0:
            if not can_check:
                ii = update_info(xe, xdu, z, y)
            pr = np.abs(ii["pv"]).max() / (max(np.abs(z).max(), np.abs(ii["Ax"]).max()) + 1e-10)
            dr = np.abs(ii["dv"]).max() / (max(np.abs(qs).max(), np.abs(ii["Aty"]).max(), np.abs(ii["Px"]).max()) + 1e-10)
            est = min(max(rho * np.sqrt(pr / (dr + 1e-10)), RHO_MIN), RHO_MAX)
            if est > rho * s.adaptive_rho_tolerance or est < rho / s.adaptive_rho_tolerance:
                rho = est
                rv = rho_vec_of(rho)
                S.factor(sp(rv), s.sigma)
                rho_updates += 1
    if not can_check:
        ii = update_info(xe, xdu, z, y)
        if check(ii, z, False):
            status = O.OSQP_SOLVED
    if status == O.OSQP_UNSOLVED:
        status = O.OSQP_SOLVED_INACCURATE if check(ii, z, True) else O.OSQP_MAX_ITER_REACHED
    xs = np.concatenate([xe.ravel(), xdu.ravel()])
    obj = (0.5 * xs @ ii["Px"] + qs @ xs) * cinv if ii is not None else None
    pe_, pdu_ = S.Ps(xe, xdu)
    obj = (0.5 * xs @ np.concatenate([pe_.ravel(), pdu_.ravel()]) + qs @ xs) * cinv
    status_polish = 0
    if s.polish and status == O.OSQP_SOLVED:
        with np.errstate(invalid="ignore"):
            low = (z - ls) < -y
            upp = ((us - z) < y) & ~low
        act = (low | upp).astype(float)
        bact = np.where(low, ls, np.where(upp, us, 0.0))
        S.factor(sp(act / s.delta), s.delta)

        def reg_solve(r1e, r1du, r2):
            ae, adu = S.Ats(sp(act * r2 / s.delta))
            de, ddu = S.solve(r1e + ae, r1du + adu)
            nu_ = act * (rows_cat(S.As(de, ddu)) - r2) / s.delta
            return de, ddu, nu_

        pe, pdu, pnu = reg_solve(-S.qe, -S.qdu, bact)
        for _ in range(s.polish_refine_iter):
            Pe_, Pdu_ = S.Ps(pe, pdu)
            ae, adu = S.Ats(sp(pnu))
            r1e, r1du = -S.qe - Pe_ - ae, -S.qdu - Pdu_ - adu
            r2 = act * (bact - rows_cat(S.As(pe, pdu)))
            de, ddu, dnu = reg_solve(r1e, r1du, r2)
            pe, pdu, pnu = pe + de, pdu + ddu, pnu + dnu
        pz = rows_cat(S.As(pe, pdu))
        t = pz + pnu
        pz2 = np.minimum(np.maximum(t, ls), us)
        py = t - pz2
        pi = update_info(pe, pdu, pz2, py)
        ok = ((pi["pri"] < ii["pri"] and pi["dua"] < ii["dua"]) or (pi["pri"] < ii["pri"] and ii["dua"] < 1e-10) or
              (pi["dua"] < ii["dua"] and ii["pri"] < 1e-10))
        if ok:
            xe, xdu, z, y = pe, pdu, pz2, py
            xs = np.concatenate([xe.ravel(), xdu.ravel()])
            obj = (0.5 * xs @ pi["Px"] + qs @ xs) * cinv
            status_polish = 1
            info["pri"], info["dua"] = pi["pri"], pi["dua"]
        else:
            status_polish = -1
    x_out = Dcat * np.concatenate([xe.ravel(), xdu.ravel()])
    y_out = cinv * Ecat * y
    state, inp, out = f.unpack(x_out)
    return dict(cmd=inp[0].copy(), solver_status=status, iter=it_done, rho_updates=rho_updates,
                status_polish=status_polish, cost=obj, x=x_out, y=y_out, rho=rho, D=Dcat, E=Ecat, c=S.c)
