"""Round-2 groundwork (CPU, numpy): the SQP of tests/nlmpc_sqp_reference.py with a BLOCK-DIAGONAL (per-stage) damped BFGS instead of
the dense one.  With a block-diagonal B the reduced KKT matrix H = B + sigma I + J' R J of the QP subproblem is block tridiagonal
over the stages (up to the move-blocking border), i.e. the LMPC kernel's O(ph b^3) factorisation applies instead of the dense
O(nz^3) one.  Result (profiles/r01_block_bfgs_experiment.txt): same optimum as the dense variant, in FEWER major iterations."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import nlmpc_sqp_reference as R
from oracle import nlmpc_slsqp as S
from oracle.nlmpc_formulation import oscnet_formulation, ugv_formulation, vanderpol_formulation

def sqp(f, x0, z0, lb, ub, groups=None, max_sqp=200, tol=1e-7, ftol=1e-12):
    """copy of tests/nlmpc_sqp_reference.sqp_solve with an optional block-diagonal BFGS (groups = list of index arrays)"""
    qp = R.QPADMM()
    z = np.clip(np.array(z0, float), lb, ub); n = z.size
    B = np.eye(n)
    fval, g = f.objective(z, x0); ce, Je = f.state_eq(z, x0)
    ci, Ji = f.ineq_con(z, x0) if f.ineq is not None else (np.zeros(0), np.zeros((0, n)))
    me, mi = ce.size, ci.size
    mu = 1.0; y_prev = None; resets = 0; just_reset = False
    viol = lambda a, b: np.abs(a).sum() + np.maximum(b, 0).sum()
    for k in range(max_sqp):
        A = np.vstack([Je, Ji, np.eye(n)])
        l = np.concatenate([-ce, np.full(mi, -np.inf), lb - z]); u = np.concatenate([-ce, -ci, ub - z])
        d, y, qit = qp.solve(B, g, A, l, u, None, y_prev); y_prev = y
        lam_e, lam_i = y[:me], y[me:me + mi]
        v0 = viol(ce, ci)
        mu = max(mu, 1.1 * (np.abs(y[:me + mi]).max() if me + mi else 0.0))
        phi0 = fval + mu * v0; dphi = g @ d - mu * v0
        if abs(g @ d) < ftol * max(1.0, abs(fval)) and v0 < 1e-8: break
        t = 1.0; ok = False
        for _ in range(25):
            zt = z + t * d
            ft, _ = f.objective(zt, x0, want_grad=False); cet, _ = f.state_eq(zt, x0, want_jac=False)
            cit = f.ineq_con(zt, x0)[0] if f.ineq is not None else np.zeros(0)
            if ft + mu * viol(cet, cit) <= phi0 + 1e-4 * t * dphi: ok = True; break
            t *= 0.5
        if not ok:
            if just_reset or resets >= 5: break
            B = np.eye(n); resets += 1; just_reset = True; continue
        just_reset = False
        s = t * d; zn = z + s
        fn, gn = f.objective(zn, x0); cen, Jen = f.state_eq(zn, x0)
        cin, Jin = f.ineq_con(zn, x0) if f.ineq is not None else (np.zeros(0), np.zeros((0, n)))
        yk = (gn + Jen.T @ lam_e + Jin.T @ lam_i) - (g + Je.T @ lam_e + Ji.T @ lam_i)
        def upd(Bm, sv, yv):
            Bs = Bm @ sv; sBs = sv @ Bs; sy = sv @ yv
            if sBs > 1e-300:
                th = 1.0 if sy >= 0.2 * sBs else 0.8 * sBs / (sBs - sy)
                r = th * yv + (1 - th) * Bs
                return Bm - np.outer(Bs, Bs) / sBs + np.outer(r, r) / (sv @ r)
            return Bm
        if groups is None: B = upd(B, s, yk)
        else:
            for G in groups:
                B[np.ix_(G, G)] = upd(B[np.ix_(G, G)], s[G], yk[G])
        step = np.abs(d).max()
        z, fval, g, ce, Je, ci, Ji = zn, fn, gn, cen, Jen, cin, Jin
        if step < tol * max(1.0, np.abs(z).max()) and viol(ce, ci) < 1e-8: break
    return dict(z=z, cost=fval, nit=k + 1, viol=viol(ce, ci))

def groups_for(f):
    ph, ch, nx, nu = f.ph, f.ch, f.nx, f.nu
    G = []
    for s in range(ph):        # stage s: X_{s+1} together with its control block if this is the block's first stage
        idx = list(range(s * nx, (s + 1) * nx))
        if s < ch: idx += list(range(ph * nx + s * nu, ph * nx + (s + 1) * nu))
        G.append(np.array(idx))
    G.append(np.array([f.nz - 1]))
    return G

cases = []
f = vanderpol_formulation(); cases.append(("vanderpol", f, True, [np.array([0.0, 1.0]), np.array([0.8, -0.4]), np.array([-1.2, 0.7])]))
f = ugv_formulation(10, 10, v_pref=(0.6, 0.8)); cases.append(("ugv10", f, False, [np.zeros(4), np.array([0.4, 0.5, 0.6, 0.8])]))
f = oscnet_formulation(4, 15, 8); cases.append(("oscnet4", f, True, [np.random.default_rng(1).uniform(-1, 1, 8)]))
for name, f, hard, starts in cases:
    lb, ub = S.default_bounds(f, hard)
    if not hard: lb[-1] = 0.0
    for x0 in starts:
        z0 = S.initial_guess(f, x0, np.zeros(f.nu), lb=lb, ub=ub)
        t = time.time(); a = sqp(f, x0, z0, lb, ub); ta = time.time() - t
        t = time.time(); b = sqp(f, x0, z0, lb, ub, groups=groups_for(f)); tb = time.time() - t
        print(name, "dense: it %d cost %.9g viol %.1e | blockdiag: it %d cost %.9g viol %.1e | dz %.2e" % (a["nit"], a["cost"], a["viol"], b["nit"], b["cost"], b["viol"], np.abs(a["z"] - b["z"]).max()), flush=True)
