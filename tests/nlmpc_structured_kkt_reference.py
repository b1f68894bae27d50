"""Executable specification (CPU, numpy; test infrastructure) of the linear algebra the stage-structured NLMPC kernel will run
(DESIGN.md 8b): the reduced KKT matrix of the SQP's QP subproblem

    H = c D B D + sigma I + (E A D)' diag(rho) (E A D),      A = [J_eq ; J_in ; I]   (nlmpc_sqp.cuh: nl_factor)

when B is block diagonal over the stages (tests/nlmpc_sqp_reference.py: stage_groups).  Reorder the decision vector stage by
stage, g_s = [U_s (if stage s opens a control block and it is not the LAST block) ; X_s], and put the last control block and
the slack into a border b.  Then, because the dynamics rows of stage s read only X_{s-1}, X_s and one control block
(Constraints.hpp:844-905) and an inequality row reads one stage (ineq_per_stage),

    H = [ T  C ]     T block tridiagonal over the stages (blocks <= nx+nu), C = coupling to the border (nu+1 columns),
        [ C' D ]

and  H x = r  is solved by a block Cholesky of T (O(ph (nx+nu)^3) instead of O(nz^3)) plus a Schur complement of size nu+1:
    S = D - C' T^-1 C ;   y = T^-1 r_T ;  x_b = S^-1 (r_b - C' y) ;  x_T = y - (T^-1 C) x_b.
The move-blocking tail (stages s >= ch share the last block) is what the border is for; with ch == ph the border is one
control block + the slack all the same.  Checked against the dense solve by tests/test_nlmpc_structured_kkt.py."""
import numpy as np


def stage_partition(f):
    """-> (groups, border): index arrays into z = [X_1..X_ph ; U_1..U_ch ; slack] (Mapping.hpp:196-201)."""
    ph, ch, nx, nu = f.ph, f.ch, f.nx, f.nu
    groups = []
    for s in range(ph):
        idx = []
        if s < ch - 1:
            idx += list(range(ph * nx + s * nu, ph * nx + (s + 1) * nu))
        idx += list(range(s * nx, (s + 1) * nx))
        groups.append(np.array(idx))
    border = np.array(list(range(ph * nx + (ch - 1) * nu, ph * nx + ch * nu)) + [f.nz - 1])
    return groups, border


class BorderedBlockTridiagonal:
    """Factor and solve H given ONLY its stage blocks: diag[s] = H[g_s, g_s], sub[s] = H[g_{s+1}, g_s], C[s] = H[g_s, b], D = H[b, b]."""

    def __init__(self, diag, sub, C, D):
        self.n = len(diag)
        self.L, self.Lsub = [], []                      # block Cholesky of T: T = L L', L block lower bidiagonal
        prev = None
        for s in range(self.n):
            A = diag[s].copy()
            if s > 0:
                A -= prev @ prev.T                      # Schur complement of the previous stage
            Ls = np.linalg.cholesky(A)
            self.L.append(Ls)
            if s + 1 < self.n:
                prev = np.linalg.solve(Ls, sub[s].T).T  # Lsub_s = sub_s L_s^-T
                self.Lsub.append(prev)
        self.TiC = self._solve_T([c.copy() for c in C])             # T^-1 C, stage by stage
        S = D - sum(c.T @ t for c, t in zip(C, self.TiC))
        self.LS = np.linalg.cholesky(S)
        self.C = C

    def _solve_T(self, r):
        y = []
        for s in range(self.n):                         # forward:  L y = r
            v = r[s] - (self.Lsub[s - 1] @ y[s - 1] if s > 0 else 0.0)
            y.append(np.linalg.solve(self.L[s], v))
        x = [None] * self.n
        for s in range(self.n - 1, -1, -1):             # backward: L' x = y
            v = y[s] - (self.Lsub[s].T @ x[s + 1] if s + 1 < self.n else 0.0)
            x[s] = np.linalg.solve(self.L[s].T, v)
        return x

    def solve(self, rT, rb):
        y = self._solve_T(rT)
        t = rb - sum(c.T @ v for c, v in zip(self.C, y))
        xb = np.linalg.solve(self.LS.T, np.linalg.solve(self.LS, t))
        return [v - tc @ xb for v, tc in zip(y, self.TiC)], xb


def assemble_blocks(H, groups, border):
    """Cut the stage blocks out of a dense H and report the largest entry OUTSIDE the bordered block-tridiagonal pattern."""
    mask = np.zeros_like(H, dtype=bool)
    for s, g in enumerate(groups):
        mask[np.ix_(g, g)] = True
        if s + 1 < len(groups):
            mask[np.ix_(groups[s + 1], g)] = True; mask[np.ix_(g, groups[s + 1])] = True
    mask[border, :] = True; mask[:, border] = True
    outside = float(np.abs(np.where(mask, 0.0, H)).max())
    diag = [H[np.ix_(g, g)] for g in groups]
    sub = [H[np.ix_(groups[s + 1], groups[s])] for s in range(len(groups) - 1)]
    C = [H[np.ix_(g, border)] for g in groups]
    return diag, sub, C, H[np.ix_(border, border)], outside


def structured_factor(f):
    """spd_factor hook for tests/nlmpc_sqp_reference.QPADMM: factor H through its stage blocks only."""
    groups, border = stage_partition(f)

    def factor(H):
        diag, sub, C, D, outside = assemble_blocks(H, groups, border)
        if outside != 0.0:
            raise ValueError(f"H is not bordered block tridiagonal (largest outside entry {outside})")
        F = BorderedBlockTridiagonal(diag, sub, C, D)

        def solve(rhs):
            xT, xb = F.solve([rhs[g] for g in groups], rhs[border])
            x = np.empty_like(rhs)
            for g, v in zip(groups, xT):
                x[g] = v
            x[border] = xb
            return x
        return solve
    return factor
