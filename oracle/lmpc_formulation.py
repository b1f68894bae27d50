"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- restatement of libmpc++'s LMPC QP formulation.

This file is a checker.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg may import it; the product path (libmpc_b200/) never does.

Restates, in plain numpy, the dense sparse-form QP that the reference hands to OSQP:

    minimise 1/2 z' P z + q' z    s.t.  l <= A z <= u

with z = [e_0 .. e_ph ; du_0 .. du_{ph-1}], e_i = [x_i ; xu_i] (xu_i = u_{i-1}).

Reference anchors (paths relative to /root/reference):
  * ProblemBuilder::onInit            include/mpc/LMPC/ProblemBuilder.hpp:88-172   (defaults)
  * setStateModel / setExogenousInput include/mpc/LMPC/ProblemBuilder.hpp:184-236  (augmented model)
  * setObjective / bounds / scalar    include/mpc/LMPC/ProblemBuilder.hpp:247-504  (column-0 duplication)
  * get()                             include/mpc/LMPC/ProblemBuilder.hpp:528-633  (per-step q,l,u)
  * buildTimeInvariantTems()          include/mpc/LMPC/ProblemBuilder.hpp:642-825  (P, A, lineq, uineq)
  * LOptimizer::run unpack            include/mpc/LMPC/LOptimizer.hpp:292-347
  * LMPC front-end slice setters      include/mpc/LMPC.hpp:111-676

Pinned by the reference's own KATs in tests/test_oracle_lmpc.py
(test/LMPC/test_constraints.cpp:169-295, test/LMPC/test_common.cpp:89-280).
"""
from __future__ import annotations

import numpy as np

INF = np.inf


class LMPCFormulation:
    """Mirror of mpc::ProblemBuilder<sizer> + the reference holders in LOptimizer (refs, exogenous)."""

    def __init__(self, nx, nu, ndu, ny, ph, ch):
        self.nx, self.nu, self.ndu, self.ny, self.ph, self.ch = nx, nu, ndu, ny, ph, ch
        ne = nx + nu
        self.ne = ne
        self.n = (ph + 1) * ne + ph * nu
        self.m_eq = (ph + 1) * ne
        self.m_ineq = (ph + 1) * ne + (ph + 1) * ny + ph * nu + (ph + 1)
        self.m = self.m_eq + self.m_ineq
        # ProblemBuilder.hpp:120-149
        self.ssA = np.zeros((ne, ne))
        self.ssB = np.zeros((ne, nu))
        self.ssC = np.zeros((nu + ny, ne))
        self.ssBv = np.zeros((ne, ndu))
        self.ssDv = np.zeros((nu + ny, ndu))
        self.wOutput = np.zeros((ny, ph + 1))
        self.wU = np.zeros((nu, ph + 1))
        self.wDeltaU = np.zeros((nu, ph))
        self.minX = np.full((nx, ph + 1), -INF)
        self.maxX = np.full((nx, ph + 1), INF)
        self.minY = np.full((ny, ph + 1), -INF)
        self.maxY = np.full((ny, ph + 1), INF)
        self.minU = np.full((nu, ph), -INF)
        self.maxU = np.full((nu, ph), INF)
        self.sMin = np.full(ph + 1, -INF)
        self.sMax = np.full(ph + 1, INF)
        self.sMultiplier = np.zeros((ph + 1, (ph + 1) * ne))
        # LOptimizer.hpp:71-79
        self.yRef = np.zeros((ny, ph))
        self.uRef = np.zeros((nu, ph))
        self.duRef = np.zeros((nu, ph))
        self.uMeas = np.zeros((ndu, ph))

    # ---- setters (ProblemBuilder.hpp:184-504) -------------------------------------------------
    def set_state_space_model(self, A, B, C):
        nx, nu, ny = self.nx, self.nu, self.ny
        self.ssA[:nx, :nx] = A
        self.ssA[:nx, nx:] = B
        self.ssA[nx:, :nx] = 0
        self.ssA[nx:, nx:] = np.eye(nu)
        self.ssB[:nx, :] = B
        self.ssB[nx:, :] = np.eye(nu)
        self.ssC[:ny, :nx] = C
        self.ssC[ny:, nx:] = np.eye(nu)

    def set_disturbances(self, Bd, Dd):
        self.ssBv[:] = 0
        self.ssBv[: self.nx, :] = Bd
        self.ssDv[:] = 0
        self.ssDv[: self.ny, :] = Dd

    def set_objective_weights(self, OW, UW, DUW):
        """Matrix form [ny x ph],[nu x ph],[nu x ph]; column 0 duplicated (ProblemBuilder.hpp:254-260)."""
        OW, UW, DUW = (np.asarray(a, float) for a in (OW, UW, DUW))
        if OW.ndim == 1:  # vector + HorizonSlice::all (LMPC.hpp:436-450)
            OW = np.tile(OW[:, None], (1, self.ph))
            UW = np.tile(UW[:, None], (1, self.ph))
            DUW = np.tile(DUW[:, None], (1, self.ph))
        self.wOutput[:, 1:] = OW
        self.wOutput[:, 0] = OW[:, 0]
        self.wU[:, 1:] = UW
        self.wU[:, 0] = UW[:, 0]
        self.wDeltaU[:, :] = DUW

    def set_objective_weights_at(self, index, OW, UW, DUW):
        self.wOutput[:, index + 1] = OW
        self.wU[:, index + 1] = UW
        if index == 0:
            self.wOutput[:, 0] = OW
            self.wU[:, 0] = UW
        self.wDeltaU[:, index] = DUW

    def set_state_bounds(self, XMin, XMax):
        XMin, XMax = np.asarray(XMin, float), np.asarray(XMax, float)
        if XMin.ndim == 1:
            XMin = np.tile(XMin[:, None], (1, self.ph))
            XMax = np.tile(XMax[:, None], (1, self.ph))
        self.minX[:, 1:] = XMin
        self.minX[:, 0] = XMin[:, 0]
        self.maxX[:, 1:] = XMax
        self.maxX[:, 0] = XMax[:, 0]

    def set_state_bounds_at(self, index, XMin, XMax):
        self.minX[:, index + 1] = XMin
        self.maxX[:, index + 1] = XMax
        if index == 0:
            self.minX[:, 0] = XMin
            self.maxX[:, 0] = XMax

    def set_input_bounds(self, UMin, UMax):
        """Matrix form [nu x ch]; tail of the prediction horizon replicated (ProblemBuilder.hpp:397-413)."""
        UMin, UMax = np.asarray(UMin, float), np.asarray(UMax, float)
        ch, ph = self.ch, self.ph
        if UMin.ndim == 1:
            UMin = np.tile(UMin[:, None], (1, ch))
            UMax = np.tile(UMax[:, None], (1, ch))
        self.minU[:, :ch] = UMin
        self.maxU[:, :ch] = UMax
        if ch < ph:
            self.minU[:, ch:] = UMin[:, ch - 1][:, None]
            self.maxU[:, ch:] = UMax[:, ch - 1][:, None]

    def set_input_bounds_at(self, index, UMin, UMax):
        self.minU[:, index] = UMin
        self.maxU[:, index] = UMax

    def set_output_bounds(self, YMin, YMax):
        YMin, YMax = np.asarray(YMin, float), np.asarray(YMax, float)
        if YMin.ndim == 1:
            YMin = np.tile(YMin[:, None], (1, self.ph))
            YMax = np.tile(YMax[:, None], (1, self.ph))
        self.minY[:, 1:] = YMin
        self.minY[:, 0] = YMin[:, 0]
        self.maxY[:, 1:] = YMax
        self.maxY[:, 0] = YMax[:, 0]

    def set_output_bounds_at(self, index, YMin, YMax):
        self.minY[:, index + 1] = YMin
        self.maxY[:, index + 1] = YMax
        if index == 0:
            self.minY[:, 0] = YMin
            self.maxY[:, 0] = YMax

    def set_scalar_constraint(self, smin, smax, X, U):
        """Whole-horizon form (ProblemBuilder.hpp:347-365). smin/smax scalars or [ph] vectors."""
        ph, ne = self.ph, self.ne
        smin = np.broadcast_to(np.asarray(smin, float), (ph,))
        smax = np.broadcast_to(np.asarray(smax, float), (ph,))
        self.sMin[1:] = smin
        self.sMin[0] = smin[0]
        self.sMax[1:] = smax
        self.sMax[0] = smax[0]
        self._set_multiplier(X, U)

    def set_scalar_constraint_at(self, index, smin, smax, X, U):
        """Per-index form; note the multiplier is overwritten for ALL stages (ProblemBuilder.hpp:329-332)."""
        self.sMin[index + 1] = smin
        self.sMax[index + 1] = smax
        if index == 0:
            self.sMin[0] = smin
            self.sMax[0] = smax
        self._set_multiplier(X, U)

    def _set_multiplier(self, X, U):
        ne = self.ne
        row = np.concatenate([np.asarray(X, float).ravel(), np.asarray(U, float).ravel()])
        for i in range(self.ph + 1):
            self.sMultiplier[i, i * ne:(i + 1) * ne] = row

    def set_references(self, yRef, uRef, duRef):
        yRef, uRef, duRef = (np.asarray(a, float) for a in (yRef, uRef, duRef))
        if yRef.ndim == 1:  # vector + slice all (LMPC.hpp:616-640)
            yRef = np.tile(yRef[:, None], (1, self.ph))
            uRef = np.tile(uRef[:, None], (1, self.ph))
            duRef = np.tile(duRef[:, None], (1, self.ph))
        self.yRef[:], self.uRef[:], self.duRef[:] = yRef, uRef, duRef

    def set_exogenous_inputs(self, uMeas):
        uMeas = np.asarray(uMeas, float)
        if uMeas.ndim == 1:
            uMeas = np.tile(uMeas[:, None], (1, self.ph))
        self.uMeas[:] = uMeas

    # ---- time-invariant terms (ProblemBuilder.hpp:642-825) -----------------------------------
    def build_PA(self):
        nx, nu, ny, ph, ch, ne = self.nx, self.nu, self.ny, self.ph, self.ch, self.ne
        n, m = self.n, self.m
        P = np.zeros((n, n))
        for i in range(ph + 1):
            W = np.zeros((ny + nu, ny + nu))
            W[:ny, :ny] = np.diag(self.wOutput[:, i])
            W[ny:, ny:] = np.diag(self.wU[:, i])
            P[i * ne:(i + 1) * ne, i * ne:(i + 1) * ne] = self.ssC.T @ W @ self.ssC
            if i < ph:
                o = (ph + 1) * ne + i * nu
                P[o:o + nu, o:o + nu] = np.diag(self.wDeltaU[:, i])
        A = np.zeros((m, n))
        # equality block  (:690-702)
        for i in range(ph + 1):
            A[i * ne:(i + 1) * ne, i * ne:(i + 1) * ne] = -np.eye(ne)
            if i > 0:
                A[i * ne:(i + 1) * ne, (i - 1) * ne:i * ne] += self.ssA
                o = (ph + 1) * ne + (i - 1) * nu
                A[i * ne:(i + 1) * ne, o:o + nu] = self.ssB
        r0 = self.m_eq
        # state/input box rows (:712-717)
        A[r0:r0 + (ph + 1) * ne, :(ph + 1) * ne] = np.eye((ph + 1) * ne)
        # output rows (:721-725)
        r1 = r0 + (ph + 1) * ne
        for i in range(ph + 1):
            A[r1 + i * ny:r1 + (i + 1) * ny, i * ne:(i + 1) * ne] = self.ssC[:ny, :]
        # delta-u rows (:768-773)
        r2 = r1 + (ph + 1) * ny
        A[r2:r2 + ph * nu, (ph + 1) * ne:] = np.eye(ph * nu)
        # scalar rows (:797-801)
        r3 = r2 + ph * nu
        A[r3:r3 + ph + 1, :(ph + 1) * ne] = self.sMultiplier
        # bounds (:727-809)
        lineq = np.zeros(self.m_ineq)
        uineq = np.zeros(self.m_ineq)
        for i in range(ph + 1):
            j = i - 1 if i == ph else i
            lineq[i * ne:(i + 1) * ne] = np.concatenate([self.minX[:, i], self.minU[:, j]])
            uineq[i * ne:(i + 1) * ne] = np.concatenate([self.maxX[:, i], self.maxU[:, j]])
        o1 = (ph + 1) * ne
        lineq[o1:o1 + (ph + 1) * ny] = self.minY.T.ravel()   # column-major flatten of [ny x (ph+1)]
        uineq[o1:o1 + (ph + 1) * ny] = self.maxY.T.ravel()
        o2 = o1 + (ph + 1) * ny
        for i in range(ph):
            frozen = i > ch  # NB: strictly greater (ProblemBuilder.hpp:784-785)
            lineq[o2 + i * nu:o2 + (i + 1) * nu] = 0.0 if frozen else -INF
            uineq[o2 + i * nu:o2 + (i + 1) * nu] = 0.0 if frozen else INF
        o3 = o2 + ph * nu
        lineq[o3:o3 + ph + 1] = self.sMin
        uineq[o3:o3 + ph + 1] = self.sMax
        return P, A, lineq, uineq

    # ---- per-step terms (ProblemBuilder.hpp:528-633) -----------------------------------------
    def build_qlu(self, x0, u0, lineq, uineq):
        nx, nu, ny, ph, ne = self.nx, self.nu, self.ny, self.ph, self.ne
        q = np.zeros(self.n)
        leq = np.zeros(self.m_eq)
        off = np.zeros(self.m_ineq)
        for i in range(ph + 1):
            j = 0 if i == 0 else i - 1
            eRef = np.concatenate([self.yRef[:, j], self.uRef[:, j]])
            d = self.uMeas[:, j]
            W = np.zeros((ny + nu, ny + nu))
            W[:ny, :ny] = np.diag(self.wOutput[:, i])
            W[ny:, ny:] = np.diag(self.wU[:, i])
            q[i * ne:(i + 1) * ne] = self.ssC.T @ W @ (-eRef + self.ssDv @ d)
            if i < ph:
                o = (ph + 1) * ne + i * nu
                q[o:o + nu] = -(self.wDeltaU[:, i] * self.duRef[:, j])
            if i > 0:
                leq[i * ne:(i + 1) * ne] = -self.ssBv @ d
            o = i * ny + (ph + 1) * ne
            off[o:o + ny] = -self.ssDv[:ny, :] @ d
        leq[:nx] = -np.asarray(x0, float)
        leq[nx:ne] = -np.asarray(u0, float)
        l = np.concatenate([leq, lineq + off])
        u = np.concatenate([leq, uineq + off])
        return q, l, u

    def build(self, x0, u0):
        P, A, lineq, uineq = self.build_PA()
        q, l, u = self.build_qlu(x0, u0, lineq, uineq)
        return P, q, A, l, u

    # ---- unpack (LOptimizer.hpp:305-341, ProblemBuilder.hpp:514-517) ---------------------------
    def unpack(self, z):
        nx, nu, ny, ph, ne = self.nx, self.nu, self.ny, self.ph, self.ne
        state = np.zeros((ph + 1, nx))
        inp = np.zeros((ph + 1, nu))
        out = np.zeros((ph + 1, ny))
        for i in range(ph + 1):
            state[i] = z[i * ne:i * ne + nx]
            k = i + 1 if i + 1 < ph + 1 else i
            inp[i] = z[k * ne + nx:(k + 1) * ne]
            j = 0 if i == 0 else i - 1
            out[i] = self.ssC[:ny, :nx] @ state[i] + self.ssDv[:ny, :] @ self.uMeas[:, j]
        return state, inp, out


def discretization(A, B, Ts):
    """c2d via matrix exponential of [[A,B],[0,0]]*Ts (include/mpc/Utils.hpp:23-47)."""
    from scipy.linalg import expm
    nx, nu = B.shape
    M = np.zeros((nx + nu, nx + nu))
    M[:nx, :nx] = A * Ts
    M[:nx, nx:] = B * Ts
    E = expm(M)
    return E[:nx, :nx], E[:nx, nx:]


def quadrotor_model():
    """Ad, Bd of examples/quadrotor_ex.cpp:19-45 (data constants, not code)."""
    Ad = np.array([
        [1, 0, 0, 0, 0, 0, 0.1, 0, 0, 0, 0, 0],
        [0, 1, 0, 0, 0, 0, 0, 0.1, 0, 0, 0, 0],
        [0, 0, 1, 0, 0, 0, 0, 0, 0.1, 0, 0, 0],
        [0.0488, 0, 0, 1, 0, 0, 0.0016, 0, 0, 0.0992, 0, 0],
        [0, -0.0488, 0, 0, 1, 0, 0, -0.0016, 0, 0, 0.0992, 0],
        [0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0.0992],
        [0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0],
        [0.9734, 0, 0, 0, 0, 0, 0.0488, 0, 0, 0.9846, 0, 0],
        [0, -0.9734, 0, 0, 0, 0, 0, -0.0488, 0, 0, 0.9846, 0],
        [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.9846]], float)
    Bd = np.array([
        [0, -0.0726, 0, 0.0726],
        [-0.0726, 0, 0.0726, 0],
        [-0.0152, 0.0152, -0.0152, 0.0152],
        [0, -0.0006, -0.0000, 0.0006],
        [0.0006, 0, -0.0006, 0],
        [0.0106, 0.0106, 0.0106, 0.0106],
        [0, -1.4512, 0, 1.4512],
        [-1.4512, 0, 1.4512, 0],
        [-0.3049, 0.3049, -0.3049, 0.3049],
        [0, -0.0236, 0, 0.0236],
        [0.0236, 0, -0.0236, 0],
        [0.2107, 0.2107, 0.2107, 0.2107]], float)
    return Ad, Bd


def quadrotor_formulation(ph=10, ch=None, kat_scalar_rows=False):
    """The quadrotor LMPC of examples/quadrotor_ex.cpp / test/LMPC/test_common.cpp:89-237."""
    ch = ph if ch is None else ch
    f = LMPCFormulation(12, 4, 4, 12, ph, ch)
    Ad, Bd = quadrotor_model()
    f.set_state_space_model(Ad, Bd, np.eye(12))
    f.set_objective_weights(np.array([0, 0, 10, 10, 10, 10, 0, 0, 0, 5, 5, 5.0]),
                            np.full(4, 0.1), np.zeros(4))
    xmin = np.full(12, -INF)
    xmax = np.full(12, INF)
    xmin[0] = xmin[1] = -np.pi / 6
    xmin[5] = -1
    xmax[0] = xmax[1] = np.pi / 6
    f.set_state_bounds(xmin, xmax)
    f.set_output_bounds(np.full(12, -INF), np.full(12, INF))
    u0 = 10.5916
    f.set_input_bounds(np.full(4, 9.6) - u0, np.full(4, 13.0) - u0)
    if kat_scalar_rows:  # test_common.cpp:209-210: multiplier of ones with infinite bounds
        f.set_scalar_constraint(-INF, INF, np.ones(12), np.ones(4))
    yref = np.zeros(12)
    yref[2] = 1.0
    f.set_references(yref, np.zeros(4), np.zeros(4))
    return f
