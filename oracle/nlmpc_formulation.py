"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- numpy restatement of libmpc++'s NLMPC problem formulation: the decision
vector mapping, the objective with its forward-difference gradient, the multiple-shooting dynamics equality
constraints with central-difference Jacobians, and the user inequality / equality constraints with their Jacobians.

Reference anchors (paths relative to /root/reference):
  Mapping::unwrapVector / computeMapping     include/mpc/NLMPC/Mapping.hpp:174-257
  Model::getOutput                           include/mpc/NLMPC/Model.hpp:72-96
  Objective::evaluate / computeGradient      include/mpc/NLMPC/Objective.hpp:91-187,198-265
  Constraints::getStateEqConstraints         include/mpc/NLMPC/Constraints.hpp:490-628
  Constraints::computeStateEqJacobian        include/mpc/NLMPC/Constraints.hpp:844-905
  Constraints::evaluateIneq / computeIneqJacobian   :211-316, 641-721
  Constraints::evaluateEq / computeEqJacobian       :365-442, 731-832
  Constraints::glueJacobian                  :455-482

The reference's finite-difference quirks are reproduced on purpose (hard part 4):
  * steps use Xa.array()(j) / Ua.array()(j): LINEAR (column-major) index j of the (ph+1) x n matrix, not (row, j);
    computeEqJacobian instead uses Xa(ix,j) and Ua(ph-1,j);
  * the objective gradient is a FORWARD difference, every constraint Jacobian a CENTRAL one;
  * objective / eq Jacobians perturb rows ph-1 and ph of U together, the ineq Jacobian perturbs each of the ph rows alone;
  * constraint Jacobians are multiplied by the state scaling, the objective gradient is not.

Pinned by the reference's own KATs in tests/test_oracle_nlmpc.py (test/NLMPC/test_common.cpp:46-106,
test_objective.cpp:9-63, test_constraints.cpp:60-274).  Only tests/, smoke() and bench.py's CPU leg may import this.
"""
from __future__ import annotations

import numpy as np

DV = np.sqrt(np.finfo(float).eps)      # Objective.hpp:283, Constraints.hpp (same constant)


class NLMPCFormulation:
    def __init__(self, nx, nu, ny, ph, ch, nineq=0, neq=0):
        self.nx, self.nu, self.ny, self.ph, self.ch, self.nineq, self.neq = nx, nu, ny, ph, ch, nineq, neq
        self.nz = ph * nx + ch * nu + 1
        self.input_scaling = np.ones(nu)
        self.state_scaling = np.ones(nx)
        self.continuous, self.Ts = False, 0.0
        self.f = None          # f(x, u, i) -> dx or x+
        self.out = None        # out(x, u, i) -> y
        self.obj = None        # obj(X, Y, U, e) -> float
        self.ineq = None       # ineq(X, Y, U, e) -> [nineq]
        self.eq = None         # eq(X, U) -> [neq]
        self._compute_mapping()

    # ---- Mapping.hpp:221-257 ------------------------------------------------------------------------------------
    def _compute_mapping(self):
        nu, ph, ch = self.nu, self.ph, self.ch
        m = np.ones(ch, dtype=int)
        m[ch - 1] = ph - ch + 1
        self.Iz2u = np.zeros((ph * nu, nu * ch))
        self.Iu2z = np.zeros((nu * ch, ph * nu))
        Sz2u = np.diag(self.input_scaling)
        Su2z = np.diag(1.0 / self.input_scaling)
        ix = jx = 0
        for i in range(ch):
            self.Iu2z[ix:ix + nu, jx:jx + nu] = Su2z
            for _ in range(m[i]):
                self.Iz2u[jx:jx + nu, ix:ix + nu] = Sz2u
                jx += nu
            ix += nu

    def set_input_scaling(self, s):
        self.input_scaling = np.asarray(s, float)
        self._compute_mapping()

    def set_state_scaling(self, s):
        self.state_scaling = np.asarray(s, float)

    # ---- Mapping.hpp:174-211 ------------------------------------------------------------------------------------
    def unwrap(self, z, x0):
        nx, nu, ph, ch = self.nx, self.nu, self.ph, self.ch
        z = np.asarray(z, float)
        uvec = z[ph * nx: ph * nx + nu * ch]
        U = np.zeros((ph + 1, nu))
        U[:ph] = (self.Iz2u @ uvec).reshape(ph, nu)
        U[ph] = U[ph - 1]
        X = np.zeros((ph + 1, nx))
        X[0] = x0
        X[1:] = z[:ph * nx].reshape(ph, nx)
        X = X / self.state_scaling[None, :]
        return X, U, float(z[-1])

    def output(self, X, U):
        Y = np.zeros((self.ph + 1, self.ny))
        if self.out is not None:
            for i in range(self.ph + 1):
                Y[i] = self.out(X[i], U[i], i)
        return Y

    # ---- Objective.hpp:91-187,198-265 --------------------------------------------------------------------------------
    def objective(self, z, x0, want_grad=True):
        nx, nu, ph, ch = self.nx, self.nu, self.ph, self.ch
        X, U, e = self.unwrap(z, x0)
        fuser = lambda Xm, Um, em: float(self.obj(Xm, self.output(Xm, Um), Um, em))
        f0 = fuser(X, U, e)
        if not want_grad:
            return f0, None
        Jx = np.zeros((nx, ph))
        Jmv = np.zeros((nu, ph))
        Xa = np.maximum(np.abs(X), 1.0).ravel(order="F")     # .array()(j): linear, column-major
        Xw = X.copy()
        for i in range(ph):
            for j in range(nx):
                dx = DV * Xa[j]
                Xw[i + 1, j] = Xw[i + 1, j] + dx
                f = fuser(Xw, U, e)
                Xw[i + 1, j] = Xw[i + 1, j] - dx
                Jx[j, i] = (f - f0) / dx
        Ua = np.maximum(np.abs(U), 1.0).ravel(order="F")
        Uw = U.copy()
        for i in range(ph - 1):
            for j in range(nu):
                du = DV * Ua[j]
                Uw[i, j] = Uw[i, j] + du
                f = fuser(Xw, Uw, e)
                Uw[i, j] = Uw[i, j] - du
                Jmv[j, i] = (f - f0) / du
        for j in range(nu):
            du = DV * Ua[j]
            Uw[ph - 1, j] += du
            Uw[ph, j] += du
            f = fuser(Xw, Uw, e)
            Uw[ph - 1, j] -= du
            Uw[ph, j] -= du
            Jmv[j, ph - 1] = (f - f0) / du
        ea = max(DV, abs(e))
        de = ea * DV
        Je = (fuser(Xw, Uw, e + de) - fuser(Xw, Uw, e - de)) / (2 * de)
        g = np.zeros(self.nz)
        g[:ph * nx] = Jx.ravel(order="F")
        g[ph * nx: ph * nx + nu * ch] = self.Iz2u.T @ Jmv.ravel(order="F")
        g[-1] = Je
        return f0, g

    # ---- Constraints.hpp:455-482 ---------------------------------------------------------------------------------------
    def _glue(self, Jstate, Jmanvar, Jcon):
        nx, nu, ph, ch = self.nx, self.nu, self.ph, self.ch
        J = np.zeros((Jstate.shape[0], self.nz))
        J[:, :ph * nx] = Jstate
        J[:, ph * nx: ph * nx + nu * ch] = Jmanvar @ self.Iz2u
        J[:, -1] = Jcon
        return J

    def _scale_state_cols(self, J):
        J = J.copy()
        for k in range(self.ph):
            J[:, k * self.nx:(k + 1) * self.nx] *= self.state_scaling[None, :]
        return J

    # ---- Constraints.hpp:844-905 ----------------------------------------------------------------------------------------
    def _state_jac(self, x, u, p):
        nx, nu = self.nx, self.nu
        A = np.zeros((nx, nx))
        B = np.zeros((nx, nu))
        Xa = np.maximum(np.abs(x), 1.0)
        for i in range(nx):
            dx = DV * Xa[i]
            xp, xm = x.copy(), x.copy()
            xp[i] += dx
            xm[i] -= dx
            A[:, i] = (np.asarray(self.f(xp, u, p), float) - np.asarray(self.f(xm, u, p), float)) / (2 * dx)
        Ua = np.maximum(np.abs(u), 1.0)
        for i in range(nu):
            du = DV * Ua[i]
            up, um = u.copy(), u.copy()
            up[i] += du
            um[i] -= du
            B[:, i] = (np.asarray(self.f(x, up, p), float) - np.asarray(self.f(x, um, p), float)) / (2 * du)
        return A, B

    # ---- Constraints.hpp:490-628 ----------------------------------------------------------------------------------------
    def state_eq(self, z, x0, want_jac=True):
        nx, nu, ph = self.nx, self.nu, self.ph
        X, U, _ = self.unwrap(z, x0)
        c = np.zeros(ph * nx)
        Jx = np.zeros((ph * nx, ph * nx))
        Jmv = np.zeros((ph * nx, ph * nu))
        Ix = np.eye(nx)
        Sx = np.diag(1.0 / self.state_scaling)
        Tx = np.diag(self.state_scaling)
        for i in range(ph):
            uk, xk, xk1 = U[i].copy(), X[i].copy(), X[i + 1].copy()
            r = slice(i * nx, (i + 1) * nx)
            if self.continuous:
                h = self.Ts / 2.0
                fk = np.asarray(self.f(xk, uk, i), float)
                fk1 = np.asarray(self.f(xk1, uk, i), float)
                c[r] = (xk + h * (fk + fk1) - xk1) / self.state_scaling
                if want_jac:
                    Ak, Bk = self._state_jac(xk, uk, i)
                    Ak1, Bk1 = self._state_jac(xk1, uk, i)
                    if i > 0:
                        Jx[r, (i - 1) * nx:i * nx] = Ix + h * Sx @ Ak @ Tx
                    Jx[r, i * nx:(i + 1) * nx] = -Ix + h * Sx @ Ak1 @ Tx
                    Jmv[r, i * nu:(i + 1) * nu] = h * Sx @ (Bk + Bk1)
            else:
                xn = np.asarray(self.f(xk, uk, i), float)
                c[r] = (xk1 - xn) / self.state_scaling
                if want_jac:
                    Ak, Bk = self._state_jac(xk, uk, i)
                    Jx[r, i * nx:(i + 1) * nx] = Ix
                    if i > 0:
                        Jx[r, (i - 1) * nx:i * nx] = -(Sx @ Ak @ Tx)
                    Jmv[r, i * nu:(i + 1) * nu] = -(Sx @ Bk)
        J = self._glue(Jx, Jmv, np.zeros(ph * nx)) if want_jac else np.zeros((ph * nx, self.nz))
        return c, J

    # ---- Constraints.hpp:211-316,641-721 ---------------------------------------------------------------------------------
    def ineq_con(self, z, x0):
        nx, nu, ph = self.nx, self.nu, self.ph
        if self.ineq is None:
            return np.zeros(self.nineq), np.zeros((self.nineq, self.nz))
        X, U, e = self.unwrap(z, x0)
        g = lambda Xm, Um, em: np.asarray(self.ineq(Xm, self.output(Xm, Um), Um, em), float)
        val = g(X, U, e)
        Jx = np.zeros((self.nineq, ph * nx))
        Jmv = np.zeros((self.nineq, ph * nu))
        Xa = np.maximum(np.abs(X), 1.0).ravel(order="F")
        Xw, Uw = X.copy(), U.copy()
        for i in range(ph):
            for j in range(nx):
                dx = DV * Xa[j]
                Xw[i + 1, j] += dx
                fp = g(Xw, Uw, e)
                Xw[i + 1, j] -= 2 * dx
                fm = g(Xw, Uw, e)
                Xw[i + 1, j] += dx
                Jx[:, i * nx + j] = (fp - fm) / (2 * dx)
        Ua = np.maximum(np.abs(U), 1.0).ravel(order="F")
        for i in range(ph):
            for j in range(nu):
                du = DV * Ua[j]
                Uw[i, j] += du
                fp = g(Xw, Uw, e)
                Uw[i, j] -= 2 * du
                fm = g(Xw, Uw, e)
                Uw[i, j] += du
                Jmv[:, i * nu + j] = (fp - fm) / (2 * du)
        ea = max(DV, abs(e))
        de = ea * DV
        Je = (g(Xw, Uw, e + de) - g(Xw, Uw, e - de)) / (2 * de)
        return val, self._scale_state_cols(self._glue(Jx, Jmv, Je))

    # ---- Constraints.hpp:365-442,731-832 ---------------------------------------------------------------------------------
    def eq_con(self, z, x0):
        nx, nu, ph = self.nx, self.nu, self.ph
        if self.eq is None:
            return np.zeros(self.neq), np.zeros((self.neq, self.nz))
        X, U, _ = self.unwrap(z, x0)
        g = lambda Xm, Um: np.asarray(self.eq(Xm, Um), float)
        val = g(X, U)
        Jx = np.zeros((self.neq, ph * nx))
        Jmv = np.zeros((self.neq, ph * nu))
        Xa = np.maximum(np.abs(X), 1.0)
        Xw, Uw = X.copy(), U.copy()
        for i in range(ph):
            for j in range(nx):
                dx = DV * Xa[i + 1, j]
                Xw[i + 1, j] += dx
                fp = g(Xw, Uw)
                Xw[i + 1, j] -= 2 * dx
                fm = g(Xw, Uw)
                Xw[i + 1, j] += dx
                Jx[:, i * nx + j] = (fp - fm) / (2 * dx)
        Ua = np.maximum(np.abs(U), 1.0)
        for i in range(ph - 1):
            for j in range(nu):
                du = DV * Ua[ph - 1, j]
                Uw[i, j] += du
                fp = g(Xw, Uw)
                Uw[i, j] -= 2 * du
                fm = g(Xw, Uw)
                Uw[i, j] += du
                Jmv[:, i * nu + j] = (fp - fm) / (2 * du)
        for j in range(nu):
            du = DV * Ua[ph - 1, j]
            Uw[ph - 1, j] += du; Uw[ph, j] += du
            fp = g(Xw, Uw)
            Uw[ph - 1, j] -= 2 * du; Uw[ph, j] -= 2 * du
            fm = g(Xw, Uw)
            Uw[ph - 1, j] += du; Uw[ph, j] += du
            Jmv[:, (ph - 1) * nu + j] = (fp - fm) / (2 * du)
        return val, self._scale_state_cols(self._glue(Jx, Jmv, np.zeros(self.neq)))


# ---- the reference's example systems (examples/*.cpp), as data for the oracle and for the device functors -----------
def vanderpol_field(x, u, i=0):
    """examples/vanderpol_ex.cpp:36-43."""
    return np.array([((1.0 - x[1] * x[1]) * x[0]) - x[1] + u[0], x[0]])


def oscillator_network_field(N, mu=1.0, k=0.1):
    """examples/networked_oscillators_ex.cpp:17-32."""
    def f(x, u, i=0):
        dx = np.zeros(2 * N)
        for a in range(N):
            dx[2 * a] = x[2 * a + 1]
            dx[2 * a + 1] = mu * (1 - x[2 * a] * x[2 * a]) * x[2 * a + 1] - x[2 * a] + u[a]
            for b in range(N):
                if a != b:
                    dx[2 * a + 1] += k * (x[2 * b] - x[2 * a])
        return dx
    return f


def sum_squares_cost(X, Y, U, e):
    """x.array().square().sum() + u.array().square().sum()  (vanderpol_ex.cpp:52-57, networked_oscillators_ex.cpp:52-57)."""
    return float((X * X).sum() + (U * U).sum())


def vanderpol_formulation():
    """examples/vanderpol_ex.cpp:9-71: nx2 nu1 ny2 ph10 ch5, Tineq=11 (u_i <= 0.5), continuous Ts=0.1."""
    f = NLMPCFormulation(2, 1, 2, 10, 5, nineq=11)
    f.continuous, f.Ts = True, 0.1
    f.f = vanderpol_field
    f.obj = sum_squares_cost
    f.ineq = lambda X, Y, U, e: U[:, 0] - 0.5
    return f


def oscnet_formulation(N, ph, ch, Ts=0.1, mu=1.0, k=0.1):
    """examples/networked_oscillators_ex.cpp with N oscillators: nx=2N, nu=N, Tineq=(ph+1)*nu (u <= 0.5), continuous."""
    f = NLMPCFormulation(2 * N, N, 2 * N, ph, ch, nineq=(ph + 1) * N)
    f.continuous, f.Ts = True, Ts
    f.f = oscillator_network_field(N, mu, k)
    f.obj = sum_squares_cost
    f.ineq = lambda X, Y, U, e: (U - 0.5).ravel()      # idx = i*nu + j
    return f


def ugv_model(Ts=0.1, m=1.0):
    """Discretised double integrator of examples/ugv_ex.cpp:36-57 (c2d via the matrix exponential, Utils.hpp:23-47)."""
    from oracle.lmpc_formulation import discretization
    A = np.zeros((4, 4)); A[0:2, 2:4] = np.eye(2)
    B = np.zeros((4, 2)); B[2:4, :] = np.eye(2) / m
    return discretization(A, B, Ts)


def ugv_formulation(ph=10, ch=10, v_pref=(0.0, 0.0), obstacles=((2.0, 1.0, 0.3), (1.0, 1.0, 0.3))):
    """examples/ugv_ex.cpp:12-126: nx4 nu2 ny4, discrete, y = x, soft obstacle constraints."""
    Ad, Bd = ugv_model()
    obs = np.asarray(obstacles, float)
    vp = np.asarray(v_pref, float)
    f = NLMPCFormulation(4, 2, 4, ph, ch, nineq=(ph + 1) * len(obs))
    f.f = lambda x, u, i=0: Ad @ x + Bd @ u
    f.out = lambda x, u, i=0: x.copy()

    def cost(X, Y, U, e):
        c = 0.0
        for i in range(ph + 1):
            d = X[i, 2:4] - vp
            c += 1e3 * (d[0] * d[0] + d[1] * d[1])
            c += 1e-2 * (U[i, 0] * U[i, 0] + U[i, 1] * U[i, 1])
        return c + 1e-5 * e * e

    def ineq(X, Y, U, e):
        out = np.zeros((ph + 1) * len(obs))
        k = 0
        for i in range(ph + 1):
            for j in range(len(obs)):
                out[k] = obs[j, 2] - np.hypot(X[i, 0] - obs[j, 0], X[i, 1] - obs[j, 1])
                k += 1
        return out
    f.obj, f.ineq = cost, ineq
    f.params = np.concatenate([Ad.ravel(), Bd.ravel(), vp, obs.ravel()])
    return f
