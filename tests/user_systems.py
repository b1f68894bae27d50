"""CUDA sources of user-defined NLMPC systems for the tests (the device-side counterpart of the std::function callbacks
the reference's NLMPC setters take; contract in include/b200mpc.h), each with the matching oracle formulation."""
import numpy as np

from oracle.nlmpc_formulation import NLMPCFormulation, sum_squares_cost, vanderpol_field

# examples/vanderpol_ex.cpp written as a user system: must reproduce the built-in SYS_VANDERPOL bit for bit.
VANDERPOL_SRC = r"""
struct UserVanDerPol {
    static constexpr int nx = 2, nu = 1, ny = 2, nparam = 1;
    static constexpr bool continuous = true;
    __device__ static double Ts(const double* p) { return p[0]; }
    __host__ __device__ static int nineq(int ph) { return ph + 1; }
    __device__ static void f(double* dx, const double* x, const double* u, int, const double*) {
        dx[0] = ((1.0 - (x[1] * x[1])) * x[0]) - x[1] + u[0];
        dx[1] = x[0];
    }
    __device__ static double cost(const Acc& a, double, int ph, const double*) {
        double sx = 0, su = 0;
        for (int j = 0; j < nx; ++j) for (int i = 0; i <= ph; ++i) { double v = a.x(i, j); sx += v * v; }
        for (int j = 0; j < nu; ++j) for (int i = 0; i <= ph; ++i) { double v = a.u(i, j); su += v * v; }
        return sx + su;
    }
    __device__ static double ineq(int r, const Acc& a, double, int, const double*) { return a.u(r, 0) - 0.5; }
};
"""

# A system the reference does not ship: damped pendulum on a cart-less pivot, discrete (explicit Euler), quadratic
# tracking cost, input limit as inequality, and a USER EQUALITY constraint tying the terminal state to the upright
# position through params (exercises NLMPC::setEqConFunction, NLMPC.hpp:261-281).  params = [Ts, damping, target angle].
PENDULUM_SRC = r"""
struct UserPendulum {
    static constexpr int nx = 2, nu = 1, ny = 2, nparam = 3;
    static constexpr bool continuous = false;
    __device__ static double Ts(const double*) { return 0.0; }
    __host__ __device__ static int nineq(int ph) { return 2 * (ph + 1); }
    __host__ __device__ static int neq(int) { return 2; }
    __device__ static void f(double* xn, const double* x, const double* u, int, const double* p) {
        xn[0] = x[0] + p[0] * x[1];
        xn[1] = x[1] + p[0] * (-9.81 * sin(x[0]) - p[1] * x[1] + u[0]);
    }
    __device__ static double cost(const Acc& a, double e, int ph, const double* p) {
        double c = 0;
        for (int i = 0; i <= ph; ++i) {
            double d0 = a.x(i, 0) - p[2], d1 = a.x(i, 1), u0 = a.u(i, 0);
            c += 10.0 * d0 * d0 + d1 * d1 + 0.1 * u0 * u0;
        }
        return c + 1e-3 * e * e;
    }
    __device__ static double ineq(int r, const Acc& a, double, int ph, const double*) {
        int i = r / 2;
        return (r & 1) ? (-a.u(i, 0) - 8.0) : (a.u(i, 0) - 8.0);
    }
    __device__ static double eq(int r, const Acc& a, int ph, const double* p) {
        return r == 0 ? (a.x(ph, 0) - p[2]) + 0.5 * a.x(ph, 1) : a.x(ph / 2, 0) * a.x(ph / 2, 0) + a.x(ph / 2, 1) - 0.1;
    }
};
"""

# A system with an OUTPUT MAP (NLMPC::setOutputFunction): the cost and the constraint read y = out(x, u), not x.
OUTPUT_MAP_SRC = r"""
struct UserWithOutput {
    static constexpr int nx = 2, nu = 1, ny = 1, nparam = 1;
    static constexpr bool continuous = false;
    __device__ static double Ts(const double*) { return 0.0; }
    __host__ __device__ static int nineq(int ph) { return ph + 1; }
    __device__ static void f(double* xn, const double* x, const double* u, int, const double* p) {
        xn[0] = x[0] + p[0] * x[1];
        xn[1] = x[1] + p[0] * u[0];
    }
    __device__ static void out(double* y, const double* x, const double* u, int, const double*) { y[0] = x[0] + 0.5 * x[1] * x[1]; }
    __device__ static double cost(const Acc& a, double, int ph, const double* p) {
        double c = 0;
        for (int i = 0; i <= ph; ++i) { double y = b200mpc::nl_y<UserWithOutput>(a, i, 0, p); c += (y - 1.0) * (y - 1.0) + 0.01 * a.u(i, 0) * a.u(i, 0); }
        return c;
    }
    __device__ static double ineq(int r, const Acc& a, double, int, const double* p) { return b200mpc::nl_y<UserWithOutput>(a, r, 0, p) - 1.5; }
};
"""

# BASELINE.json configs[2] at its stated shape: the CUDA source lives with the workload definitions
# (bench.py measures it); the oracle-side formulation is unicycle_formulation below.
from libmpc_b200.workloads import UNICYCLE_SRC  # noqa: E402,F401

BROKEN_SRC = "struct Broken { static constexpr int nx = 2; int oops( };"


def vanderpol_user_formulation():
    f = NLMPCFormulation(2, 1, 2, 10, 5, nineq=11)
    f.continuous, f.Ts, f.f, f.obj = True, 0.1, vanderpol_field, sum_squares_cost
    f.ineq = lambda X, Y, U, e: U[:, 0] - 0.5
    f.params = np.array([0.1])
    return f


def pendulum_formulation(ph=10, ch=5, Ts=0.1, damping=0.3, target=0.4):
    f = NLMPCFormulation(2, 1, 2, ph, ch, nineq=2 * (ph + 1), neq=2)
    f.continuous = False
    f.f = lambda x, u, i=0: np.array([x[0] + Ts * x[1], x[1] + Ts * (-9.81 * np.sin(x[0]) - damping * x[1] + u[0])])

    def cost(X, Y, U, e):
        c = 0.0
        for i in range(ph + 1):
            d0 = X[i, 0] - target
            c += 10.0 * d0 * d0 + X[i, 1] * X[i, 1] + 0.1 * U[i, 0] * U[i, 0]
        return c + 1e-3 * e * e
    f.obj = cost
    f.ineq = lambda X, Y, U, e: np.stack([U[:, 0] - 8.0, -U[:, 0] - 8.0], axis=1).ravel()
    f.eq = lambda X, U: np.array([(X[ph, 0] - target) + 0.5 * X[ph, 1], X[ph // 2, 0] * X[ph // 2, 0] + X[ph // 2, 1] - 0.1])
    f.params = np.array([Ts, damping, target])
    return f


def output_map_formulation(ph=6, ch=3, Ts=0.1):
    f = NLMPCFormulation(2, 1, 1, ph, ch, nineq=ph + 1)
    f.continuous = False
    f.f = lambda x, u, i=0: np.array([x[0] + Ts * x[1], x[1] + Ts * u[0]])
    f.out = lambda x, u, i=0: np.array([x[0] + 0.5 * x[1] * x[1]])
    f.obj = lambda X, Y, U, e: float(((Y[:, 0] - 1.0) ** 2).sum() + 0.01 * (U[:, 0] ** 2).sum())
    f.ineq = lambda X, Y, U, e: Y[:, 0] - 1.5
    f.params = np.array([Ts])
    return f


def unicycle_formulation(ph=30, ch=30, Ts=0.1, goal=(2.0, 2.0), obstacles=((1.0, 0.6, 0.3), (1.4, 1.7, 0.3)), params=None):
    """Oracle side of UNICYCLE_SRC (BASELINE.json configs[2] shape).  `params` = one row of
    libmpc_b200.workloads.unicycle_inputs ([Ts, goal(2), obs0(3), obs1(3)]) overrides the keyword values."""
    if params is not None:
        params = np.asarray(params, float)
        Ts, goal, obstacles = float(params[0]), params[1:3], params[3:9].reshape(2, 3)
    obs = np.asarray(obstacles, float); goal = np.asarray(goal, float)
    f = NLMPCFormulation(3, 2, 3, ph, ch, nineq=(ph + 1) * len(obs))
    f.continuous = False
    f.f = lambda x, u, i=0: np.array([x[0] + Ts * (u[0] * np.cos(x[2])), x[1] + Ts * (u[0] * np.sin(x[2])), x[2] + Ts * u[1]])

    def cost(X, Y, U, e):
        c = 0.0
        for i in range(ph + 1):
            d0, d1 = X[i, 0] - goal[0], X[i, 1] - goal[1]
            c += 1e1 * (d0 * d0 + d1 * d1)
            c += 1e-2 * (U[i, 0] * U[i, 0] + U[i, 1] * U[i, 1])
        return c + 1e-5 * e * e

    def ineq(X, Y, U, e):
        out = np.zeros((ph + 1) * len(obs))
        for i in range(ph + 1):
            for j in range(len(obs)):
                out[i * len(obs) + j] = obs[j, 2] - np.sqrt((X[i, 0] - obs[j, 0]) ** 2 + (X[i, 1] - obs[j, 1]) ** 2)
        return out
    f.obj, f.ineq = cost, ineq
    f.params = np.concatenate([[Ts], goal, obs.ravel()])
    return f
