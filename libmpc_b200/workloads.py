"""Synthetic workloads of BASELINE.json `configs`, shared by bench.py, tools/ and tests/.

Only data: model constants of the reference's example programs, the seeded per-instance inputs, and -- for the one
BASELINE shape the reference does not ship (configs[2], unicycle nx=3 nu=2 Tph=30) -- the CUDA source of the user-defined
system.  Nothing here computes a solve; the oracle (oracle/) does not import this and this does not import the oracle.
RNG convention of SURVEY 8d: instance b is drawn from numpy PCG64(seed + b), so a shard is reproducible on any GPU count.
"""
import numpy as np

INF = float("inf")

# ---- configs[1] / configs[4]: quadrotor LMPC (examples/quadrotor_ex.cpp:19-83, data constants) ----------------------
QUAD_NX, QUAD_NU, QUAD_NDU, QUAD_NY = 12, 4, 4, 12
QUAD_X0_SCALE = np.array([0.2, 0.2, 0.5, 0.5, 0.5, 0.5] + [0.3] * 6)


def quadrotor_model():
    """Ad, Bd of examples/quadrotor_ex.cpp:19-45."""
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


def quadrotor_setup():
    """Weights and bounds of examples/quadrotor_ex.cpp:47-83 / test/LMPC/test_common.cpp:150-215 (vectors, one stage)."""
    w_out = np.array([0, 0, 10, 10, 10, 10, 0, 0, 0, 5, 5, 5.0])
    w_u = np.full(4, 0.1)
    w_du = np.zeros(4)
    xmin = np.full(12, -INF); xmax = np.full(12, INF)
    xmin[0] = xmin[1] = -np.pi / 6; xmax[0] = xmax[1] = np.pi / 6
    xmin[5] = -1
    u_eq = 10.5916
    umin = np.full(4, 9.6) - u_eq; umax = np.full(4, 13.0) - u_eq
    return dict(w_out=w_out, w_u=w_u, w_du=w_du, xmin=xmin, xmax=xmax, umin=umin, umax=umax)


def quadrotor_inputs(first, count, seed=20):
    """x0 ~ U(-1,1)*scale clipped into the state box, u0 = 0, yRef = [0,0,r,0..], r ~ U(0.5,1.5)."""
    x0 = np.empty((count, QUAD_NX))
    r = np.empty(count)
    for k in range(count):
        g = np.random.Generator(np.random.PCG64(seed + first + k))
        x0[k] = g.uniform(-1, 1, QUAD_NX) * QUAD_X0_SCALE
        r[k] = g.uniform(0.5, 1.5)
    x0[:, 0:2] = np.clip(x0[:, 0:2], -np.pi / 6, np.pi / 6)
    x0[:, 5] = np.maximum(x0[:, 5], -1.0)
    return x0, r


def build_quadrotor_controller(L, ph, batch, max_iter, per_instance_model=False, device=0):
    """mpc::LMPC<12,4,4,12,ph,ph> of quadrotor_ex.cpp for `batch` controllers through the Python mirror `L`."""
    c = L.LMPC(QUAD_NX, QUAD_NU, QUAD_NDU, QUAD_NY, ph, ph, batch=batch, device=device)
    Ad, Bd = quadrotor_model()
    s = quadrotor_setup()
    if per_instance_model:
        c.setStateSpaceModel(np.broadcast_to(Ad, (batch, 12, 12)), np.broadcast_to(Bd, (batch, 12, 4)),
                             np.broadcast_to(np.eye(12), (batch, 12, 12)))
    else:
        c.setStateSpaceModel(Ad, Bd, np.eye(12))
    c.setObjectiveWeights(s["w_out"], s["w_u"], s["w_du"], (0, ph))
    c.setStateBounds(s["xmin"], s["xmax"], (0, ph))
    c.setInputBounds(s["umin"], s["umax"], (0, ph))
    c.setOptimizerParameters(L.LParameters(maximum_iteration=max_iter))
    return c


# ---- configs[2]: unicycle NLMPC nx=3 nu=2 Tph=Tch=30 with two obstacles -----------------------------
# x = [px, py, theta], u = [v, omega], discrete x+ = x + Ts [v cos(theta), v sin(theta), omega], two circular obstacles on
# (px, py) -> Tineq = 62, soft constraints; cost 10 |p - p_goal|^2 + 1e-2 |u|^2 + 1e-5 e^2.  The model is NOT in the reference
# (its ugv_ex is a double integrator with nx = 4): the shape exists only as a user-defined system (NVRTC), and parity is
# against the restated oracle only.  params = [Ts, goal(2), obs0(x,y,r), obs1(x,y,r)].
UNICYCLE_SRC = r"""
struct UserUnicycle {
    static constexpr int nx = 3, nu = 2, ny = 3, nparam = 9, nobs = 2;
    static constexpr bool continuous = false;
    static constexpr int ineq_per_stage = nobs;
    __device__ static double Ts(const double*) { return 0.0; }
    __host__ __device__ static int nineq(int ph) { return (ph + 1) * nobs; }
    __device__ static void f(double* xn, const double* x, const double* u, int, const double* p) {
        xn[0] = x[0] + p[0] * (u[0] * cos(x[2]));
        xn[1] = x[1] + p[0] * (u[0] * sin(x[2]));
        xn[2] = x[2] + p[0] * u[1];
    }
    __device__ static double cost(const Acc& a, double e, int ph, const double* p) {
        double c = 0;
        for (int i = 0; i <= ph; ++i) {
            double d0 = a.x(i, 0) - p[1], d1 = a.x(i, 1) - p[2], u0 = a.u(i, 0), u1 = a.u(i, 1);
            c += 1e1 * (d0 * d0 + d1 * d1);
            c += 1e-2 * (u0 * u0 + u1 * u1);
        }
        return c + 1e-5 * e * e;
    }
    __device__ static double ineq(int r, const Acc& a, double, int, const double* p) {
        int i = r / nobs, j = r % nobs;
        double dx = a.x(i, 0) - p[3 + 3 * j], dy = a.x(i, 1) - p[3 + 3 * j + 1];
        return p[3 + 3 * j + 2] - sqrt(dx * dx + dy * dy);
    }
};
"""
UNICYCLE_TYPE = "UserUnicycle"
UNICYCLE_OBSTACLES = ((1.0, 0.6, 0.3), (1.4, 1.7, 0.3))


def unicycle_inputs(first, count, seed=30, Ts=0.1, goal=(2.0, 2.0)):
    """Per instance: start ~ U([-0.5,0.5]^2) with heading ~ U(-0.3,0.3), goal (2,2), obstacle centres jittered +-0.1.
    Returns x0 [count,3], params [count,9]."""
    x0 = np.empty((count, 3)); params = np.empty((count, 9))
    for k in range(count):
        g = np.random.Generator(np.random.PCG64(seed + first + k))
        x0[k, :2] = g.uniform(-0.5, 0.5, 2)
        x0[k, 2] = g.uniform(-0.3, 0.3)
        obs = np.array(UNICYCLE_OBSTACLES)
        obs[:, :2] += g.uniform(-0.1, 0.1, (2, 2))
        params[k] = np.concatenate([[Ts], goal, obs.ravel()])
    return x0, params


# ---- configs[3]: networked oscillators N=4 (nx=8, nu=4), Tph=15, Tch=8 --------------------------------
def oscnet4_inputs(first, count, seed=40):
    """x0 ~ U(-1,1)^8; params [Ts, mu, k] = [0.1, 1, 0.1] (examples/networked_oscillators_ex.cpp:17-32)."""
    x0 = np.empty((count, 8))
    for k in range(count):
        g = np.random.Generator(np.random.PCG64(seed + first + k))
        x0[k] = g.uniform(-1, 1, 8)
    return x0, np.array([0.1, 1.0, 0.1])


def cold_start(x0, u0, ph, ch):
    """NLOptimizer::run cold initial guess (NLOptimizer.hpp:431-451): X_i = x0, U_i = u0, slack 0."""
    x0 = np.atleast_2d(x0); B = x0.shape[0]
    u0 = np.broadcast_to(np.atleast_2d(u0), (B, np.atleast_2d(u0).shape[1]))
    return np.concatenate([np.tile(x0, (1, ph)), np.tile(u0, (1, ch)), np.zeros((B, 1))], axis=1)


FLT_INF = float(np.float32(np.inf))


def soft_bounds(nz):
    """lb/ub default to +-float infinity (NLOptimizer.hpp:70-73); soft constraints: slack in [0, inf)."""
    lb = np.full(nz, -FLT_INF); ub = np.full(nz, FLT_INF)
    lb[-1] = 0.0
    return lb, ub


def hard_bounds(nz):
    lb = np.full(nz, -FLT_INF); ub = np.full(nz, FLT_INF)
    lb[-1] = ub[-1] = 0.0
    return lb, ub
