// A controller whose model, cost and constraints are the USER's own -- what the reference does with
// NLMPC::setStateSpaceFunction / setObjectiveFunction / setIneqConFunction / setEqConFunction (NLMPC.hpp:139-281) --
// written against the C++ mirror include/mpc_b200/NLMPC.hpp: the callbacks become CUDA source handed to setSystemSource,
// compiled by the engine (NVRTC) into its solver kernels.  Damped pendulum, explicit Euler, 22 input-limit inequalities,
// 2 user equality constraints (Teq = 2).  Golden numbers: the same problem under the SLSQP oracle (tests/user_systems.py
// pendulum_formulation + oracle/nlmpc_slsqp.py): x0 = (0.1, 0) -> first command 2.17558, cost 11.42719517.
// Exit code: 0 pass, 2 mismatch, 3 no CUDA device.
#include <mpc_b200/NLMPC.hpp>

#include <cmath>
#include <cstdio>
#include <cstring>

static const char* kPendulum = R"CUDA(
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
)CUDA";

int main() {
    constexpr int nx = 2, ny = 2, nu = 1, ph = 10, ch = 5, ineq_c = 2 * (ph + 1), eq_c = 2;
    try {
        if (b200mpc_device_count() <= 0) { std::printf("no CUDA device\n"); return 3; }
        mpc::NLMPC<nx, nu, ny, ph, ch, ineq_c, eq_c> controller;
        controller.setSystemSource(kPendulum, "UserPendulum", {0.1, 0.3, 0.4});
        mpc::NLParameters params;
        params.maximum_iteration = 100;
        params.hard_constraints = true;
        controller.setOptimizerParameters(params);
        controller.setEqTolerance(1e-7);
        mpc::cvec<nx> x;
        x(0) = 0.1; x(1) = 0.0;
        mpc::cvec<nu> u;
        u.setZero();
        auto r = controller.optimize(x, u);
        std::printf("first cmd %.8f cost %.8f status %d feasible %d\n", r.cmd(0), r.cost, (int)r.status, (int)r.is_feasible);
        if (std::fabs(r.cmd(0) - 2.17558) > 3e-5 || std::fabs(r.cost - 11.42719517) > 1e-6) return 2;
        if (r.status != mpc::ResultStatus::SUCCESS || !r.is_feasible) return 2;
        // the terminal equality holds on the returned sequence
        auto seq = controller.getOptimalSequence();
        if (std::fabs((seq.state(ph, 0) - 0.4) + 0.5 * seq.state(ph, 1)) > 1e-7) return 2;
        // a second controller with a wrong Teq is rejected
        bool threw = false;
        try { mpc::NLMPC<nx, nu, ny, ph, ch, ineq_c, 0> bad; bad.setSystemSource(kPendulum, "UserPendulum", {0.1, 0.3, 0.4}); }
        catch (const std::runtime_error&) { threw = true; }
        return threw ? 0 : 2;
    } catch (const std::exception& e) {
        std::printf("%s\n", e.what());
        return std::strstr(e.what(), "no CUDA device") ? 3 : 2;
    }
}
