// Host emulation of the stage-structured NLMPC solver (libmpc_b200/csrc/nlmpc_structured.cuh) -- TEST INFRASTRUCTURE.
// The device routines are cooperative loops over a thread group; compiled by g++ with a group of one thread they run
// sequentially, so tests/test_nlmpc_structured_host.py can check the exact kernel source against the Python specification
// (tests/nlmpc_sqp_reference.py with stage_groups + tests/nlmpc_structured_kkt_reference.py) and the SLSQP oracle on the CPU.
// USER_SYSTEM_SOURCE / USER_SYSTEM_TYPE (optional, -D / -include) add one user-defined system (e.g. the unicycle of
// BASELINE configs[2]) as system id 100.
#include <nlmpc_structured.cuh>

#include <vector>

using namespace b200mpc;
#ifdef USER_SYSTEM_HEADER
#include USER_SYSTEM_HEADER
#endif

template <class S>
static int run(int ph, int ch, const double* z0, const double* x0, const double* params, const double* lb, const double* ub, int max_sqp,
               int max_qp, double* z, double* out) {
    if (!nls_supported<S>(ph, ch)) return -2;
    constexpr int K = NlIneqPerStage<S>::value;
    std::vector<double> mem(nls_doubles(ph, ch, S::nx, S::nu, K) + 64, 0.0);
    NlSW<S::nx, S::nu, K> w;
    w.carve(mem.data(), ph, ch);
    NlSParams a{max_sqp, max_qp, 1e-7, 1e-12, 1e-5, 0.1, lb, ub, nullptr, nullptr};
    NlGrpHost g;
    NlSResult r = nls_solve_instance<S>(g, w, a, z0, x0, params, z);
    out[0] = r.cost; out[1] = r.viol; out[2] = r.status; out[3] = r.iters; out[4] = r.qp_iters;
    return 0;
}

extern "C" int nls_host_solve(int system, int ph, int ch, const double* z0, const double* x0, const double* params, const double* lb,
                              const double* ub, int max_sqp, int max_qp, double* z, double* out) {
    switch (system) {
    case 0: return run<SysVanDerPol>(ph, ch, z0, x0, params, lb, ub, max_sqp, max_qp, z, out);
    case 1: return run<SysOscNet<4>>(ph, ch, z0, x0, params, lb, ub, max_sqp, max_qp, z, out);
    case 3: return run<SysUgv>(ph, ch, z0, x0, params, lb, ub, max_sqp, max_qp, z, out);
#ifdef USER_SYSTEM_HEADER
    case 100: return run<USER_SYSTEM_TYPE>(ph, ch, z0, x0, params, lb, ub, max_sqp, max_qp, z, out);
#endif
    default: return -1;
    }
}
