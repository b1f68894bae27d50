// examples/vanderpol_ex.cpp written against the C++ mirror include/mpc_b200/NLMPC.hpp: same template sizes, parameters,
// bounds and closed loop; the three callbacks are replaced by setSystem(B200MPC_SYS_VANDERPOL, {Ts}).  The first command
// and cost are the numbers the restated example gives under the SLSQP oracle (oracle/nlmpc_slsqp.py: 0.09098442,
// 11.1952468).  Exit code: 0 pass, 2 mismatch, 3 no CUDA device.
#include <mpc_b200/NLMPC.hpp>

#include <cmath>
#include <cstdio>
#include <cstring>

int main() {
    constexpr int num_states = 2, num_output = 2, num_inputs = 1, pred_hor = 10, ctrl_hor = 5, ineq_c = pred_hor + 1, eq_c = 0;
    const double ts = 0.1;
    try {
        mpc::NLMPC<num_states, num_inputs, num_output, pred_hor, ctrl_hor, ineq_c, eq_c> controller;
        controller.setSystem(B200MPC_SYS_VANDERPOL, {ts});
        if (!controller.setDiscretizationSamplingTime(ts)) return 2;
        bool threw = false;
        try { controller.setObjectiveFunction([](int) { return 0.0; }); } catch (const std::runtime_error&) { threw = true; }
        if (!threw) return 2;

        mpc::NLParameters params;
        params.maximum_iteration = 100;
        params.relative_ftol = -1;
        params.relative_xtol = -1;
        params.hard_constraints = true;
        params.enable_warm_start = true;
        controller.setOptimizerParameters(params);

        mpc::cvec<num_states> modelX, modeldX;
        modelX(0) = 0; modelX(1) = 1.0;
        mpc::cvec<num_inputs> u;
        u.setZero();
        auto r = controller.optimize(modelX, u);
        std::printf("first cmd %.8f cost %.7f status %d feasible %d\n", r.cmd(0), r.cost, (int)r.status, (int)r.is_feasible);
        if (std::fabs(r.cmd(0) - 0.09098442) > 1e-6 || std::fabs(r.cost - 11.1952468) > 1e-6) return 2;
        if (r.status != mpc::ResultStatus::SUCCESS || !r.is_feasible) return 2;
        auto seq = controller.getOptimalSequence();
        if (seq.state.rows() != pred_hor + 1 || seq.input(0, 0) != r.cmd(0) || seq.state(0, 1) != 1.0) return 2;

        // closed loop as in the example: forward Euler on the Van der Pol field until the state is near the origin
        int steps = 0;
        for (; steps < 300; ++steps) {
            r = controller.optimize(modelX, u);
            if (r.status != mpc::ResultStatus::SUCCESS || r.cmd(0) > 0.5 + 1e-9) return 2;
            u = r.cmd;
            modeldX(0) = ((1.0 - modelX(1) * modelX(1)) * modelX(0)) - modelX(1) + u(0);
            modeldX(1) = modelX(0);
            modelX(0) += modeldX(0) * ts; modelX(1) += modeldX(1) * ts;
            if (std::fabs(modelX(0)) <= 1e-2 && std::fabs(modelX(1)) <= 1e-2) break;
        }
        std::printf("closed loop: %d steps, x = (%.5f, %.5f)\n", steps, modelX(0), modelX(1));
        if (steps >= 300) return 2;
        return 0;
    } catch (const std::exception& e) {
        std::printf("%s\n", e.what());
        return std::strstr(e.what(), "no CUDA device") ? 3 : 2;
    }
}
