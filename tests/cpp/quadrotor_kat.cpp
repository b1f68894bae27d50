// The reference's "LMPC interface test" (test/LMPC/test_common.cpp:89-237) written against the C++ mirror
// include/mpc_b200/LMPC.hpp: same calls, same golden vector.  Exit code: 0 pass, 2 mismatch, 3 no CUDA device.
#include <mpc_b200/LMPC.hpp>

#include <cmath>
#include <cstdio>
#include <cstring>

int main() {
    constexpr int Tnx = 12, Tny = 12, Tnu = 4, Tndu = 4, Tph = 10, Tch = 10;
    try {
        mpc::LMPC<Tnx, Tnu, Tndu, Tny, Tph, Tch> optsolver;
        mpc::mat<Tnx, Tnx> Ad;
        Ad.fillRowMajor({1, 0, 0, 0, 0, 0, 0.1, 0, 0, 0, 0, 0,
                         0, 1, 0, 0, 0, 0, 0, 0.1, 0, 0, 0, 0,
                         0, 0, 1, 0, 0, 0, 0, 0, 0.1, 0, 0, 0,
                         0.0488, 0, 0, 1, 0, 0, 0.0016, 0, 0, 0.0992, 0, 0,
                         0, -0.0488, 0, 0, 1, 0, 0, -0.0016, 0, 0, 0.0992, 0,
                         0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0.0992,
                         0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0,
                         0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0,
                         0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0,
                         0.9734, 0, 0, 0, 0, 0, 0.0488, 0, 0, 0.9846, 0, 0,
                         0, -0.9734, 0, 0, 0, 0, 0, -0.0488, 0, 0, 0.9846, 0,
                         0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.9846});
        mpc::mat<Tnx, Tnu> Bd;
        Bd.fillRowMajor({0, -0.0726, 0, 0.0726,
                         -0.0726, 0, 0.0726, 0,
                         -0.0152, 0.0152, -0.0152, 0.0152,
                         0, -0.0006, -0.0000, 0.0006,
                         0.0006, 0, -0.0006, 0,
                         0.0106, 0.0106, 0.0106, 0.0106,
                         0, -1.4512, 0, 1.4512,
                         -1.4512, 0, 1.4512, 0,
                         -0.3049, 0.3049, -0.3049, 0.3049,
                         0, -0.0236, 0, 0.0236,
                         0.0236, 0, -0.0236, 0,
                         0.2107, 0.2107, 0.2107, 0.2107});
        mpc::mat<Tny, Tnx> Cd;
        Cd.setIdentity();
        if (!optsolver.setStateSpaceModel(Ad, Bd, Cd)) return 2;
        if (!optsolver.setDisturbances(mpc::mat<Tnx, Tndu>::Zero(), mpc::mat<Tny, Tndu>::Zero())) return 2;

        mpc::cvec<Tnu> InputW, DeltaInputW;
        mpc::cvec<Tny> OutputW;
        OutputW.fillRowMajor({0, 0, 10, 10, 10, 10, 0, 0, 0, 5, 5, 5});
        InputW.setConstant(0.1);
        DeltaInputW.setZero();
        if (!optsolver.setObjectiveWeights(OutputW, InputW, DeltaInputW, {0, Tph})) return 2;

        mpc::cvec<Tnx> xmin, xmax;
        xmin.setConstant(-mpc::inf); xmax.setConstant(mpc::inf);
        xmin(0) = xmin(1) = -M_PI / 6; xmin(5) = -1;
        xmax(0) = xmax(1) = M_PI / 6;
        mpc::cvec<Tny> ymin, ymax;
        ymin.setConstant(-mpc::inf); ymax.setConstant(mpc::inf);
        mpc::cvec<Tnu> umin, umax;
        double u0 = 10.5916;
        umin.setConstant(9.6 - u0); umax.setConstant(13 - u0);
        if (!optsolver.setStateBounds(xmin, xmax, {0, Tph})) return 2;
        if (!optsolver.setInputBounds(umin, umax, {0, Tph})) return 2;
        if (!optsolver.setOutputBounds(ymin, ymax, {0, Tph})) return 2;
        mpc::cvec<Tnx> onesx; onesx.setOnes();
        mpc::cvec<Tnu> onesu; onesu.setOnes();
        if (!optsolver.setScalarConstraint(-mpc::inf, mpc::inf, onesx, onesu, {-1, -1})) return 2;
        if (!optsolver.setScalarConstraint(0, -mpc::inf, mpc::inf, onesx, onesu)) return 2;

        mpc::cvec<Tny> yRef; yRef.setZero(); yRef(2) = 1;
        if (!optsolver.setReferences(yRef, mpc::cvec<Tnu>::Zero(), mpc::cvec<Tnu>::Zero(), {0, Tph})) return 2;
        mpc::LParameters params;
        params.maximum_iteration = 250;
        optsolver.setOptimizerParameters(params);
        if (!optsolver.setExogenousInputs(mpc::cvec<Tndu>::Zero(), {0, Tph})) return 2;

        // slice validation and unsupported calls behave like the reference
        if (optsolver.setStateBounds(xmin, xmax, {3, 2})) return 2;
        bool threw = false;
        try { optsolver.setDiscretizationSamplingTime(0.1); } catch (const std::runtime_error&) { threw = true; }
        if (!threw) return 2;

        auto res = optsolver.step(mpc::cvec<Tnx>::Zero(), mpc::cvec<Tnu>::Zero());
        auto seq = optsolver.getOptimalSequence();
        const double golden[4] = {-0.9916, 1.74839, -0.9916, 1.74839};
        double num = 0, den = 0;
        for (int k = 0; k < 4; ++k) { num += (res.cmd(k) - golden[k]) * (res.cmd(k) - golden[k]); den += golden[k] * golden[k]; }
        std::printf("cmd = [%.8f %.8f %.8f %.8f] status=%d solver_status=%d cost=%.9g seq.input(0,1)=%.8f\n", res.cmd(0), res.cmd(1),
                    res.cmd(2), res.cmd(3), (int)res.status, res.solver_status, res.cost, seq.input(0, 1));
        if (std::sqrt(num) > 1e-4 * std::sqrt(den)) return 2;      // Eigen isApprox(…, 1e-4)
        if (res.status != mpc::SUCCESS || !res.is_feasible) return 2;
        if (std::fabs(seq.input(0, 1) - res.cmd(1)) > 1e-15) return 2;

        // the loop of examples/quadrotor_ex.cpp on the device: the first command of the trajectory is the golden one, and the
        // plant follows x+ = A x + B u
        std::vector<double> tx, tu, xinit(Tnx, 0.0), uinit(Tnu, 0.0);
        optsolver.closedLoop(xinit.data(), uinit.data(), 3, tx, tu);
        for (int k = 0; k < 4; ++k) if (std::fabs(tu[k] - res.cmd(k)) > 1e-9) return 2;
        double x1_2 = 0;                                            // row 2 of A x0 + B u0 with x0 = 0
        for (int k = 0; k < 4; ++k) x1_2 += Bd(2, k) * tu[k];
        if (std::fabs(tx[Tnx + 2] - x1_2) > 1e-12) return 2;

        // mpc::discretization (test/test_utils.cpp:10-63): 2-dof double integrator, Ts = 0.02
        mpc::mat<4, 4> Ac, Adz; mpc::mat<4, 2> Bc, Bdz;
        Ac(0, 2) = 1; Ac(1, 3) = 1; Bc(2, 0) = 1; Bc(3, 1) = 1;
        mpc::discretization<4, 2>(Ac, Bc, 0.02, Adz, Bdz);
        if (std::fabs(Adz(0, 2) - 0.02) > 1e-15 || std::fabs(Adz(0, 0) - 1) > 1e-15 || std::fabs(Bdz(0, 0) - 0.0002) > 1e-15 ||
            std::fabs(Bdz(2, 0) - 0.02) > 1e-15) return 2;
        return 0;
    } catch (const std::exception& e) {
        std::printf("exception: %s\n", e.what());
        return std::strstr(e.what(), "no CUDA device") ? 3 : 2;
    }
}
