// B200Optimizer.hpp -- the reference-side binding of the B200 engine: an mpc::IOptimizer<sizer> (include/mpc/IOptimizer.hpp:24-58)
// that a libmpc++ maintainer drops next to LOptimizer.hpp.  It compiles AGAINST THE REFERENCE'S OWN HEADERS (<mpc/IOptimizer.hpp>,
// Eigen types) and forwards every call below the seam to the C ABI in b200mpc.h; LMPC::onSetup (LMPC.hpp:728-735) would create
// it instead of LOptimizer + ProblemBuilder:
//     optPtr = new B200Optimizer<sizer>();        // and the LMPC<> setters call the same-named members below
// The setter names / argument types are those of ProblemBuilder (ProblemBuilder.hpp:184-504) and LOptimizer
// (LOptimizer.hpp:100-186) -- the two classes LMPC<> reaches by downcast (LMPC.hpp:537-721) -- so the front-end changes by one
// type name.  Memory layout: the reference's column-major mat<dim,ph> is byte-for-byte the stage-major [ph][dim] block the C ABI
// takes; model matrices (row-major in the ABI) go through a transposed temporary.
//
// tests/test_b200optimizer_header.py compiles this header against /root/reference/include with a minimal Eigen stand-in
// (tests/cpp/eigen_stub; Eigen itself is not in the build image) -- it cannot be linked into the reference here (no OSQP /
// NLopt / Eigen), which is why include/mpc_b200/LMPC.hpp exists as the self-contained mirror.
#pragma once
#include <mpc/IDimensionable.hpp>
#include <mpc/IOptimizer.hpp>

#include <stdexcept>
#include <vector>

#include "../b200mpc.h"

namespace mpc {

template <MPCSize sizer>
class B200Optimizer : public IOptimizer<sizer> {
    using IComponent<sizer>::checkOrQuit;
    using IDimensionable<sizer>::nx;
    using IDimensionable<sizer>::nu;
    using IDimensionable<sizer>::ndu;
    using IDimensionable<sizer>::ny;
    using IDimensionable<sizer>::ph;
    using IDimensionable<sizer>::ch;

public:
    using IOptimizer<sizer>::result;
    using IOptimizer<sizer>::sequence;

    B200Optimizer() = default;
    ~B200Optimizer() override { if (h) b200mpc_lmpc_destroy(h); }
    B200Optimizer(const B200Optimizer&) = delete;
    B200Optimizer& operator=(const B200Optimizer&) = delete;

    // LOptimizer::onInit (LOptimizer.hpp:59-82) + ProblemBuilder::onInit (ProblemBuilder.hpp:88-172)
    void onInit() override {
        b200mpc_lmpc_dims d{(int)nx(), (int)nu(), (int)ndu(), (int)ny(), (int)ph(), (int)ch()};
        if (b200mpc_lmpc_create(&d, /*batch=*/1, /*device=*/0, &h) != B200MPC_OK) throw std::runtime_error(b200mpc_last_error());
        COND_RESIZE_MAT(sizer, sequence.state, ph() + 1, nx());
        COND_RESIZE_MAT(sizer, sequence.input, ph() + 1, nu());
        COND_RESIZE_MAT(sizer, sequence.output, ph() + 1, ny());
        COND_RESIZE_CVEC(sizer, result.cmd, nu());
        sequence.state.setZero(); sequence.input.setZero(); sequence.output.setZero(); result.cmd.setZero();
        rs.assign((size_t)(ph() + 1) * nx(), 0.0); ri.assign((size_t)(ph() + 1) * nu(), 0.0); ro.assign((size_t)(ph() + 1) * ny(), 0.0);
    }

    // LOptimizer::setParameters (LOptimizer.hpp:100-108)
    void setParameters(const Parameters& param) override {
        checkOrQuit();
        const LParameters& lp = *dynamic_cast<const LParameters*>(&param);
        b200mpc_lmpc_params q;
        b200mpc_lmpc_default_params(&q);
        q.maximum_iteration = lp.maximum_iteration; q.enable_warm_start = lp.enable_warm_start ? 1 : 0; q.time_limit = lp.time_limit;
        q.alpha = lp.alpha; q.rho = lp.rho; q.eps_rel = lp.eps_rel; q.eps_abs = lp.eps_abs;
        q.eps_prim_inf = lp.eps_prim_inf; q.eps_dual_inf = lp.eps_dual_inf;
        q.adaptive_rho = lp.adaptive_rho ? 1 : 0; q.polish = lp.polish ? 1 : 0;
        ok(b200mpc_lmpc_set_params(h, &q));
    }

    // ---- what LMPC<> forwards to ProblemBuilder (same names / argument types) ----
    bool setStateModel(const mat<sizer.nx, sizer.nx>& A, const mat<sizer.nx, sizer.nu>& B, const mat<sizer.ny, sizer.nx>& C) {
        checkOrQuit();
        std::vector<double> a = rowmajor(A), b = rowmajor(B), c = rowmajor(C);
        return b200mpc_lmpc_set_model(h, a.data(), b.data(), c.data(), 0, 0) == B200MPC_OK;
    }
    bool setExogenousInput(const mat<sizer.nx, sizer.ndu>& Bd, const mat<sizer.ny, sizer.ndu>& Dd) {
        checkOrQuit();
        std::vector<double> b = rowmajor(Bd), d = rowmajor(Dd);
        return b200mpc_lmpc_set_disturbances(h, b.data(), d.data(), 0, 0) == B200MPC_OK;
    }
    bool setObjective(const mat<sizer.ny, sizer.ph>& OWeight, const mat<sizer.nu, sizer.ph>& UWeight,
                      const mat<sizer.nu, sizer.ph>& DeltaUWeight) {
        checkOrQuit();
        return b200mpc_lmpc_set_weights(h, OWeight.data(), UWeight.data(), DeltaUWeight.data(), 0, 0) == B200MPC_OK;
    }
    bool setScalarConstraint(const cvec<sizer.ph>& MinMat, const cvec<sizer.ph>& MaxMat, const cvec<sizer.nx>& X,
                             const cvec<sizer.nu>& U) {
        checkOrQuit();
        return b200mpc_lmpc_set_scalar_constraint(h, MinMat.data(), MaxMat.data(), X.data(), U.data(), 0, 0) == B200MPC_OK;
    }
    bool setStateBounds(const mat<sizer.nx, sizer.ph> XMinMat, const mat<sizer.nx, sizer.ph> XMaxMat) {
        checkOrQuit();
        return b200mpc_lmpc_set_state_bounds(h, XMinMat.data(), XMaxMat.data(), 0, 0) == B200MPC_OK;
    }
    bool setInputBounds(const mat<sizer.nu, sizer.ch> UMinMat, const mat<sizer.nu, sizer.ch> UMaxMat) {
        checkOrQuit();
        return b200mpc_lmpc_set_input_bounds(h, UMinMat.data(), UMaxMat.data(), 0, 0) == B200MPC_OK;
    }
    bool setOutputBounds(const mat<sizer.ny, sizer.ph> YMinMat, const mat<sizer.ny, sizer.ph> YMaxMat) {
        checkOrQuit();
        return b200mpc_lmpc_set_output_bounds(h, YMinMat.data(), YMaxMat.data(), 0, 0) == B200MPC_OK;
    }
    // ---- what LMPC<> forwards to LOptimizer ----
    bool setReferences(const mat<sizer.ny, sizer.ph>& outRef, const mat<sizer.nu, sizer.ph>& cmdRef,
                       const mat<sizer.nu, sizer.ph>& deltaCmdRef) {
        return b200mpc_lmpc_set_references(h, outRef.data(), cmdRef.data(), deltaCmdRef.data(), 0, 0) == B200MPC_OK;
    }
    bool setExogenousInputs(const mat<sizer.ndu, sizer.ph>& uMeas) {
        return b200mpc_lmpc_set_exogenous_inputs(h, uMeas.data(), 0, 0) == B200MPC_OK;
    }

    // LOptimizer::run (LOptimizer.hpp:189-368): solve, then Result / OptSequence exactly as :292-361 fills them
    void run(const cvec<sizer.nx>& x0, const cvec<sizer.nu>& u0) override {
        checkOrQuit();
        ok(b200mpc_lmpc_solve(h, x0.data(), u0.data(), 0));
        int32_t st = 0, sst = 0, feas = 0;
        double cost = 0;
        ok(b200mpc_lmpc_get_result(h, result.cmd.data(), &cost, &st, &sst, &feas, nullptr, nullptr, nullptr, 0));
        result.cost = cost;
        result.status = (ResultStatus)st;
        result.solver_status = sst;
        result.is_feasible = feas != 0;
        ok(b200mpc_lmpc_get_sequence(h, rs.data(), ri.data(), ro.data(), 0));       // row-major [(ph+1) x dim]
        for (int i = 0; i <= (int)ph(); ++i) {
            for (int j = 0; j < (int)nx(); ++j) sequence.state(i, j) = rs[(size_t)i * nx() + j];
            for (int j = 0; j < (int)nu(); ++j) sequence.input(i, j) = ri[(size_t)i * nu() + j];
            for (int j = 0; j < (int)ny(); ++j) sequence.output(i, j) = ro[(size_t)i * ny() + j];
        }
    }

    b200mpc_lmpc_t handle() const { return h; }

private:
    template <class M>
    static std::vector<double> rowmajor(const M& m) {
        std::vector<double> v((size_t)m.rows() * m.cols());
        for (int r = 0; r < (int)m.rows(); ++r)
            for (int c = 0; c < (int)m.cols(); ++c) v[(size_t)r * m.cols() + c] = m(r, c);
        return v;
    }
    static void ok(int rc) { if (rc != B200MPC_OK) throw std::runtime_error(b200mpc_last_error()); }

    b200mpc_lmpc_t h = nullptr;
    std::vector<double> rs, ri, ro;
};

}  // namespace mpc
