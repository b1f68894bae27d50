// mpc_b200/NLMPC.hpp -- C++20 host-side mirror of mpc::NLMPC<Tnx,Tnu,Tny,Tph,Tch,Tineq,Teq> (libmpc++ v0.7.1,
// include/mpc/NLMPC.hpp:24-476) over the C ABI of the B200 engine (include/b200mpc.h: b200mpc_nlmpc_solve).
//
// What carries over unchanged: bounds setters (matrix and vector + HorizonSlice forms), setOptimizerParameters
// (NLParameters), optimize()/step(), getLastResult(), getOptimalSequence(), the warm-start / bound-repair / one-stage
// shift of the decision vector (NLOptimizer::run, NLOptimizer.hpp:412-510,705-716), Result / OptSequence / ResultStatus.
//
// What cannot carry over: the reference receives the model, the objective and the constraints as host std::function
// callbacks (setStateSpaceFunction / setObjectiveFunction / setIneqConFunction / setEqConFunction, NLMPC.hpp:139-281); a
// device kernel cannot call them.  They are replaced by ONE call: setSystem(system_id, params) selects a device functor
// compiled into the engine (the reference's three example systems: B200MPC_SYS_VANDERPOL, _OSCNET4/6, _UGV), or
// setSystemSource(cuda_source, type_name, params) takes the user's own model / cost / constraints as CUDA source, which
// the engine compiles with NVRTC into its kernels (contract in include/b200mpc.h).  The callback setters are kept so that
// existing call sites fail with a clear message, not silently.
//
// Batch extension: NLMPC(batch) holds `batch` independent controllers; optimizeBatch solves them in one launch.
#pragma once
#include <mpc_b200/LMPC.hpp>

#include <algorithm>
#include <functional>

namespace mpc {

// include/mpc/Types.hpp:121-140
struct NLParameters : Parameters {
    double relative_ftol = -1, relative_xtol = -1, absolute_ftol = -1, absolute_xtol = -1;
    bool hard_constraints = true;
};

template <int Tnx = Dynamic, int Tnu = Dynamic, int Tny = Dynamic, int Tph = Dynamic, int Tch = Dynamic, int Tineq = Dynamic, int Teq = Dynamic>
class NLMPC {
public:
    explicit NLMPC(int batch = 1) requires(Tnx > 0) { init(Tnx, Tnu, Tny, Tph, Tch, Tineq, Teq, batch); }
    NLMPC(int nx, int nu, int ny, int ph, int ch, int ineq, int eq, int batch = 1) { init(nx, nu, ny, ph, ch, ineq, eq, batch); }

    int batch() const { return batch_; }

    // ---- the device-functor replacement of the three callbacks
    bool setSystem(int system_id, const std::vector<double>& params, bool per_instance = false) {
        int nx, nu, np, ni;
        detail::check(b200mpc_nlmpc_system_dims(system_id, ph_, &nx, &nu, &np, &ni));
        if (nx != nx_ || nu != nu_) throw std::runtime_error("setSystem: system dimensions do not match the template sizes");
        if (ineq_ >= 0 && ni != ineq_) throw std::runtime_error("setSystem: Tineq does not match the system's inequality count");
        int ne = 0;
        detail::check(b200mpc_nlmpc_system_neq(system_id, ph_, &ne));
        if (eq_ >= 0 && ne != eq_) throw std::runtime_error("setSystem: Teq does not match the system's equality count");
        neq_ = ne;
        if (params.size() != (size_t)np * (per_instance ? batch_ : 1)) throw std::runtime_error("setSystem: wrong parameter count");
        system_ = system_id; params_ = params; per_instance_ = per_instance; nparam_ = np;
        return true;
    }
    // the user's own system as CUDA source: struct `type_name` with static f / cost / ineq [/ eq] device functions
    bool setSystemSource(const std::string& cuda_source, const std::string& type_name, const std::vector<double>& params,
                         bool per_instance = false) {
        int id = -1;
        detail::check(b200mpc_nlmpc_register_system(cuda_source.c_str(), type_name.c_str(), &id));
        return setSystem(id, params, per_instance);
    }
    template <class F> bool setStateSpaceFunction(F&&, float = 1e-10f) { return noCallback("setStateSpaceFunction"); }
    template <class F> bool setObjectiveFunction(F&&) { return noCallback("setObjectiveFunction"); }
    template <class F> bool setOutputFunction(F&&) { return noCallback("setOutputFunction"); }
    template <class F> bool setIneqConFunction(F&&, float tol = 1e-10f) { (void)tol; return noCallback("setIneqConFunction"); }
    template <class F> bool setEqConFunction(F&&, float = 1e-10f) { return noCallback("setEqConFunction"); }
    void setIneqTolerance(double tol) { ineq_tol_ = tol; }                       // the `tol` of setIneqConFunction (NLMPC.hpp:229)
    void setEqTolerance(double tol) { eq_tol_ = tol; }                           // the `tol` of setEqConFunction (NLMPC.hpp:262)

    // NLMPC.hpp:80-95: continuous-time systems take their sampling time through the parameter vector (slot 0)
    bool setDiscretizationSamplingTime(const double ts) {
        if (system_ < 0 || params_.empty()) throw std::runtime_error("call setSystem first");
        if (system_ == B200MPC_SYS_UGV) return false;                            // discrete-time functor
        for (int b = 0; b < (per_instance_ ? batch_ : 1); ++b) params_[(size_t)b * nparam_] = ts;
        return true;
    }
    // NLMPC.hpp:108-130 -> Mapping::setInputScaling / setStateScaling (Mapping.hpp:108-150)
    void setInputScale(const cvec<Tnu> scaling) { su_.assign(scaling.data(), scaling.data() + nu_); }
    void setStateScale(const cvec<Tnx> scaling) { sx_.assign(scaling.data(), scaling.data() + nx_); }

    void setOptimizerParameters(const Parameters& param) {                       // NLMPC.hpp:97-100, NLOptimizer.hpp:150-190
        const auto* np = dynamic_cast<const NLParameters*>(&param);
        if (!np) throw std::runtime_error("NLMPC expects NLParameters");
        p_ = *np;
        applySlackBound();
    }

    // bounds: NLMPC.hpp:292-330 (matrix forms), :362-400 (vector + slice), NLOptimizer.hpp:346-400
    bool setStateBounds(const mat<Tnx, Tph>& lo, const mat<Tnx, Tph>& hi) {
        for (int i = 0; i < ph_; ++i) for (int j = 0; j < nx_; ++j) { lb_[i * nx_ + j] = lo(j, i); ub_[i * nx_ + j] = hi(j, i); }
        return true;
    }
    bool setInputBounds(const mat<Tnu, Tch>& lo, const mat<Tnu, Tch>& hi) {
        for (int i = 0; i < ch_; ++i) for (int j = 0; j < nu_; ++j) { lb_[ph_ * nx_ + i * nu_ + j] = lo(j, i); ub_[ph_ * nx_ + i * nu_ + j] = hi(j, i); }
        return true;
    }
    bool setOutputBounds(const mat<Tny, Tph>&, const mat<Tny, Tph>&) { return false; }           // ignored upstream too (:342-349)
    bool setStateBounds(const cvec<Tnx>& lo, const cvec<Tnx>& hi, const HorizonSlice& s) {
        int a, b; if (!slice(s, ph_, a, b)) return false;
        for (int i = a; i < b; ++i) for (int j = 0; j < nx_; ++j) { lb_[i * nx_ + j] = lo(j); ub_[i * nx_ + j] = hi(j); }
        return true;
    }
    bool setInputBounds(const cvec<Tnu>& lo, const cvec<Tnu>& hi, const HorizonSlice& s) {
        int a, b; if (!slice(s, ch_, a, b)) return false;
        for (int i = a; i < b; ++i) for (int j = 0; j < nu_; ++j) { lb_[ph_ * nx_ + i * nu_ + j] = lo(j); ub_[ph_ * nx_ + i * nu_ + j] = hi(j); }
        return true;
    }
    bool setOutputBounds(const cvec<Tny>&, const cvec<Tny>&, const HorizonSlice&) { return false; }

    // IMPC::optimize (IMPC.hpp:149-166); step() is the pre-0.5.0 name
    Result<Tnu> optimize(const cvec<Tnx> x0, const cvec<Tnu> lastU) {
        if (batch_ != 1) throw std::runtime_error("optimize() is the batch==1 call; use optimizeBatch()");
        optimizeBatch(x0.data(), lastU.data());
        return last_[0];
    }
    Result<Tnu> step(const cvec<Tnx> x0, const cvec<Tnu> lastU) { return optimize(x0, lastU); }

    // x0[batch*nx], lastU[batch*nu], row-major per controller
    const std::vector<Result<Tnu>>& optimizeBatch(const double* x0, const double* lastU) {
        if (system_ < 0) throw std::runtime_error("NLMPC: setSystem has not been called");
        const int nz = nz_;
        std::vector<double> z0((size_t)batch_ * nz);
        for (int b = 0; b < batch_; ++b) initialGuess(b, x0 + (size_t)b * nx_, lastU + (size_t)b * nu_, z0.data() + (size_t)b * nz);
        b200mpc_nlmpc_params q;
        b200mpc_nlmpc_default_params(&q);
        q.max_sqp = p_.maximum_iteration;
        double xt = minPositive(p_.relative_xtol, p_.absolute_xtol), ft = minPositive(p_.relative_ftol, p_.absolute_ftol);
        if (xt > 0) q.tol = xt;
        if (ft > 0) q.ftol = ft;
        std::vector<double> cost(batch_), viol(batch_);
        std::vector<int32_t> st(batch_), it(batch_), qit(batch_);
        const b200mpc_nlmpc_scaling sc{sx_.empty() ? nullptr : sx_.data(), su_.empty() ? nullptr : su_.data()};
        detail::check(b200mpc_nlmpc_solve_ex(system_, ph_, ch_, batch_, &q, z0.data(), x0, params_.data(), per_instance_ ? 1 : 0, &sc,
                                             lb_.data(), ub_.data(), opt_.data(), cost.data(), viol.data(), st.data(), it.data(),
                                             qit.data(), 0, nullptr));
        first_ = false;
        // feasibility as Constraints::isFeasible (Constraints.hpp:157-201): user inequalities against their tolerance
        int ni = 0;
        detail::check(b200mpc_nlmpc_system_dims(system_, ph_, nullptr, nullptr, nullptr, &ni));
        std::vector<double> cin((size_t)batch_ * std::max(ni, 1)), cue((size_t)batch_ * std::max(neq_, 1));
        if (ni > 0 || neq_ > 0)
            detail::check(b200mpc_nlmpc_eval_ex(system_, ph_, ch_, batch_, opt_.data(), x0, params_.data(), per_instance_ ? 1 : 0, &sc,
                                                nullptr, nullptr, nullptr, nullptr, ni > 0 ? cin.data() : nullptr, nullptr,
                                                neq_ > 0 ? cue.data() : nullptr, nullptr, 0, nullptr));
        // OptSequence::output = Model::getOutput(Xmat, Umat) (NLOptimizer.hpp:611, Model.hpp:72-96), on the device
        int sys_ny = 0;
        detail::check(b200mpc_nlmpc_system_ny(system_, ph_, &sys_ny, nullptr));
        yseq_.assign((size_t)batch_ * (ph_ + 1) * std::max(sys_ny, 1), 0.0);
        sys_ny_ = sys_ny;
        if (sys_ny > 0)
            detail::check(b200mpc_nlmpc_output(system_, ph_, ch_, batch_, opt_.data(), x0, params_.data(), per_instance_ ? 1 : 0, &sc,
                                               yseq_.data(), 0, nullptr));
        last_.assign(batch_, Result<Tnu>());
        x0_.assign(x0, x0 + (size_t)batch_ * nx_);
        for (int b = 0; b < batch_; ++b) {
            auto& r = last_[b];
            const double* z = opt_.data() + (size_t)b * nz;
            slack_[b] = z[nz - 1];
            r.cmd.resize(nu_, 1);
            for (int k = 0; k < nu_; ++k) r.cmd(k) = (su_.empty() ? 1.0 : su_[k]) * z[ph_ * nx_ + k];
            r.cost = cost[b];
            r.status = st[b] == 0 ? ResultStatus::SUCCESS : ResultStatus::MAX_ITERATION;
            r.solver_status = st[b] == 0 ? 4 : 5;                                // nlopt::XTOL_REACHED / MAXEVAL_REACHED
            r.is_feasible = true;
            for (int k = 0; k < ni; ++k) if (cin[(size_t)b * ni + k] > ineq_tol_) r.is_feasible = false;
            for (int k = 0; k < neq_; ++k) if (std::fabs(cue[(size_t)b * neq_ + k]) > eq_tol_) r.is_feasible = false;
        }
        return last_;
    }
    Result<Tnu> getLastResult() { return last_.empty() ? Result<Tnu>() : last_[0]; }

    // Mapping::unwrapVector on the stored optimum (NLOptimizer.hpp:596-611): rows 0..ph, row ph of U repeats row ph-1
    OptSequence<Tnx, Tny, Tnu, detail::dimp1(Tph)> getOptimalSequence(int instance = 0) {
        OptSequence<Tnx, Tny, Tnu, detail::dimp1(Tph)> q;
        q.state.resize(ph_ + 1, nx_); q.input.resize(ph_ + 1, nu_); q.output.resize(ph_ + 1, ny_);
        if (x0_.empty()) return q;
        const double* z = opt_.data() + (size_t)instance * nz_;
        for (int t = 0; t <= ph_; ++t) {
            for (int k = 0; k < nx_; ++k)                                                     // X / state scaling, row 0 = x0 included
                q.state(t, k) = (t == 0 ? x0_[(size_t)instance * nx_ + k] : z[(t - 1) * nx_ + k]) / (sx_.empty() ? 1.0 : sx_[k]);
            int blk = std::min(std::min(t, ph_ - 1), ch_ - 1);
            for (int k = 0; k < nu_; ++k) q.input(t, k) = (su_.empty() ? 1.0 : su_[k]) * z[ph_ * nx_ + blk * nu_ + k];
            for (int k = 0; k < ny_; ++k)                                                     // the system's output map (zeros without one)
                q.output(t, k) = k < sys_ny_ ? yseq_[((size_t)instance * (ph_ + 1) + t) * sys_ny_ + k] : 0.0;
        }
        return q;
    }

private:
    void init(int nx, int nu, int ny, int ph, int ch, int ineq, int eq, int batch) {
        nx_ = nx; nu_ = nu; ny_ = ny; ph_ = ph; ch_ = ch; ineq_ = ineq; eq_ = eq; batch_ = batch;
        nz_ = ph * nx + ch * nu + 1;
        const double finf = std::numeric_limits<float>::infinity();                  // NLOptimizer.hpp:69-73
        lb_.assign(nz_, -finf); ub_.assign(nz_, finf);
        opt_.assign((size_t)batch * nz_, 0.0);
        slack_.assign(batch, 0.0);
        applySlackBound();
    }
    void applySlackBound() {                                                         // NLOptimizer.hpp:160-190
        if (p_.hard_constraints) { lb_[nz_ - 1] = 0; ub_[nz_ - 1] = 0; }
        else { lb_[nz_ - 1] = 0; ub_[nz_ - 1] = std::numeric_limits<float>::infinity(); }
    }
    static bool slice(const HorizonSlice& s, int horizon, int& a, int& b) {
        if (s.start == -1 && s.end == -1) { a = 0; b = horizon; return true; }
        a = s.start == -1 ? 0 : s.start; b = s.end == -1 ? horizon : s.end;
        return a >= 0 && a < b && b <= horizon;
    }
    static double minPositive(double a, double b) { return a > 0 && b > 0 ? std::min(a, b) : (a > 0 ? a : b); }
    bool noCallback(const char* name) const {
        throw std::runtime_error(std::string("b200mpc NLMPC: ") + name + " takes a host callback, which a CUDA kernel cannot call; use "
                                 "setSystem(system_id, params) to select the device functor instead");
    }
    // NLOptimizer::run :431-510 -- cold tile or previous optimum, fixOptimalSolution (:705-716), one-stage left shift
    void initialGuess(int b, const double* x0, const double* u0, double* out) {
        double* z = opt_.data() + (size_t)b * nz_;
        if (first_ || !p_.enable_warm_start) {
            for (int i = 0; i < ph_; ++i) for (int j = 0; j < nx_; ++j) z[i * nx_ + j] = x0[j];
            for (int i = 0; i < ch_; ++i) for (int j = 0; j < nu_; ++j) z[ph_ * nx_ + i * nu_ + j] = u0[j];
        }
        for (int i = 0; i < nz_; ++i) if (z[i] < lb_[i] || z[i] > ub_[i]) z[i] = (ub_[i] - lb_[i]) / 2.0;   // sic (:713)
        for (int i = 0; i < ph_; ++i) for (int j = 0; j < nx_; ++j) out[i * nx_ + j] = z[std::min(i + 1, ph_ - 1) * nx_ + j];
        // Iz2u expands the ch blocks to ph stages, the stages shift left by one, Iu2z keeps the first stage of each block
        for (int i = 0; i < ch_; ++i) {
            int stage = std::min(i + 1, ph_ - 1), blk = std::min(stage, ch_ - 1);
            for (int j = 0; j < nu_; ++j) out[ph_ * nx_ + i * nu_ + j] = z[ph_ * nx_ + blk * nu_ + j];
        }
        out[nz_ - 1] = slack_[b];
    }

    int nx_ = 0, nu_ = 0, ny_ = 0, ph_ = 0, ch_ = 0, ineq_ = -1, eq_ = -1, neq_ = 0, batch_ = 1, nz_ = 0;
    int system_ = -1, nparam_ = 0;
    bool per_instance_ = false, first_ = true;
    double ineq_tol_ = 1e-10, eq_tol_ = 1e-10;
    std::vector<double> sx_, su_;
    NLParameters p_;
    std::vector<double> params_, lb_, ub_, opt_, slack_, x0_, yseq_;
    int sys_ny_ = 0;
    std::vector<Result<Tnu>> last_;
};

}  // namespace mpc
