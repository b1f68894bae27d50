// mpc_b200/LMPC.hpp -- C++20 host-side mirror of mpc::LMPC<Tnx,Tnu,Tndu,Tny,Tph,Tch> (libmpc++ v0.7.1) over the C ABI
// of the B200 engine (include/b200mpc.h).  Same setter names, argument meaning, bool returns and throwing calls as
// the reference front-end (include/mpc/LMPC.hpp:23-749, include/mpc/IMPC.hpp:149-217), so user code that only touches
// the public API switches by changing the include and linking libb200mpc.so.  Eigen is not available in this image,
// so `mat`/`cvec` are a minimal column-major stand-in with the handful of members the reference's examples use
// (operator(), setZero/setOnes/setIdentity/setConstant, data(), rows()/cols(), col()).
//
// Batch extension: `LMPC(..., batch)` creates `batch` independent controllers that share every setter and differ in
// (x0, lastU) and, optionally, in per-instance references; `optimizeBatch` solves them in one kernel launch.
#pragma once
#include <b200mpc.h>

#include <cmath>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

namespace mpc {

constexpr int Dynamic = -1;
constexpr double inf = std::numeric_limits<double>::infinity();

// ---- minimal dense matrix (column-major, like Eigen's default) -----------------------------------------------------
template <int M = Dynamic, int N = Dynamic>
class mat {
public:
    mat() : r_(M > 0 ? M : 0), c_(N > 0 ? N : 0), v_((size_t)r_ * c_, 0.0) {}
    mat(int r, int c) : r_(r), c_(c), v_((size_t)r * c, 0.0) {}
    void resize(int r, int c) { r_ = r; c_ = c; v_.assign((size_t)r * c, 0.0); }
    int rows() const { return r_; }
    int cols() const { return c_; }
    size_t size() const { return v_.size(); }
    double& operator()(int i, int j) { return v_[(size_t)j * r_ + i]; }
    double operator()(int i, int j) const { return v_[(size_t)j * r_ + i]; }
    double& operator()(int i) { return v_[i]; }
    double operator()(int i) const { return v_[i]; }
    double& operator[](int i) { return v_[i]; }
    double operator[](int i) const { return v_[i]; }
    double* data() { return v_.data(); }
    const double* data() const { return v_.data(); }
    mat& setZero() { std::fill(v_.begin(), v_.end(), 0.0); return *this; }
    mat& setOnes() { std::fill(v_.begin(), v_.end(), 1.0); return *this; }
    mat& setConstant(double x) { std::fill(v_.begin(), v_.end(), x); return *this; }
    mat& setIdentity() { setZero(); for (int i = 0; i < (r_ < c_ ? r_ : c_); ++i) (*this)(i, i) = 1.0; return *this; }
    static mat Zero() { return mat(); }
    static mat Zero(int r, int c) { return mat(r, c); }
    static mat Ones() { mat m; m.setOnes(); return m; }
    // row-major fill from a flat list (the reference's `<<` initialiser order)
    mat& fillRowMajor(std::initializer_list<double> l) {
        int k = 0;
        for (double x : l) { int i = k / c_, j = k % c_; if (i < r_) (*this)(i, j) = x; ++k; }
        return *this;
    }
    void setCol(int j, const mat<M, 1>& v) { for (int i = 0; i < r_; ++i) (*this)(i, j) = v(i); }
    // row-major copy (what the C ABI wants for A,B,C,Bd,Dd)
    std::vector<double> rowMajor() const {
        std::vector<double> o((size_t)r_ * c_);
        for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) o[(size_t)i * c_ + j] = (*this)(i, j);
        return o;
    }
private:
    int r_, c_;
    std::vector<double> v_;
};
template <int N = Dynamic> using cvec = mat<N, 1>;

// include/mpc/Types.hpp:57-79
struct HorizonSlice {
    int start, end;
    HorizonSlice(int s, int e) : start(s), end(e) {}
    static HorizonSlice all() { return HorizonSlice{-1, -1}; }
};
// include/mpc/Types.hpp:84-91
enum ResultStatus { SUCCESS, MAX_ITERATION, INFEASIBLE, ERROR, UNKNOWN };
// include/mpc/Types.hpp:99-160
struct Parameters {
    virtual ~Parameters() = default;
    int maximum_iteration = 100;
    double time_limit = 0;
    bool enable_warm_start = false;
protected:
    Parameters() = default;
};
struct LParameters : Parameters {
    double alpha = 1.6, rho = 1e-6, eps_rel = 1e-4, eps_abs = 1e-4, eps_prim_inf = 1e-3, eps_dual_inf = 1e-3;
    bool verbose = false, adaptive_rho = true, polish = true;
};
// include/mpc/Types.hpp:168-199
template <int Tnu = Dynamic>
struct Result {
    int solver_status = 0;
    bool is_feasible = false;
    std::string solver_status_msg;
    double cost = 0;
    ResultStatus status = ResultStatus::UNKNOWN;
    cvec<Tnu> cmd;
};
template <int Tnx = Dynamic, int Tny = Dynamic, int Tnu = Dynamic, int Tph = Dynamic>
struct OptSequence {
    mat<Tph, Tnx> state;
    mat<Tph, Tny> output;
    mat<Tph, Tnu> input;
};

namespace detail {
inline void check(int rc) {
    if (rc != B200MPC_OK) throw std::runtime_error(std::string("b200mpc: ") + b200mpc_last_error());
}
constexpr int dimp1(int d) { return d == Dynamic ? Dynamic : d + 1; }
}  // namespace detail

// mpc::discretization<nx,nu>(A, B, Ts, Ad, Bd)  (include/mpc/Utils.hpp:23-47): zero-order-hold c2d, on the device
template <int nx, int nu>
void discretization(const mat<nx, nx>& A, const mat<nx, nu>& B, const double& Ts, mat<nx, nx>& Ad, mat<nx, nu>& Bd) {
    const int n = A.rows(), m = B.cols();
    std::vector<double> a = A.rowMajor(), b = B.rowMajor(), ad((size_t)n * n), bd((size_t)n * m);
    detail::check(b200mpc_c2d(n, m, 1, a.data(), b.data(), 0, &Ts, 0, ad.data(), bd.data(), 0, nullptr));
    Ad.resize(n, n); Bd.resize(n, m);
    for (int i = 0; i < n; ++i) { for (int j = 0; j < n; ++j) Ad(i, j) = ad[(size_t)i * n + j]; for (int j = 0; j < m; ++j) Bd(i, j) = bd[(size_t)i * m + j]; }
}

// ---- mpc::LMPC<Tnx,Tnu,Tndu,Tny,Tph,Tch>  (include/mpc/LMPC.hpp:23-26) ------------------------------------------------
template <int Tnx = Dynamic, int Tnu = Dynamic, int Tndu = Dynamic, int Tny = Dynamic, int Tph = Dynamic, int Tch = Dynamic>
class LMPC {
public:
    LMPC() requires(Tnx > 0) { init(Tnx, Tnu, Tndu, Tny, Tph, Tch, 1); }
    LMPC(int nx, int nu, int ndu, int ny, int ph, int ch, int batch = 1) { init(nx, nu, ndu, ny, ph, ch, batch); }
    LMPC(const LMPC&) = delete;
    LMPC& operator=(const LMPC&) = delete;
    ~LMPC() { if (h_) b200mpc_lmpc_destroy(h_); }

    int batch() const { return batch_; }

    // -- unsupported, exactly as LMPC.hpp:68-100
    bool setDiscretizationSamplingTime(const double) { throw std::runtime_error("Linear MPC supports only discrete time systems"); }
    void setInputScale(const cvec<Tnu>) { throw std::runtime_error("Linear MPC does not support input scaling"); }
    void setStateScale(const cvec<Tnx>) { throw std::runtime_error("Linear MPC does not support state scaling"); }

    void setOptimizerParameters(const Parameters& param) {   // LMPC.hpp:79-82
        const auto* lp = dynamic_cast<const LParameters*>(&param);
        if (!lp) throw std::runtime_error("LMPC expects LParameters");
        b200mpc_lmpc_params q;
        b200mpc_lmpc_default_params(&q);
        q.maximum_iteration = lp->maximum_iteration; q.enable_warm_start = lp->enable_warm_start;
        q.alpha = lp->alpha; q.rho = lp->rho; q.eps_rel = lp->eps_rel; q.eps_abs = lp->eps_abs;
        q.eps_prim_inf = lp->eps_prim_inf; q.eps_dual_inf = lp->eps_dual_inf; q.adaptive_rho = lp->adaptive_rho; q.polish = lp->polish;
        q.time_limit = lp->time_limit;                                                       // LOptimizer.hpp:256
        detail::check(b200mpc_lmpc_set_params(h_, &q));
    }

    bool setStateSpaceModel(const mat<Tnx, Tnx>& A, const mat<Tnx, Tnu>& B, const mat<Tny, Tnx>& C) {   // LMPC.hpp:493
        auto a = A.rowMajor(), b = B.rowMajor(), c = C.rowMajor();
        detail::check(b200mpc_lmpc_set_model(h_, a.data(), b.data(), c.data(), 0, 0));
        return sync();
    }
    bool setDisturbances(const mat<Tnx, Tndu>& Bd, const mat<Tny, Tndu>& Dd) {                           // LMPC.hpp:518
        auto a = Bd.rowMajor(), b = Dd.rowMajor();
        detail::check(b200mpc_lmpc_set_disturbances(h_, a.data(), b.data(), 0, 0));
        return sync();
    }

    // matrix forms (LMPC.hpp:111-141, 306-313, 534, 596): mat<dim,ph> column-major == the ABI's stage-major layout
    bool setObjectiveWeights(const mat<Tny, Tph>& OW, const mat<Tnu, Tph>& UW, const mat<Tnu, Tph>& DUW) {
        ow_ = OW; uw_ = UW; duw_ = DUW; return pushWeights();
    }
    bool setStateBounds(const mat<Tnx, Tph>& lo, const mat<Tnx, Tph>& hi) { xmin_ = lo; xmax_ = hi; return pushState(); }
    bool setOutputBounds(const mat<Tny, Tph>& lo, const mat<Tny, Tph>& hi) { ymin_ = lo; ymax_ = hi; return pushOutput(); }
    bool setInputBounds(const mat<Tnu, Tch>& lo, const mat<Tnu, Tch>& hi) {
        for (int j = 0; j < ph_; ++j) for (int i = 0; i < nu_; ++i) {            // ProblemBuilder.hpp:397-413
            int sc = j < ch_ ? j : ch_ - 1;
            umin_(i, j) = lo(i, sc); umax_(i, j) = hi(i, sc);
        }
        return pushInput();
    }
    bool setReferences(const mat<Tny, Tph> yRef, const mat<Tnu, Tph> uRef, const mat<Tnu, Tph> duRef) {
        yref_ = yRef; uref_ = uRef; duref_ = duRef; return pushRefs();
    }
    bool setExogenousInputs(const mat<Tndu, Tph>& uMeas) { umeas_ = uMeas; return pushMeas(); }

    // vector + HorizonSlice forms (LMPC.hpp:153-292, 436-478, 550-676)
    bool setObjectiveWeights(const cvec<Tny>& OW, const cvec<Tnu>& UW, const cvec<Tnu>& DUW, const HorizonSlice& s) {
        int a, b; if (!predSlice(s, a, b)) return false;
        for (int j = a; j < b; ++j) { ow_.setCol(j, OW); uw_.setCol(j, UW); duw_.setCol(j, DUW); }
        return pushWeights();
    }
    bool setStateBounds(const cvec<Tnx>& lo, const cvec<Tnx>& hi, const HorizonSlice& s) {
        int a, b; if (!predSlice(s, a, b)) return false;
        for (int j = a; j < b; ++j) { xmin_.setCol(j, lo); xmax_.setCol(j, hi); }
        return pushState();
    }
    bool setOutputBounds(const cvec<Tny>& lo, const cvec<Tny>& hi, const HorizonSlice& s) {
        int a, b; if (!predSlice(s, a, b)) return false;
        for (int j = a; j < b; ++j) { ymin_.setCol(j, lo); ymax_.setCol(j, hi); }
        return pushOutput();
    }
    bool setInputBounds(const cvec<Tnu>& lo, const cvec<Tnu>& hi, const HorizonSlice& s) {
        if (s.start == -1 && s.end == -1) {                                        // LMPC.hpp:200-216
            for (int j = 0; j < ph_; ++j) { umin_.setCol(j, lo); umax_.setCol(j, hi); }
            return pushInput();
        }
        if (s.start >= s.end || s.start > ch_ || s.end > ch_) return false;        // IMPC.hpp:263-272
        for (int j = s.start; j < s.end; ++j) { umin_.setCol(j, lo); umax_.setCol(j, hi); }
        return pushInput();
    }
    bool setReferences(const cvec<Tny> yRef, const cvec<Tnu> uRef, const cvec<Tnu> duRef, const HorizonSlice& s) {
        int a, b; if (!predSlice(s, a, b)) return false;
        for (int j = a; j < b; ++j) { yref_.setCol(j, yRef); uref_.setCol(j, uRef); duref_.setCol(j, duRef); }
        return pushRefs();
    }
    bool setExogenousInputs(const cvec<Tndu>& uMeas, const HorizonSlice& s) {
        int a, b;
        if (s.start == -1 && s.end == -1) { a = 0; b = ph_; }
        else { if (s.start >= s.end || s.start > ch_ || s.end > ch_) return false; a = s.start; b = s.end; }   // LMPC.hpp:571
        for (int j = a; j < b; ++j) umeas_.setCol(j, uMeas);
        return pushMeas();
    }
    bool setScalarConstraint(const double min, const double max, const cvec<Tnx> X, const cvec<Tnu> U, const HorizonSlice& s) {
        int a, b; if (!predSlice(s, a, b)) return false;                            // LMPC.hpp:355-407
        for (int j = a; j < b; ++j) { smin_[j] = min; smax_[j] = max; }
        sx_ = X; su_ = U;
        return pushScalar();
    }
    bool setScalarConstraint(const unsigned int index, const double min, const double max, const cvec<Tnx> X, const cvec<Tnu> U) {
        if ((int)index >= ph_) return false;                                        // LMPC.hpp:409-422
        smin_[index] = min; smax_[index] = max; sx_ = X; su_ = U;
        return pushScalar();
    }
    bool setConstraints(const unsigned int index, const cvec<Tnx> XMin, const cvec<Tnu> UMin, const cvec<Tny> YMin,
                        const cvec<Tnx> XMax, const cvec<Tnu> UMax, const cvec<Tny> YMax) {          // LMPC.hpp:328-341
        if ((int)index >= ph_) return false;
        xmin_.setCol(index, XMin); xmax_.setCol(index, XMax); ymin_.setCol(index, YMin); ymax_.setCol(index, YMax);
        if ((int)index < ch_) { umin_.setCol(index, UMin); umax_.setCol(index, UMax); }
        return pushState() && pushOutput() && pushInput();
    }

    // warm start accessors (LMPC.hpp:677-722), batch == 1 view
    std::vector<double> getSolverWarmStartPrimal() { std::vector<double> x((size_t)batch_ * n()); detail::check(b200mpc_lmpc_get_warm_start(h_, x.data(), nullptr, 0)); return x; }
    std::vector<double> getSolverWarmStartDual() { std::vector<double> y((size_t)batch_ * m()); detail::check(b200mpc_lmpc_get_warm_start(h_, nullptr, y.data(), 0)); return y; }
    void setSolverWarmStart(std::vector<double> primal, std::vector<double> dual) {
        if (primal.size() != (size_t)batch_ * n() || dual.size() != (size_t)batch_ * m()) throw std::runtime_error("warm start size");
        detail::check(b200mpc_lmpc_set_warm_start(h_, primal.data(), dual.data(), 0));
        sync();
    }

    // IMPC::optimize (IMPC.hpp:149-166); step() is the pre-0.5.0 name of the same call (CHANGELOG.md:64-65)
    Result<Tnu> optimize(const cvec<Tnx> x0, const cvec<Tnu> lastU) {
        if (batch_ != 1) throw std::runtime_error("optimize() is the batch==1 call; use optimizeBatch()");
        optimizeBatch(x0.data(), lastU.data());
        return last_[0];
    }
    Result<Tnu> step(const cvec<Tnx> x0, const cvec<Tnu> lastU) { return optimize(x0, lastU); }
    // x0[batch*nx], lastU[batch*nu] row-major per instance
    const std::vector<Result<Tnu>>& optimizeBatch(const double* x0, const double* lastU) {
        detail::check(b200mpc_lmpc_solve(h_, x0, lastU, 0));
        std::vector<double> cmd((size_t)batch_ * nu_), cost(batch_);
        std::vector<int32_t> st(batch_), sst(batch_), feas(batch_);
        detail::check(b200mpc_lmpc_get_result(h_, cmd.data(), cost.data(), st.data(), sst.data(), feas.data(), nullptr, nullptr, nullptr, 0));
        last_.assign(batch_, Result<Tnu>());
        for (int b = 0; b < batch_; ++b) {
            auto& r = last_[b];
            r.cmd.resize(nu_, 1);
            for (int k = 0; k < nu_; ++k) r.cmd(k) = cmd[(size_t)b * nu_ + k];
            r.cost = cost[b]; r.status = (ResultStatus)st[b]; r.solver_status = sst[b]; r.is_feasible = feas[b] != 0;
        }
        return last_;
    }
    // `steps` control steps on the device (optimize -> apply cmd -> x+ = A x + B u with the controller's own model): the loop
    // every example wraps around optimize().  traj_x[(steps+1)*batch*nx], traj_u[steps*batch*nu], row-major per instance.
    void closedLoop(const double* x0, const double* lastU, int steps, std::vector<double>& traj_x, std::vector<double>& traj_u,
                    std::vector<int32_t>* status = nullptr) {
        traj_x.assign((size_t)(steps + 1) * batch_ * nx_, 0.0); traj_u.assign((size_t)steps * batch_ * nu_, 0.0);
        if (status) status->assign((size_t)steps * batch_, 0);
        detail::check(b200mpc_lmpc_closed_loop(h_, x0, lastU, steps, nullptr, nullptr, 0, traj_x.data(), traj_u.data(),
                                               status ? status->data() : nullptr, nullptr, 0));
    }
    Result<Tnu> getLastResult() { return last_.empty() ? Result<Tnu>() : last_[0]; }
    OptSequence<Tnx, Tny, Tnu, detail::dimp1(Tph)> getOptimalSequence(int instance = 0) {
        std::vector<double> s((size_t)batch_ * (ph_ + 1) * nx_), i((size_t)batch_ * (ph_ + 1) * nu_), o((size_t)batch_ * (ph_ + 1) * ny_);
        detail::check(b200mpc_lmpc_get_sequence(h_, s.data(), i.data(), o.data(), 0));
        OptSequence<Tnx, Tny, Tnu, detail::dimp1(Tph)> q;
        q.state.resize(ph_ + 1, nx_); q.input.resize(ph_ + 1, nu_); q.output.resize(ph_ + 1, ny_);
        for (int t = 0; t <= ph_; ++t) {
            for (int k = 0; k < nx_; ++k) q.state(t, k) = s[((size_t)instance * (ph_ + 1) + t) * nx_ + k];
            for (int k = 0; k < nu_; ++k) q.input(t, k) = i[((size_t)instance * (ph_ + 1) + t) * nu_ + k];
            for (int k = 0; k < ny_; ++k) q.output(t, k) = o[((size_t)instance * (ph_ + 1) + t) * ny_ + k];
        }
        return q;
    }
    b200mpc_lmpc_t handle() { return h_; }
    /// Multi-GPU exchange step (no counterpart in the single-controller reference): all-gather of the command blocks
    /// of every rank into cmd_all_dev[nranks * batch * nu] (device memory), enqueued on the handle's stream behind the solve.
    bool allgatherCommands(b200mpc_comm_t comm, double* cmd_all_dev) { return b200mpc_lmpc_allgather_cmd(h_, comm, cmd_all_dev) == B200MPC_OK; }

private:
    int n() const { return (ph_ + 1) * (nx_ + nu_) + ph_ * nu_; }
    int m() const { return 2 * (ph_ + 1) * (nx_ + nu_) + (ph_ + 1) * ny_ + ph_ * nu_ + (ph_ + 1); }
    void init(int nx, int nu, int ndu, int ny, int ph, int ch, int batch) {
        nx_ = nx; nu_ = nu; ndu_ = ndu; ny_ = ny; ph_ = ph; ch_ = ch; batch_ = batch;
        b200mpc_lmpc_dims d{nx, nu, ndu, ny, ph, ch};
        detail::check(b200mpc_lmpc_create(&d, batch, 0, &h_));
        ow_.resize(ny, ph); uw_.resize(nu, ph); duw_.resize(nu, ph);
        xmin_.resize(nx, ph); xmax_.resize(nx, ph); ymin_.resize(ny, ph); ymax_.resize(ny, ph); umin_.resize(nu, ph); umax_.resize(nu, ph);
        xmin_.setConstant(-inf); xmax_.setConstant(inf); ymin_.setConstant(-inf); ymax_.setConstant(inf); umin_.setConstant(-inf); umax_.setConstant(inf);
        yref_.resize(ny, ph); uref_.resize(nu, ph); duref_.resize(nu, ph); umeas_.resize(ndu, ph);
        smin_.assign(ph, -inf); smax_.assign(ph, inf); sx_.resize(nx, 1); su_.resize(nu, 1);
    }
    bool predSlice(const HorizonSlice& s, int& a, int& b) const {                   // IMPC.hpp:245-261
        if (s.start == -1 && s.end == -1) { a = 0; b = ph_; return true; }
        if (s.start >= s.end || s.start > ph_ || s.end > ph_) return false;
        a = s.start; b = s.end; return true;
    }
    bool sync() { detail::check(b200mpc_sync(h_)); return true; }
    bool pushWeights() { detail::check(b200mpc_lmpc_set_weights(h_, ow_.data(), uw_.data(), duw_.data(), 0, 0)); return sync(); }
    bool pushState() { detail::check(b200mpc_lmpc_set_state_bounds(h_, xmin_.data(), xmax_.data(), 0, 0)); return sync(); }
    bool pushOutput() { detail::check(b200mpc_lmpc_set_output_bounds(h_, ymin_.data(), ymax_.data(), 0, 0)); return sync(); }
    bool pushInput() {   // all ph internal columns verbatim: the per-index setter never re-replicates the tail (ProblemBuilder.hpp:469-477)
        detail::check(b200mpc_lmpc_set_input_bounds_full(h_, umin_.data(), umax_.data(), 0, 0)); return sync();
    }
    bool pushRefs() { detail::check(b200mpc_lmpc_set_references(h_, yref_.data(), uref_.data(), duref_.data(), 0, 0)); return sync(); }
    bool pushMeas() { detail::check(b200mpc_lmpc_set_exogenous_inputs(h_, umeas_.data(), 0, 0)); return sync(); }
    bool pushScalar() { detail::check(b200mpc_lmpc_set_scalar_constraint(h_, smin_.data(), smax_.data(), sx_.data(), su_.data(), 0, 0)); return sync(); }

    b200mpc_lmpc_t h_ = nullptr;
    int nx_ = 0, nu_ = 0, ndu_ = 0, ny_ = 0, ph_ = 0, ch_ = 0, batch_ = 1;
    mat<Tny, Tph> ow_, ymin_, ymax_, yref_;
    mat<Tnu, Tph> uw_, duw_, umin_, umax_, uref_, duref_;
    mat<Tnx, Tph> xmin_, xmax_;
    mat<Tndu, Tph> umeas_;
    std::vector<double> smin_, smax_;
    cvec<Tnx> sx_;
    cvec<Tnu> su_;
    std::vector<Result<Tnu>> last_;
};

}  // namespace mpc
