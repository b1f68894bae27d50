// b200mpc_nlmpc.cu -- NLMPC entry points of the C ABI (include/b200mpc.h): problem evaluation (K5) and solve (K6/K7).
// The kernels themselves are instantiated per system in nlmpc_sys_*.cu.
#include "capi_common.h"
#include "nlmpc_sqp.cuh"
#include "nlmpc_launch_choice.h"

using namespace b200mpc;

namespace b200mpc {
template <class S> int nl_eval_t(const NlEvalArgs& a, cudaStream_t stream);
template <class S> int nl_solve_t(NlSolveArgs& a, cudaStream_t stream, std::vector<void*>& tofree);
extern template int nl_eval_t<SysVanDerPol>(const NlEvalArgs&, cudaStream_t);
extern template int nl_eval_t<SysOscNet<4>>(const NlEvalArgs&, cudaStream_t);
extern template int nl_eval_t<SysOscNet<6>>(const NlEvalArgs&, cudaStream_t);
extern template int nl_eval_t<SysUgv>(const NlEvalArgs&, cudaStream_t);
template <class S> int nl_plant_t(const NlPlantArgs& a, cudaStream_t stream);
extern template int nl_plant_t<SysVanDerPol>(const NlPlantArgs&, cudaStream_t);
extern template int nl_plant_t<SysOscNet<4>>(const NlPlantArgs&, cudaStream_t);
extern template int nl_plant_t<SysOscNet<6>>(const NlPlantArgs&, cudaStream_t);
extern template int nl_plant_t<SysUgv>(const NlPlantArgs&, cudaStream_t);
extern template int nl_solve_t<SysVanDerPol>(NlSolveArgs&, cudaStream_t, std::vector<void*>&);
extern template int nl_solve_t<SysOscNet<4>>(NlSolveArgs&, cudaStream_t, std::vector<void*>&);
extern template int nl_solve_t<SysOscNet<6>>(NlSolveArgs&, cudaStream_t, std::vector<void*>&);
extern template int nl_solve_t<SysUgv>(NlSolveArgs&, cudaStream_t, std::vector<void*>&);
// user-defined systems (b200mpc_nlmpc_rtc.cu)
bool rtc_is_user(int system);
int rtc_dims(int system, int ph, int* nx, int* nu, int* nparam, int* nineq, int* neq, int* ny = nullptr, int* has_out = nullptr,
             int* struct_ok = nullptr, int* K = nullptr);
int rtc_eval(int system, const NlEvalArgs& a, cudaStream_t stream);
int rtc_solve(int system, NlSolveArgs& a, cudaStream_t stream, std::vector<void*>& tofree);
int rtc_plant(int system, const NlPlantArgs& a, cudaStream_t stream);
}

static int g_nl_solver = 0;
int b200mpc::nl_solver_override() { return g_nl_solver; }
extern "C" int b200mpc_nlmpc_set_solver(int solver) {
    if (solver < 0 || solver > 2) return fail(B200MPC_EINVAL, "solver must be 0 (automatic), 1 (dense) or 2 (stage-structured)");
    g_nl_solver = solver;
    return B200MPC_OK;
}

// ---- NLMPC problem evaluation (K5) --------------------------------------------------------------------------------
static int nl_dims(int system, int* nx, int* nu, int* nparam, int ph, int* nineq, int* neq) {
    *neq = 0;
    switch (system) {
    case B200MPC_SYS_VANDERPOL: *nx = 2; *nu = 1; *nparam = 1; *nineq = ph + 1; return 0;
    case B200MPC_SYS_OSCNET4: *nx = 8; *nu = 4; *nparam = 3; *nineq = (ph + 1) * 4; return 0;
    case B200MPC_SYS_OSCNET6: *nx = 12; *nu = 6; *nparam = 3; *nineq = (ph + 1) * 6; return 0;
    case B200MPC_SYS_UGV: *nx = 4; *nu = 2; *nparam = 32; *nineq = (ph + 1) * 2; return 0;
    default:
        if (rtc_is_user(system)) return rtc_dims(system, ph, nx, nu, nparam, nineq, neq);
        return fail(B200MPC_EINVAL, "unknown system id");
    }
}

extern "C" int b200mpc_nlmpc_system_dims(int system, int ph, int* nx, int* nu, int* nparam, int* nineq) {
    int a, b, c, d, e;
    int rc = nl_dims(system, &a, &b, &c, ph, &d, &e);
    if (rc) return rc;
    if (nx) *nx = a; if (nu) *nu = b; if (nparam) *nparam = c; if (nineq) *nineq = d;
    return B200MPC_OK;
}
// Tny of the system and whether it defines an output map (NLMPC::setOutputFunction, NLMPC.hpp:202)
extern "C" int b200mpc_nlmpc_system_ny(int system, int ph, int* ny, int* has_output_map) {
    int v = 0, ho = 0;
    switch (system) {
    case B200MPC_SYS_VANDERPOL: v = 2; break;
    case B200MPC_SYS_OSCNET4: v = 8; break;
    case B200MPC_SYS_OSCNET6: v = 12; break;
    case B200MPC_SYS_UGV: v = 4; ho = 1; break;
    default: {
        if (!rtc_is_user(system)) return fail(B200MPC_EINVAL, "unknown system id");
        int rc = rtc_dims(system, ph, nullptr, nullptr, nullptr, nullptr, nullptr, &v, &ho);
        if (rc) return rc;
    }
    }
    if (ny) *ny = v; if (has_output_map) *has_output_map = ho;
    return B200MPC_OK;
}
extern "C" int b200mpc_nlmpc_system_neq(int system, int ph, int* neq) {
    int a, b, c, d, e;
    int rc = nl_dims(system, &a, &b, &c, ph, &d, &e);
    if (rc) return rc;
    if (neq) *neq = e;
    return B200MPC_OK;
}

// Temporaries come from the stream-ordered allocator with a pool that keeps its memory: after the first call a solve does no
// cudaMalloc / cudaFree (each costs up to milliseconds and cudaFree synchronises the device).
static void keep_pool(int dev) {
    static bool done[64] = {false};
    if (dev < 0 || dev >= 64 || done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[dev] = true;
}

// state / input scaling vectors are tiny host arrays: stage them with stream-ordered allocations
static int stage_scaling(const b200mpc_nlmpc_scaling* sc, int nx, int nu, const double** dsx, const double** dsu, cudaStream_t stream,
                         std::vector<void*>& async_free) {
    *dsx = *dsu = nullptr;
    if (!sc) return 0;
    auto up = [&](const double* h, int n, const double** d) -> int {
        if (!h) return 0;
        for (int i = 0; i < n; ++i) if (!(h[i] > 0.0)) return fail(B200MPC_EINVAL, "scaling factors must be positive");
        double* p = nullptr;
        CK(cudaMallocAsync(&p, n * sizeof(double), stream)); async_free.push_back(p);
        CK(cudaMemcpyAsync(p, h, n * sizeof(double), cudaMemcpyHostToDevice, stream));
        *d = p; return 0;
    };
    int rc;
    if ((rc = up(sc->state_scale, nx, dsx))) return rc;
    return up(sc->input_scale, nu, dsu);
}

extern "C" int b200mpc_nlmpc_eval(int system, int ph, int ch, int batch, const double* z, const double* x0, const double* params,
                                  int params_per_instance, double* fval, double* grad, double* ceq, double* Jeq, double* cin,
                                  double* Jin, int dev, void* stream_) {
    return b200mpc_nlmpc_eval_ex(system, ph, ch, batch, z, x0, params, params_per_instance, nullptr, fval, grad, ceq, Jeq, cin, Jin,
                                 nullptr, nullptr, dev, stream_);
}

static int nl_eval_impl(int system, int ph, int ch, int batch, const double* z, const double* x0, const double* params,
                        int params_per_instance, const b200mpc_nlmpc_scaling* scaling, double* fval, double* grad,
                        double* ceq, double* Jeq, double* cin, double* Jin, double* cue, double* Jue, double* yout, int dev,
                        void* stream_);

extern "C" int b200mpc_nlmpc_eval_ex(int system, int ph, int ch, int batch, const double* z, const double* x0, const double* params,
                                     int params_per_instance, const b200mpc_nlmpc_scaling* scaling, double* fval, double* grad,
                                     double* ceq, double* Jeq, double* cin, double* Jin, double* cue, double* Jue, int dev,
                                     void* stream_) {
    return nl_eval_impl(system, ph, ch, batch, z, x0, params, params_per_instance, scaling, fval, grad, ceq, Jeq, cin, Jin, cue, Jue,
                        nullptr, dev, stream_);
}

// OptSequence::output of NLOptimizer::run (NLOptimizer.hpp:596-611): Model::getOutput (Model.hpp:72-96) of the unwrapped
// sequences of z, y[batch*(ph+1)*ny]; zeros for a system without an output map, exactly as the reference.
extern "C" int b200mpc_nlmpc_output(int system, int ph, int ch, int batch, const double* z, const double* x0, const double* params,
                                    int params_per_instance, const b200mpc_nlmpc_scaling* scaling, double* y, int dev, void* stream_) {
    if (!y) return fail(B200MPC_EINVAL, "null output pointer");
    return nl_eval_impl(system, ph, ch, batch, z, x0, params, params_per_instance, scaling, nullptr, nullptr, nullptr, nullptr, nullptr,
                        nullptr, nullptr, nullptr, y, dev, stream_);
}

static int nl_eval_impl(int system, int ph, int ch, int batch, const double* z, const double* x0, const double* params,
                        int params_per_instance, const b200mpc_nlmpc_scaling* scaling, double* fval, double* grad,
                        double* ceq, double* Jeq, double* cin, double* Jin, double* cue, double* Jue, double* yout, int dev,
                        void* stream_) {
    if (b200mpc_device_count() <= 0) return fail(B200MPC_ENOGPU, "no CUDA device: b200mpc has no CPU fallback");
    int nx, nu, np, ni, nue, rc0;
    if ((rc0 = nl_dims(system, &nx, &nu, &np, ph, &ni, &nue))) return rc0;
    { int cur = 0; if (cudaGetDevice(&cur) == cudaSuccess) keep_pool(cur); }
    if (ph < 1 || ch < 1 || ch > ph || batch < 1 || !z || !x0 || !params) return fail(B200MPC_EINVAL, "bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int nz = ph * nx + ch * nu + 1;
    NlEvalArgs a;
    a.ph = ph; a.ch = ch; a.batch = batch; a.param_stride = params_per_instance ? np : 0;
    int ny = 0;
    if (yout && (rc0 = b200mpc_nlmpc_system_ny(system, ph, &ny, nullptr))) return rc0;
    std::vector<void*> tofree;
    // every temporary is released on every exit path (stream-ordered: after the work enqueued so far)
    struct Free { std::vector<void*>& v; cudaStream_t s; ~Free() { for (void* p : v) cudaFreeAsync(p, s); } } freer{tofree, stream};
    auto in = [&](const double* h, size_t n, const double** d) -> int {
        if (dev) { *d = h; return 0; }
        double* p = nullptr;
        CK(cudaMallocAsync(&p, n * sizeof(double), stream)); tofree.push_back(p);
        CK(cudaMemcpyAsync(p, h, n * sizeof(double), cudaMemcpyHostToDevice, stream));
        *d = p; return 0;
    };
    auto out = [&](double* h, size_t n, double** d) -> int {
        if (!h) { *d = nullptr; return 0; }
        if (dev) { *d = h; return 0; }
        double* p = nullptr;
        CK(cudaMallocAsync(&p, n * sizeof(double), stream)); tofree.push_back(p);
        *d = p; return 0;
    };
    int rc;
    if ((rc = in(z, (size_t)batch * nz, &a.z))) return rc;
    if ((rc = in(x0, (size_t)batch * nx, &a.x0))) return rc;
    if ((rc = in(params, (size_t)(params_per_instance ? batch : 1) * np, &a.params))) return rc;
    if ((rc = out(fval, batch, &a.fval))) return rc;
    if ((rc = out(grad, (size_t)batch * nz, &a.grad))) return rc;
    if ((rc = out(ceq, (size_t)batch * ph * nx, &a.ceq))) return rc;
    if ((rc = out(Jeq, (size_t)batch * ph * nx * nz, &a.Jeq))) return rc;
    if ((rc = out(cin, (size_t)batch * ni, &a.cin))) return rc;
    if ((rc = out(Jin, (size_t)batch * ni * nz, &a.Jin))) return rc;
    if ((rc = out(nue ? cue : nullptr, (size_t)batch * nue, &a.cue))) return rc;
    if ((rc = out(nue ? Jue : nullptr, (size_t)batch * nue * nz, &a.Jue))) return rc;
    if ((rc = out(ny ? yout : nullptr, (size_t)batch * (ph + 1) * ny, &a.yout))) return rc;
    if ((rc = stage_scaling(scaling, nx, nu, &a.sx, &a.su, stream, tofree))) return rc;
    switch (system) {
    case B200MPC_SYS_VANDERPOL: rc = nl_eval_t<SysVanDerPol>(a, stream); break;
    case B200MPC_SYS_OSCNET4: rc = nl_eval_t<SysOscNet<4>>(a, stream); break;
    case B200MPC_SYS_OSCNET6: rc = nl_eval_t<SysOscNet<6>>(a, stream); break;
    case B200MPC_SYS_UGV: rc = nl_eval_t<SysUgv>(a, stream); break;
    default: rc = rtc_eval(system, a, stream); break;
    }
    if (rc) return rc;
    if (!dev) {
        auto back = [&](double* h, const double* d, size_t n) -> int { if (h) CK(cudaMemcpyAsync(h, d, n * sizeof(double), cudaMemcpyDeviceToHost, stream)); return 0; };
        if ((rc = back(fval, a.fval, batch))) return rc;
        if ((rc = back(grad, a.grad, (size_t)batch * nz))) return rc;
        if ((rc = back(ceq, a.ceq, (size_t)batch * ph * nx))) return rc;
        if ((rc = back(Jeq, a.Jeq, (size_t)batch * ph * nx * nz))) return rc;
        if ((rc = back(cin, a.cin, (size_t)batch * ni))) return rc;
        if ((rc = back(Jin, a.Jin, (size_t)batch * ni * nz))) return rc;
        if (nue && (rc = back(cue, a.cue, (size_t)batch * nue))) return rc;
        if (nue && (rc = back(Jue, a.Jue, (size_t)batch * nue * nz))) return rc;
        if (ny && (rc = back(yout, a.yout, (size_t)batch * (ph + 1) * ny))) return rc;
        CK(cudaStreamSynchronize(stream));
    }
    return B200MPC_OK;
}


// ---- NLMPC solve (K6/K7) ---------------------------------------------------------------------------------------------
extern "C" void b200mpc_nlmpc_default_params(b200mpc_nlmpc_params* p) {
    if (!p) return;
    p->max_sqp = 100; p->max_qp = 200; p->tol = 1e-7; p->ftol = 1e-12; p->qp_eps = 1e-5; p->rho = 0.1;
}

// shared memory one controller needs when the matrices are shared-memory resident (the fast path)
static size_t nl_solve_smem(int system, int ph, int ch) {
    int nx, nu, np, ni, nue;
    if (nl_dims(system, &nx, &nu, &np, ph, &ni, &nue)) return 0;
    int n = ph * nx + ch * nu + 1, me = ph * nx;
    return NlWs::smem_doubles(0, n, me, ni + nue, ph, nx, nu) * sizeof(double);
}
extern "C" long long b200mpc_nlmpc_solve_smem_bytes(int system, int ph, int ch) { return (long long)nl_solve_smem(system, ph, ch); }

extern "C" int b200mpc_nlmpc_solve(int system, int ph, int ch, int batch, const b200mpc_nlmpc_params* prm, const double* z0,
                                   const double* x0, const double* sys_params, int params_per_instance, const double* lb,
                                   const double* ub, double* z, double* cost, double* viol, int32_t* status, int32_t* iters,
                                   int32_t* qp_iters, int dev, void* stream_) {
    return b200mpc_nlmpc_solve_ex(system, ph, ch, batch, prm, z0, x0, sys_params, params_per_instance, nullptr, lb, ub, z, cost, viol,
                                  status, iters, qp_iters, dev, stream_);
}

static int nl_solve_impl(int system, int ph, int ch, int batch, const b200mpc_nlmpc_params* prm, const double* z0,
                         const double* x0, const double* sys_params, int params_per_instance,
                         const b200mpc_nlmpc_scaling* scaling, const double* dsx, const double* dsu, const double* lb, const double* ub,
                         double* z, double* cost, double* viol, int32_t* status, int32_t* iters, int32_t* qp_iters, int dev,
                         void* stream_, bool sync);

extern "C" int b200mpc_nlmpc_solve_ex(int system, int ph, int ch, int batch, const b200mpc_nlmpc_params* prm, const double* z0,
                                      const double* x0, const double* sys_params, int params_per_instance,
                                      const b200mpc_nlmpc_scaling* scaling, const double* lb, const double* ub, double* z,
                                      double* cost, double* viol, int32_t* status, int32_t* iters, int32_t* qp_iters, int dev,
                                      void* stream_) {
    return nl_solve_impl(system, ph, ch, batch, prm, z0, x0, sys_params, params_per_instance, scaling, nullptr, nullptr, lb, ub, z, cost,
                         viol, status, iters, qp_iters, dev, stream_, true);
}

// dsx / dsu: scaling vectors already on the device (closed loop), used instead of `scaling`
static int nl_solve_impl(int system, int ph, int ch, int batch, const b200mpc_nlmpc_params* prm, const double* z0,
                         const double* x0, const double* sys_params, int params_per_instance,
                         const b200mpc_nlmpc_scaling* scaling, const double* dsx, const double* dsu, const double* lb, const double* ub,
                         double* z, double* cost, double* viol, int32_t* status, int32_t* iters, int32_t* qp_iters, int dev,
                         void* stream_, bool sync) {
    if (b200mpc_device_count() <= 0) return fail(B200MPC_ENOGPU, "no CUDA device: b200mpc has no CPU fallback");
    int nx, nu, np, ni, nue, rc0;
    if ((rc0 = nl_dims(system, &nx, &nu, &np, ph, &ni, &nue))) return rc0;
    { int cur = 0; if (cudaGetDevice(&cur) == cudaSuccess) keep_pool(cur); }
    if (ph < 1 || ch < 1 || ch > ph || batch < 1 || !z0 || !x0 || !sys_params || !lb || !ub || !z) return fail(B200MPC_EINVAL, "bad arguments");
    b200mpc_nlmpc_params q;
    if (prm) q = *prm; else b200mpc_nlmpc_default_params(&q);
    if (q.max_sqp < 1 || q.max_qp < 25 || !(q.tol > 0) || !(q.ftol >= 0) || !(q.qp_eps > 0) || !(q.rho > 0)) return fail(B200MPC_EINVAL, "bad NLMPC parameters");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int nz = ph * nx + ch * nu + 1;
    NlSolveArgs a;
    a.ph = ph; a.ch = ch; a.batch = batch; a.param_stride = params_per_instance ? np : 0;
    a.max_sqp = q.max_sqp; a.max_qp = q.max_qp; a.tol = q.tol; a.ftol = q.ftol; a.qp_eps = q.qp_eps; a.rho0 = q.rho;
    std::vector<void*> tofree;
    struct Free { std::vector<void*>& v; cudaStream_t s; ~Free() { for (void* p : v) cudaFreeAsync(p, s); } } freer{tofree, stream};
    auto in = [&](const double* h, size_t n, const double** d) -> int {
        if (dev) { *d = h; return 0; }
        double* p = nullptr;
        CK(cudaMallocAsync(&p, n * sizeof(double), stream)); tofree.push_back(p);
        CK(cudaMemcpyAsync(p, h, n * sizeof(double), cudaMemcpyHostToDevice, stream));
        *d = p; return 0;
    };
    auto out = [&](void* h, size_t bytes, void** d) -> int {
        if (dev && h) { *d = h; return 0; }
        void* p = nullptr;
        CK(cudaMallocAsync(&p, bytes, stream)); tofree.push_back(p);
        *d = p; return 0;
    };
    int rc;
    if ((rc = in(z0, (size_t)batch * nz, &a.z0))) return rc;
    if ((rc = in(x0, (size_t)batch * nx, &a.x0))) return rc;
    if ((rc = in(sys_params, (size_t)(params_per_instance ? batch : 1) * np, &a.params))) return rc;
    if ((rc = in(lb, nz, &a.lb))) return rc;
    if ((rc = in(ub, nz, &a.ub))) return rc;
    if ((rc = out(z, (size_t)batch * nz * 8, (void**)&a.z_out))) return rc;
    if ((rc = out(cost, (size_t)batch * 8, (void**)&a.cost))) return rc;
    if ((rc = out(viol, (size_t)batch * 8, (void**)&a.viol))) return rc;
    if ((rc = out(status, (size_t)batch * 4, (void**)&a.status))) return rc;
    if ((rc = out(iters, (size_t)batch * 4, (void**)&a.iters))) return rc;
    if ((rc = out(qp_iters, (size_t)batch * 4, (void**)&a.qp_iters))) return rc;
    if ((rc = stage_scaling(scaling, nx, nu, &a.sx, &a.su, stream, tofree))) return rc;      // freed after the final synchronise
    {   // work counter of the stage-structured kernel (controllers are drawn dynamically)
        void* cnt = nullptr;
        CK(cudaMallocAsync(&cnt, sizeof(int), stream)); tofree.push_back(cnt);
        CK(cudaMemsetAsync(cnt, 0, sizeof(int), stream));
        a.counter = (int*)cnt;
    }
    if (dsx) a.sx = dsx;
    if (dsu) a.su = dsu;
    switch (system) {
    case B200MPC_SYS_VANDERPOL: rc = nl_solve_t<SysVanDerPol>(a, stream, tofree); break;
    case B200MPC_SYS_OSCNET4: rc = nl_solve_t<SysOscNet<4>>(a, stream, tofree); break;
    case B200MPC_SYS_OSCNET6: rc = nl_solve_t<SysOscNet<6>>(a, stream, tofree); break;
    case B200MPC_SYS_UGV: rc = nl_solve_t<SysUgv>(a, stream, tofree); break;
    default: rc = rtc_solve(system, a, stream, tofree); break;
    }
    if (rc) return rc;
    if (!dev) {
        auto back = [&](void* h, const void* d, size_t bytes) -> int { if (h) CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, stream)); return 0; };
        if ((rc = back(z, a.z_out, (size_t)batch * nz * 8))) return rc;
        if ((rc = back(cost, a.cost, (size_t)batch * 8))) return rc;
        if ((rc = back(viol, a.viol, (size_t)batch * 8))) return rc;
        if ((rc = back(status, a.status, (size_t)batch * 4))) return rc;
        if ((rc = back(iters, a.iters, (size_t)batch * 4))) return rc;
        if ((rc = back(qp_iters, a.qp_iters, (size_t)batch * 4))) return rc;
    }
    if (sync || !dev) CK(cudaStreamSynchronize(stream));     // (temporaries are released in stream order either way)
    return B200MPC_OK;
}

// ---- plant step / RK4 and the closed loop on the device ---------------------------------------------
static int nl_plant_dispatch(int system, const NlPlantArgs& a, cudaStream_t stream) {
    switch (system) {
    case B200MPC_SYS_VANDERPOL: return nl_plant_t<SysVanDerPol>(a, stream);
    case B200MPC_SYS_OSCNET4: return nl_plant_t<SysOscNet<4>>(a, stream);
    case B200MPC_SYS_OSCNET6: return nl_plant_t<SysOscNet<6>>(a, stream);
    case B200MPC_SYS_UGV: return nl_plant_t<SysUgv>(a, stream);
    default: return rtc_plant(system, a, stream);
    }
}

// mpc::RK4<N>::run(t, in, h, integration_step) (include/mpc/Integrator.hpp:38-56) for `batch` states at once; the vector field is
// the system's model with the input held: dx/dt = f(x, u, stage, params).
extern "C" int b200mpc_nlmpc_rk4(int system, int batch, int stage, const double* x, const double* u, const double* sys_params,
                                 int params_per_instance, double h, int integration_steps, double* x_out, int dev, void* stream_) {
    if (b200mpc_device_count() <= 0) return fail(B200MPC_ENOGPU, "no CUDA device: b200mpc has no CPU fallback");
    int nx, nu, np, ni, nue, rc;
    if ((rc = nl_dims(system, &nx, &nu, &np, 1, &ni, &nue))) return rc;
    if (batch < 1 || !x || (!u && nu) || !sys_params || !x_out || integration_steps < 0) return fail(B200MPC_EINVAL, "bad arguments");
    { int cur = 0; if (cudaGetDevice(&cur) == cudaSuccess) keep_pool(cur); }
    cudaStream_t stream = (cudaStream_t)stream_;
    std::vector<void*> tofree;
    struct Free { std::vector<void*>& v; cudaStream_t s; ~Free() { for (void* p : v) cudaFreeAsync(p, s); } } freer{tofree, stream};
    auto in = [&](const double* h_, size_t n, const double** d) -> int {
        if (dev) { *d = h_; return 0; }
        double* p = nullptr;
        CK(cudaMallocAsync(&p, (n ? n : 1) * sizeof(double), stream)); tofree.push_back(p);
        if (n) CK(cudaMemcpyAsync(p, h_, n * sizeof(double), cudaMemcpyHostToDevice, stream));
        *d = p; return 0;
    };
    NlPlantArgs a{};
    a.batch = batch; a.mode = 2; a.substeps = integration_steps; a.stage = stage; a.h = h;
    a.param_stride = params_per_instance ? np : 0;
    if ((rc = in(x, (size_t)batch * nx, &a.x))) return rc;
    if ((rc = in(u, (size_t)batch * nu, &a.u_in))) return rc;
    if ((rc = in(sys_params, (size_t)(params_per_instance ? batch : 1) * np, &a.params))) return rc;
    double* dout = x_out;
    if (!dev) { CK(cudaMallocAsync(&dout, (size_t)batch * nx * sizeof(double), stream)); tofree.push_back(dout); }
    a.x_out = dout;
    if ((rc = nl_plant_dispatch(system, a, stream))) return rc;
    if (!dev) {
        CK(cudaMemcpyAsync(x_out, dout, (size_t)batch * nx * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }
    return B200MPC_OK;
}

// The control loop of the NLMPC examples (examples/vanderpol_ex.cpp:76-85, ugv_ex.cpp:143-166) for `steps` control steps without
// leaving the GPU: every step builds NLOptimizer::run's initial guess on the device (cold tile or previous optimum, bound repair,
// one-stage shift, slack carry-over: NLOptimizer.hpp:425-510), solves, applies cmd = first control block and steps the plant.
extern "C" int b200mpc_nlmpc_closed_loop(int system, int ph, int ch, int batch, const b200mpc_nlmpc_params* prm, const double* x0,
                                         const double* u0, const double* sys_params, int params_per_instance,
                                         const b200mpc_nlmpc_scaling* scaling, const double* lb, const double* ub, int steps,
                                         int enable_warm_start, int plant_mode, int plant_substeps, double plant_h, double* traj_x,
                                         double* traj_u, int32_t* traj_status, int32_t* traj_iters, double* traj_cost, int dev,
                                         void* stream_) {
    if (b200mpc_device_count() <= 0) return fail(B200MPC_ENOGPU, "no CUDA device: b200mpc has no CPU fallback");
    int nx, nu, np, ni, nue, rc;
    if ((rc = nl_dims(system, &nx, &nu, &np, ph, &ni, &nue))) return rc;
    if (ph < 1 || ch < 1 || ch > ph || batch < 1 || steps < 1 || !x0 || !u0 || !sys_params || !lb || !ub || !traj_x || !traj_u ||
        plant_mode < 0 || plant_mode > 2 || (plant_mode == 2 && plant_substeps < 1))
        return fail(B200MPC_EINVAL, "bad arguments");
    { int cur = 0; if (cudaGetDevice(&cur) == cudaSuccess) keep_pool(cur); }
    cudaStream_t stream = (cudaStream_t)stream_;
    const int nz = ph * nx + ch * nu + 1;
    const size_t Bn = (size_t)batch, nX = Bn * nx, nU = Bn * nu, nZ = Bn * nz;
    std::vector<void*> tofree;
    struct Free { std::vector<void*>& v; cudaStream_t s; ~Free() { for (void* p : v) cudaFreeAsync(p, s); } } freer{tofree, stream};
    auto dbuf = [&](void** p, size_t bytes) -> int { CK(cudaMallocAsync(p, bytes ? bytes : 8, stream)); tofree.push_back(*p); return 0; };
    auto in = [&](const double* h_, size_t n, const double** d) -> int {
        if (dev) { *d = h_; return 0; }
        double* p = nullptr;
        if (dbuf((void**)&p, n * sizeof(double))) return B200MPC_ECUDA;
        CK(cudaMemcpyAsync(p, h_, n * sizeof(double), cudaMemcpyHostToDevice, stream));
        *d = p; return 0;
    };
    const double *dx0, *du0, *dpar, *dlb, *dub, *dsx = nullptr, *dsu = nullptr;
    if ((rc = in(x0, nX, &dx0)) || (rc = in(u0, nU, &du0)) || (rc = in(sys_params, (size_t)(params_per_instance ? batch : 1) * np, &dpar)) ||
        (rc = in(lb, nz, &dlb)) || (rc = in(ub, nz, &dub)))
        return rc;
    if ((rc = stage_scaling(scaling, nx, nu, &dsx, &dsu, stream, tofree))) return rc;
    double *dX = traj_x, *dU = traj_u, *dC = traj_cost; int *dS = traj_status, *dI = traj_iters;
    if (!dev) {
        if ((rc = dbuf((void**)&dX, (steps + 1) * nX * 8)) || (rc = dbuf((void**)&dU, steps * nU * 8))) return rc;
        if (traj_status && (rc = dbuf((void**)&dS, steps * Bn * 4))) return rc;
        if (traj_iters && (rc = dbuf((void**)&dI, steps * Bn * 4))) return rc;
        if (traj_cost && (rc = dbuf((void**)&dC, steps * Bn * 8))) return rc;
    }
    double *zprev, *zguess, *dviol; int* dqp; int* dstat_tmp = nullptr; int* dit_tmp = nullptr; double* dcost_tmp = nullptr;
    if ((rc = dbuf((void**)&zprev, nZ * 8)) || (rc = dbuf((void**)&zguess, nZ * 8)) || (rc = dbuf((void**)&dviol, Bn * 8)) ||
        (rc = dbuf((void**)&dqp, Bn * 4)))
        return rc;
    if (!dS && (rc = dbuf((void**)&dstat_tmp, Bn * 4))) return rc;
    if (!dI && (rc = dbuf((void**)&dit_tmp, Bn * 4))) return rc;
    if (!dC && (rc = dbuf((void**)&dcost_tmp, Bn * 8))) return rc;
    CK(cudaMemcpyAsync(dX, dx0, nX * 8, cudaMemcpyDeviceToDevice, stream));
    CK(cudaMemsetAsync(zprev, 0, nZ * 8, stream));
    const unsigned gblocks = (unsigned)((nZ + 255) / 256);
    for (int k = 0; k < steps; ++k) {
        const double* xk = dX + (size_t)k * nX;
        const double* uk = k == 0 ? du0 : dU + (size_t)(k - 1) * nU;
        const int first = k == 0, cold = (first || !enable_warm_start) ? 1 : 0;
        nlmpc_guess_kernel<0><<<gblocks, 256, 0, stream>>>(batch, nx, nu, ph, ch, cold, xk, uk, zprev, first ? nullptr : zprev + (nz - 1), dlb,
                                                        dub, zguess);
        CK(cudaGetLastError());
        if ((rc = nl_solve_impl(system, ph, ch, batch, prm, zguess, xk, dpar, params_per_instance, nullptr, dsx, dsu, dlb, dub, zprev,
                                dC ? dC + (size_t)k * Bn : dcost_tmp, dviol, dS ? dS + (size_t)k * Bn : dstat_tmp,
                                dI ? dI + (size_t)k * Bn : dit_tmp, dqp, 1, stream_, false)))
            return rc;
        NlPlantArgs a{};
        a.batch = batch; a.mode = plant_mode; a.substeps = plant_substeps; a.stage = 0; a.h = plant_h;
        a.x = xk; a.u_in = nullptr; a.z = zprev; a.u_off = ph * nx; a.nz = nz; a.su = dsu; a.params = dpar;
        a.param_stride = params_per_instance ? np : 0;
        a.x_out = dX + (size_t)(k + 1) * nX; a.u_out = dU + (size_t)k * nU;
        if ((rc = nl_plant_dispatch(system, a, stream))) return rc;
    }
    if (!dev) {
        CK(cudaMemcpyAsync(traj_x, dX, (steps + 1) * nX * 8, cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(traj_u, dU, steps * nU * 8, cudaMemcpyDeviceToHost, stream));
        if (traj_status) CK(cudaMemcpyAsync(traj_status, dS, steps * Bn * 4, cudaMemcpyDeviceToHost, stream));
        if (traj_iters) CK(cudaMemcpyAsync(traj_iters, dI, steps * Bn * 4, cudaMemcpyDeviceToHost, stream));
        if (traj_cost) CK(cudaMemcpyAsync(traj_cost, dC, steps * Bn * 8, cudaMemcpyDeviceToHost, stream));
    }
    CK(cudaStreamSynchronize(stream));
    return B200MPC_OK;
}
