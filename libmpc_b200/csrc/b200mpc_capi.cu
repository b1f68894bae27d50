// b200mpc_capi.cu -- host side of the C ABI declared in include/b200mpc.h.
// Owns the device-resident controller state of a batch of LMPC instances (what ProblemBuilder + LOptimizer hold per
// object in the reference: include/mpc/LMPC/ProblemBuilder.hpp:827-857, include/mpc/LMPC/LOptimizer.hpp:518-527) and
// launches the single persistent solve kernel.  There is no CPU fallback: without a CUDA device every entry point
// that would compute returns B200MPC_ENOGPU.
#include "capi_common.h"
#include "lmpc_kernels.cuh"
#include "lmpc_cta_launch.h"

#include <cstdio>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <vector>

using namespace b200mpc;

static thread_local std::string g_err;
int b200mpc::fail(int code, const std::string& msg) { g_err = msg; return code; }

struct DevBuf {
    double* p = nullptr;
    size_t count = 0;       // doubles per instance
    bool per_instance = false;
};

struct b200mpc_lmpc {
    Dm d;
    int batch = 0, device = 0;
    cudaStream_t stream = 0;
    Params p;
    int enable_warm_start = 0;
    bool has_prev = false, model_set = false, has_iters = false;
    DevBuf A, B, C, Bd, Dd, OW, UW, DUW, XMin, XMax, YMin, YMax, UMin, UMax, SMin, SMax, SX, SU, yRef, uRef, duRef, uMeas;
    double *x0 = nullptr, *u0 = nullptr, *xnext = nullptr;
    // results
    double *cmd = nullptr, *prev_cmd = nullptr, *cost = nullptr, *seq_state = nullptr, *seq_input = nullptr, *seq_output = nullptr;
    double *sol_x = nullptr, *sol_y = nullptr;
    long long* prof = nullptr;
    int *status = nullptr, *solver_status = nullptr, *feasible = nullptr, *iters = nullptr, *rho_updates = nullptr, *polish = nullptr;
    // engine
    double* workspace = nullptr;
    size_t ws_stride = 0;
    int* counter = nullptr;
    int* order = nullptr;          // drawing order of the next solve (history ordering), [batch]
    int history_order = getenv("B200MPC_HISTORY_ORDER") ? atoi(getenv("B200MPC_HISTORY_ORDER")) : 1;
    int warps_per_cta = 0, ctas_per_sm = 0, grid = 0, num_sms = 0;
    int req_wpc = 0, req_cps = 0;
    size_t smem_cta = 0;
    bool force_generic = false;
    int gang = getenv("B200MPC_GANG") ? atoi(getenv("B200MPC_GANG")) : 1;     // B200MPC_SCHEDULE_GANG unless overridden
    int model_shared = 0;
    long long launches = 0;
    // engine 2 (CTA per controller, lmpc_cta_kernels.cuh)
    int engine_req = getenv("B200MPC_ENGINE") ? atoi(getenv("B200MPC_ENGINE")) : 0;   // 0 auto, 1 warp per controller, 2 CTA per controller
    int engine = 0;                // what configure_launch selected
    CtaLaunchCfg ctacfg;
    int req_cta_threads = getenv("B200MPC_CTA_THREADS") ? atoi(getenv("B200MPC_CTA_THREADS")) : 0;
    double time_limit = 0.0;
    std::vector<double> stage;   // host staging
};

extern "C" const char* b200mpc_last_error(void) { return g_err.c_str(); }

extern "C" int b200mpc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" void b200mpc_lmpc_default_params(b200mpc_lmpc_params* p) {
    if (!p) return;
    p->maximum_iteration = 100; p->enable_warm_start = 0;
    p->alpha = 1.6; p->rho = 1e-6; p->eps_rel = 1e-4; p->eps_abs = 1e-4; p->eps_prim_inf = 1e-3; p->eps_dual_inf = 1e-3;
    p->adaptive_rho = 1; p->polish = 1;
    p->sigma = 1e-6; p->delta = 1e-6; p->adaptive_rho_tolerance = 5.0; p->scaling = 10; p->check_termination = 25;
    p->adaptive_rho_interval = 25; p->polish_refine_iter = 3;
    p->time_limit = 0.0;
}

static int fill_const(b200mpc_lmpc* h, DevBuf& b, size_t count, double v) {
    b.count = count; b.per_instance = false;
    size_t bytes = (count ? count : 1) * sizeof(double);
    CK(cudaMalloc(&b.p, bytes));
    std::vector<double> tmp(count ? count : 1, v);
    CK(cudaMemcpy(b.p, tmp.data(), bytes, cudaMemcpyHostToDevice));
    (void)h;
    return B200MPC_OK;
}

static int upload(b200mpc_lmpc* h, DevBuf& b, const double* src, int per_instance, int dev) {
    if (!src && b.count) return fail(B200MPC_EINVAL, "null pointer");
    bool pi = per_instance != 0;
    size_t total = b.count * (pi ? (size_t)h->batch : 1);
    if (pi != b.per_instance) {
        double* np = nullptr;
        CK(cudaMalloc(&np, (total ? total : 1) * sizeof(double)));
        CK(cudaStreamSynchronize(h->stream));
        cudaFree(b.p);
        b.p = np; b.per_instance = pi;
        if (h->workspace) { cudaFree(h->workspace); h->workspace = nullptr; }   // launch geometry depends on what is shared
    }
    if (total) CK(cudaMemcpyAsync(b.p, src, total * sizeof(double), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    return B200MPC_OK;
}

// Entry points run on the handle's device and put the caller's current device back when they return.
struct DeviceScope {
    int prev = -1; bool switched = false; cudaError_t err = cudaSuccess;
    explicit DeviceScope(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) { err = cudaSetDevice(dev); switched = (err == cudaSuccess); }
    }
    ~DeviceScope() { if (switched) cudaSetDevice(prev); }
};

template <class T>
static int dalloc(T** p, size_t n) {
    CK(cudaMalloc(p, (n ? n : 1) * sizeof(T)));
    CK(cudaMemset(*p, 0, (n ? n : 1) * sizeof(T)));
    return B200MPC_OK;
}

// device state of a new handle (any failure: the caller destroys the partially built handle, nothing leaks)
static int create_alloc(b200mpc_lmpc* h, const b200mpc_lmpc_dims* dims, int batch, int device) {
    h->d.nx = dims->nx; h->d.nu = dims->nu; h->d.ndu = dims->ndu; h->d.ny = dims->ny; h->d.ph = dims->ph; h->d.ch = dims->ch;
    h->d.derive();
    h->batch = batch; h->device = device;
    b200mpc_lmpc_params dp; b200mpc_lmpc_default_params(&dp);
    const Dm& d = h->d;
    const double inf = std::numeric_limits<double>::infinity();
    int rc;
#define F(buf, cnt, val) if ((rc = fill_const(h, h->buf, (size_t)(cnt), (val))) != 0) return rc
    F(A, d.nx * d.nx, 0.0); F(B, d.nx * d.nu, 0.0); F(C, d.ny * d.nx, 0.0); F(Bd, d.nx * d.ndu, 0.0); F(Dd, d.ny * d.ndu, 0.0);
    F(OW, d.ph * d.ny, 0.0); F(UW, d.ph * d.nu, 0.0); F(DUW, d.ph * d.nu, 0.0);
    F(XMin, d.ph * d.nx, -inf); F(XMax, d.ph * d.nx, inf); F(YMin, d.ph * d.ny, -inf); F(YMax, d.ph * d.ny, inf);
    F(UMin, d.ph * d.nu, -inf); F(UMax, d.ph * d.nu, inf); F(SMin, d.ph, -inf); F(SMax, d.ph, inf);
    F(SX, d.nx, 0.0); F(SU, d.nu, 0.0);
    F(yRef, d.ph * d.ny, 0.0); F(uRef, d.ph * d.nu, 0.0); F(duRef, d.ph * d.nu, 0.0); F(uMeas, d.ph * d.ndu, 0.0);
#undef F
    size_t Bn = (size_t)batch;
    if ((rc = dalloc(&h->x0, Bn * d.nx))) return rc;
    if ((rc = dalloc(&h->u0, Bn * d.nu))) return rc;
    if ((rc = dalloc(&h->cmd, Bn * d.nu))) return rc;
    if ((rc = dalloc(&h->prev_cmd, Bn * d.nu))) return rc;
    if ((rc = dalloc(&h->cost, Bn))) return rc;
    if ((rc = dalloc(&h->seq_state, Bn * (d.ph + 1) * d.nx))) return rc;
    if ((rc = dalloc(&h->seq_input, Bn * (d.ph + 1) * d.nu))) return rc;
    if ((rc = dalloc(&h->seq_output, Bn * (d.ph + 1) * d.ny))) return rc;
    if ((rc = dalloc(&h->sol_x, Bn * d.n))) return rc;
    if ((rc = dalloc(&h->sol_y, Bn * d.m))) return rc;
    if ((rc = dalloc(&h->status, Bn))) return rc;
    if ((rc = dalloc(&h->solver_status, Bn))) return rc;
    if ((rc = dalloc(&h->feasible, Bn))) return rc;
    if ((rc = dalloc(&h->iters, Bn))) return rc;
    if ((rc = dalloc(&h->rho_updates, Bn))) return rc;
    if ((rc = dalloc(&h->polish, Bn))) return rc;
    if ((rc = dalloc(&h->counter, 1))) return rc;
    if ((rc = dalloc(&h->order, (size_t)batch))) return rc;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    h->num_sms = prop.multiProcessorCount;
    return b200mpc_lmpc_set_params(h, &dp);
}


extern "C" int b200mpc_lmpc_create(const b200mpc_lmpc_dims* dims, int batch, int device, b200mpc_lmpc_t* out) {
    if (!dims || !out || batch <= 0) return fail(B200MPC_EINVAL, "bad arguments");
    if (dims->nx < 0 || dims->nu < 0 || dims->ndu < 0 || dims->ny < 0 || dims->ph < 1 || dims->ch < 1 || dims->ch > dims->ph ||
        dims->nx + dims->nu <= 0)
        return fail(B200MPC_EINVAL, "bad dimensions");
    int ndev = b200mpc_device_count();
    if (ndev <= 0) return fail(B200MPC_ENOGPU, "no CUDA device: b200mpc has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(B200MPC_EINVAL, "bad device index");
    DeviceScope device_scope_(device);
    if (device_scope_.err != cudaSuccess) return fail(B200MPC_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(device_scope_.err));
    b200mpc_lmpc* h = new (std::nothrow) b200mpc_lmpc();
    if (!h) return fail(B200MPC_EINVAL, "out of host memory");
    h->device = device;
    const int rc = create_alloc(h, dims, batch, device);
    if (rc != B200MPC_OK) { b200mpc_lmpc_destroy(h); return rc; }
    *out = h;
    return B200MPC_OK;
}

static void free_buf(DevBuf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; }

extern "C" int b200mpc_lmpc_destroy(b200mpc_lmpc_t h) {
    if (!h) return B200MPC_OK;
    DeviceScope device_scope_(h->device);
    cudaStreamSynchronize(h->stream);
    DevBuf* bufs[] = {&h->A, &h->B, &h->C, &h->Bd, &h->Dd, &h->OW, &h->UW, &h->DUW, &h->XMin, &h->XMax, &h->YMin, &h->YMax,
                      &h->UMin, &h->UMax, &h->SMin, &h->SMax, &h->SX, &h->SU, &h->yRef, &h->uRef, &h->duRef, &h->uMeas};
    for (DevBuf* b : bufs) free_buf(*b);
    void* ptrs[] = {h->x0, h->u0, h->cmd, h->prev_cmd, h->cost, h->seq_state, h->seq_input, h->seq_output, h->sol_x, h->sol_y,
                    h->status, h->solver_status, h->feasible, h->iters, h->rho_updates, h->polish, h->counter, h->order, h->workspace, h->prof, h->xnext};
    for (void* p : ptrs) if (p) cudaFree(p);
    delete h;
    return B200MPC_OK;
}

extern "C" int b200mpc_lmpc_set_stream(b200mpc_lmpc_t h, void* stream) {
    if (!h) return fail(B200MPC_EINVAL, "null handle");
    h->stream = (cudaStream_t)stream;
    return B200MPC_OK;
}

extern "C" int b200mpc_lmpc_set_params(b200mpc_lmpc_t h, const b200mpc_lmpc_params* q) {
    if (!h || !q) return fail(B200MPC_EINVAL, "null argument");
    if (q->check_termination < 0 || q->scaling < 0 || q->adaptive_rho_interval < 0) return fail(B200MPC_EINVAL, "bad parameter");
    Params& p = h->p;
    p.max_iter = q->maximum_iteration; p.adaptive_rho = q->adaptive_rho; p.polish = q->polish; p.scaling = q->scaling;
    p.check_termination = q->check_termination; p.adaptive_rho_interval = q->adaptive_rho_interval;
    p.polish_refine_iter = q->polish_refine_iter;
    p.alpha = q->alpha; p.rho = q->rho; p.sigma = q->sigma; p.delta = q->delta; p.eps_abs = q->eps_abs; p.eps_rel = q->eps_rel;
    p.eps_prim_inf = q->eps_prim_inf; p.eps_dual_inf = q->eps_dual_inf; p.adaptive_rho_tolerance = q->adaptive_rho_tolerance;
    h->enable_warm_start = q->enable_warm_start;
    h->time_limit = q->time_limit > 0.0 ? q->time_limit : 0.0;
    return B200MPC_OK;
}

#define HCHECK()                                                    \
    if (!h) return fail(B200MPC_EINVAL, "null handle");             \
    DeviceScope device_scope_(h->device);                           \
    if (device_scope_.err != cudaSuccess) return fail(B200MPC_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(device_scope_.err))

extern "C" int b200mpc_lmpc_set_model(b200mpc_lmpc_t h, const double* A, const double* B, const double* C, int pi, int dev) {
    HCHECK();
    int rc;
    if ((rc = upload(h, h->A, A, pi, dev))) return rc;
    if ((rc = upload(h, h->B, B, pi, dev))) return rc;
    if ((rc = upload(h, h->C, C, pi, dev))) return rc;
    h->model_set = true;
    return B200MPC_OK;
}
extern "C" int b200mpc_lmpc_set_disturbances(b200mpc_lmpc_t h, const double* Bd, const double* Dd, int pi, int dev) {
    HCHECK();
    int rc;
    if ((rc = upload(h, h->Bd, Bd, pi, dev))) return rc;
    return upload(h, h->Dd, Dd, pi, dev);
}
extern "C" int b200mpc_lmpc_set_weights(b200mpc_lmpc_t h, const double* OW, const double* UW, const double* DUW, int pi, int dev) {
    HCHECK();
    int rc;
    if ((rc = upload(h, h->OW, OW, pi, dev))) return rc;
    if ((rc = upload(h, h->UW, UW, pi, dev))) return rc;
    return upload(h, h->DUW, DUW, pi, dev);
}
extern "C" int b200mpc_lmpc_set_state_bounds(b200mpc_lmpc_t h, const double* lo, const double* hi, int pi, int dev) {
    HCHECK();
    int rc;
    if ((rc = upload(h, h->XMin, lo, pi, dev))) return rc;
    return upload(h, h->XMax, hi, pi, dev);
}
extern "C" int b200mpc_lmpc_set_output_bounds(b200mpc_lmpc_t h, const double* lo, const double* hi, int pi, int dev) {
    HCHECK();
    int rc;
    if ((rc = upload(h, h->YMin, lo, pi, dev))) return rc;
    return upload(h, h->YMax, hi, pi, dev);
}

__global__ void expand_tail_kernel(double* dst, const double* src, int nu, int ch, int ph, long long copies) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = copies * ph * nu;
    if (idx >= total) return;
    long long inst = idx / ((long long)ph * nu);
    int rem = (int)(idx - inst * ph * nu);
    int col = rem / nu, r = rem - col * nu;
    int sc = col < ch ? col : ch - 1;       // ProblemBuilder.hpp:402-410
    dst[idx] = src[inst * ch * nu + (long long)sc * nu + r];
}

static int upload_input_bounds(b200mpc_lmpc* h, DevBuf& b, const double* src, int per_instance, int dev) {
    const Dm& d = h->d;
    if (!src && b.count) return fail(B200MPC_EINVAL, "null pointer");
    size_t copies = per_instance ? (size_t)h->batch : 1;
    if (!dev) {
        h->stage.resize(copies * d.ph * d.nu);
        for (size_t i = 0; i < copies; ++i)
            for (int col = 0; col < d.ph; ++col) {
                int sc = col < d.ch ? col : d.ch - 1;
                for (int r = 0; r < d.nu; ++r) h->stage[(i * d.ph + col) * d.nu + r] = src[(i * d.ch + sc) * d.nu + r];
            }
        int rc = upload(h, b, h->stage.data(), per_instance, 0);
        if (rc) return rc;
        CK(cudaStreamSynchronize(h->stream));   // staging buffer is reused
        return B200MPC_OK;
    }
    // device source: make sure the destination has the right shape, then expand on the device
    bool pi = per_instance != 0;
    if (pi != b.per_instance) {
        double* np = nullptr;
        CK(cudaMalloc(&np, (copies * b.count ? copies * b.count : 1) * sizeof(double)));
        CK(cudaStreamSynchronize(h->stream));
        cudaFree(b.p); b.p = np; b.per_instance = pi;
    }
    long long total = (long long)copies * d.ph * d.nu;
    if (total) expand_tail_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(b.p, src, d.nu, d.ch, d.ph, (long long)copies);
    CK(cudaGetLastError());
    return B200MPC_OK;
}
extern "C" int b200mpc_lmpc_set_input_bounds(b200mpc_lmpc_t h, const double* lo, const double* hi, int pi, int dev) {
    HCHECK();
    int rc;
    if ((rc = upload_input_bounds(h, h->UMin, lo, pi, dev))) return rc;
    return upload_input_bounds(h, h->UMax, hi, pi, dev);
}
// ProblemBuilder::setInputBounds(index, ...) (ProblemBuilder.hpp:469-477) writes ONE internal column and never re-replicates
// the tail: the host mirrors keep the reference's internal [nu x ph] matrices and push all ph columns through this entry.
extern "C" int b200mpc_lmpc_set_input_bounds_full(b200mpc_lmpc_t h, const double* lo, const double* hi, int pi, int dev) {
    HCHECK();
    int rc;
    if ((rc = upload(h, h->UMin, lo, pi, dev))) return rc;
    return upload(h, h->UMax, hi, pi, dev);
}
extern "C" int b200mpc_lmpc_set_scalar_constraint(b200mpc_lmpc_t h, const double* SMin, const double* SMax, const double* X,
                                                  const double* U, int pi, int dev) {
    HCHECK();
    int rc;
    if ((rc = upload(h, h->SMin, SMin, pi, dev))) return rc;
    if ((rc = upload(h, h->SMax, SMax, pi, dev))) return rc;
    if ((rc = upload(h, h->SX, X, pi, dev))) return rc;
    return upload(h, h->SU, U, pi, dev);
}
extern "C" int b200mpc_lmpc_set_references(b200mpc_lmpc_t h, const double* yRef, const double* uRef, const double* duRef, int pi, int dev) {
    HCHECK();
    int rc;
    if ((rc = upload(h, h->yRef, yRef, pi, dev))) return rc;
    if ((rc = upload(h, h->uRef, uRef, pi, dev))) return rc;
    return upload(h, h->duRef, duRef, pi, dev);
}
extern "C" int b200mpc_lmpc_set_exogenous_inputs(b200mpc_lmpc_t h, const double* uMeas, int pi, int dev) {
    HCHECK();
    return upload(h, h->uMeas, uMeas, pi, dev);
}

extern "C" int b200mpc_lmpc_set_warm_start(b200mpc_lmpc_t h, const double* primal, const double* dual, int dev) {
    HCHECK();
    if (!primal || !dual) return fail(B200MPC_EINVAL, "null pointer");
    cudaMemcpyKind k = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CK(cudaMemcpyAsync(h->sol_x, primal, (size_t)h->batch * h->d.n * sizeof(double), k, h->stream));
    CK(cudaMemcpyAsync(h->sol_y, dual, (size_t)h->batch * h->d.m * sizeof(double), k, h->stream));
    h->has_prev = true;
    return B200MPC_OK;
}
extern "C" int b200mpc_lmpc_get_warm_start(b200mpc_lmpc_t h, double* primal, double* dual, int dev) {
    HCHECK();
    cudaMemcpyKind k = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (primal) CK(cudaMemcpyAsync(primal, h->sol_x, (size_t)h->batch * h->d.n * sizeof(double), k, h->stream));
    if (dual) CK(cudaMemcpyAsync(dual, h->sol_y, (size_t)h->batch * h->d.m * sizeof(double), k, h->stream));
    if (!dev) CK(cudaStreamSynchronize(h->stream));
    return B200MPC_OK;
}

// ---- kernel instantiations: compile-time dimensions for the named workloads, runtime dimensions otherwise --------
typedef SDm<12, 4, 4, 12> DmQuad;    // quadrotor_ex / BASELINE configs[1],[4]

template <class DM>
static int configure_t(b200mpc_lmpc* h) {
    DM dm; dm.from(h->d);
    const bool mshared = !h->A.per_instance && !h->B.per_instance && !h->C.per_instance && !h->SX.per_instance && !h->SU.per_instance;
    h->model_shared = mshared ? 1 : 0;
    size_t smem_model = (size_t)dm.model_doubles() * sizeof(double);
    size_t smem_warp = (size_t)dm.smem_doubles() * sizeof(double) + (mshared ? 0 : smem_model);
    size_t smem_fixed = mshared ? smem_model : 0;
    int dev_max_smem = 0;
    CK(cudaDeviceGetAttribute(&dev_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    // warps per CTA: the request, else the value in 1..4 that maximises resident warps per SM
    int best_wpc = 0, best_occ = 0;
    for (int wpc = (h->req_wpc > 0 ? h->req_wpc : B200_MAX_THREADS / 32); wpc >= 1; --wpc) {
        size_t smem_cta = smem_warp * wpc + smem_fixed;
        if (smem_cta > (size_t)dev_max_smem) { if (h->req_wpc > 0 && wpc == h->req_wpc) continue; else continue; }
        CK(cudaFuncSetAttribute(lmpc_solve_kernel<DM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lmpc_solve_kernel<DM>, wpc * 32, smem_cta));
        if (occ * wpc > best_occ * best_wpc) { best_occ = occ; best_wpc = wpc; }
        if (h->req_wpc > 0) break;
    }
    if (best_wpc == 0 || best_occ < 1) return fail(B200MPC_EINVAL, "problem dimensions exceed the shared memory of an SM");
    int wpc = best_wpc, occ = best_occ;
    size_t smem_cta = smem_warp * wpc + smem_fixed;
    CK(cudaFuncSetAttribute(lmpc_solve_kernel<DM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
    int cps = h->req_cps > 0 ? (h->req_cps < occ ? h->req_cps : occ) : occ;
    int grid = h->num_sms * cps;
    int need = (h->batch + wpc - 1) / wpc;
    if (grid > need) grid = need;
    h->warps_per_cta = wpc; h->ctas_per_sm = cps; h->grid = grid;
    h->smem_cta = smem_cta;
    size_t wsd = (dm.ws_doubles() + 31) & ~(size_t)31;
    h->ws_stride = wsd;
    size_t slots = (size_t)grid * wpc;
    CK(cudaMalloc(&h->workspace, slots * wsd * sizeof(double)));
    return B200MPC_OK;
}
template <class DM>
static int launch_t(b200mpc_lmpc* h, const Prob& pr, const Out& o) {
    DM dm; dm.from(h->d);
    const int* order = nullptr;
    if (h->history_order && h->has_iters && h->batch > h->warps_per_cta) {      // needs a previous solve of this handle
        order_by_history_kernel<<<1, 1024, 0, h->stream>>>(h->iters, h->batch, h->p.check_termination > 0 ? h->p.check_termination : 25, h->order);
        order = h->order;
    }
    // the attribute belongs to the kernel function, not to the handle: another handle of the same instantiation may have
    // lowered it since this one was configured
    CK(cudaFuncSetAttribute(lmpc_solve_kernel<DM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_cta));
    lmpc_solve_kernel<DM><<<h->grid, h->warps_per_cta * 32, h->smem_cta, h->stream>>>(dm, h->p, pr, o, h->batch, h->workspace,
                                                                                      h->ws_stride, h->counter, h->model_shared, h->gang, order);
    CK(cudaGetLastError());
    return B200MPC_OK;
}
static bool is_quad(const Dm& d) { return d.nx == 12 && d.nu == 4 && d.ndu == 4 && d.ny == 12; }

// Engine selection.  Engine 2 (one CTA per controller, everything in shared memory: lmpc_cta_kernels.cuh) is used whenever
// the controller's vectors fit the shared memory of an SM; engine 1 (one warp per controller, state streamed through TMA rings:
// lmpc_kernels.cuh) otherwise, or when asked for (b200mpc_lmpc_set_engine / B200MPC_ENGINE).
static int configure_launch(b200mpc_lmpc* h) {
    if (h->workspace) return B200MPC_OK;
    const bool quad = !h->force_generic && is_quad(h->d);
    const bool mshared = !h->A.per_instance && !h->B.per_instance && !h->C.per_instance && !h->SX.per_instance && !h->SU.per_instance;
    // automatic choice: the CTA engine runs one controller per SM at a time (best latency, no HBM streaming); the warp engine
    // keeps 12 controllers per SM in flight and wins on throughput once the batch exceeds its resident slots (measured on the
    // quadrotor, profiles/r02_engine_crossover.jsonl: batch 1184: 17.8 vs 25.2 ms, 1776: 26.1 vs 25.4 ms, 4096: 59 vs 39 ms)
    static const int kAutoMax = getenv("B200MPC_ENGINE2_MAX_BATCH") ? atoi(getenv("B200MPC_ENGINE2_MAX_BATCH")) : 0;
    const int auto_max = kAutoMax > 0 ? kAutoMax : 11 * h->num_sms;
    if (h->engine_req == 2 || (h->engine_req == 0 && h->batch <= auto_max)) {
        CtaLaunchCfg cfg;
        int rc = cta_configure(h->d, quad, h->device, h->num_sms, h->batch, h->req_cta_threads, &cfg);
        if (rc == B200MPC_OK) {
            h->ctacfg = cfg; h->engine = 2; h->model_shared = mshared ? 1 : 0;
            CK(cudaMalloc(&h->workspace, (size_t)cfg.grid * (size_t)cfg.L.gtotal * sizeof(double)));
            h->warps_per_cta = cfg.threads / 32; h->ctas_per_sm = 1; h->grid = cfg.grid; h->smem_cta = cfg.smem_bytes;
            h->ws_stride = (size_t)cfg.L.gtotal;
            return B200MPC_OK;
        }
        if (h->engine_req == 2) return rc;
    }
    h->engine = 1;
    if (quad) return configure_t<DmQuad>(h);
    return configure_t<Dm>(h);
}
static int launch(b200mpc_lmpc* h, const Prob& pr, const Out& o) {
    if (h->engine == 2) {
        const int* order = nullptr;
        if (h->history_order && h->has_iters && h->batch > h->grid) {      // longest-expected first (LPT over the persistent CTAs)
            order_by_history_kernel<<<1, 1024, 0, h->stream>>>(h->iters, h->batch, h->p.check_termination > 0 ? h->p.check_termination : 25, h->order);
            order = h->order;
        }
        return cta_launch(h->ctacfg, h->d, h->p, pr, o, h->batch, h->workspace, h->counter, h->model_shared, order, h->time_limit, h->stream);
    }
    if (!h->force_generic && is_quad(h->d)) return launch_t<DmQuad>(h, pr, o);
    return launch_t<Dm>(h, pr, o);
}

extern "C" int b200mpc_lmpc_set_engine(b200mpc_lmpc_t h, int engine, int cta_threads) {
    HCHECK();
    if (engine < 0 || engine > 2 || (cta_threads != 0 && cta_threads != 256 && cta_threads != 384)) return fail(B200MPC_EINVAL, "bad engine");
    CK(cudaStreamSynchronize(h->stream));
    if (h->workspace) { cudaFree(h->workspace); h->workspace = nullptr; }
    h->engine_req = engine; h->req_cta_threads = cta_threads;
    return configure_launch(h);
}
extern "C" int b200mpc_lmpc_get_engine(b200mpc_lmpc_t h, int* engine, int* threads_per_cta, int* factor_in_shared_memory) {
    HCHECK();
    int rc = configure_launch(h);
    if (rc) return rc;
    if (engine) *engine = h->engine;
    if (threads_per_cta) *threads_per_cta = h->warps_per_cta * 32;
    if (factor_in_shared_memory) *factor_in_shared_memory = h->engine == 2 ? h->ctacfg.L.fac_shared : 0;
    return B200MPC_OK;
}

extern "C" int b200mpc_lmpc_set_schedule(b200mpc_lmpc_t h, int schedule) {
    if (!h || schedule < 0 || schedule > 8) return fail(B200MPC_EINVAL, "bad schedule");
    h->gang = schedule;
    return B200MPC_OK;
}
extern "C" int b200mpc_lmpc_set_history_order(b200mpc_lmpc_t h, int enable) {
    if (!h) return fail(B200MPC_EINVAL, "null handle");
    h->history_order = enable ? 1 : 0;
    return B200MPC_OK;
}
extern "C" int b200mpc_lmpc_set_launch(b200mpc_lmpc_t h, int warps_per_cta, int ctas_per_sm) {
    HCHECK();
    if (warps_per_cta < -(B200_MAX_THREADS / 32) || warps_per_cta > B200_MAX_THREADS / 32 || ctas_per_sm < 0) return fail(B200MPC_EINVAL, "bad launch geometry");
    h->force_generic = warps_per_cta < 0;   // negative: use the runtime-dimension kernel (parity testing of both instantiations)
    if (warps_per_cta < 0) warps_per_cta = -warps_per_cta;
    if (warps_per_cta > 0 || ctas_per_sm > 0) h->engine_req = 1;      // an explicit warp geometry is a request for the warp-per-controller engine
    CK(cudaStreamSynchronize(h->stream));
    if (h->workspace) { cudaFree(h->workspace); h->workspace = nullptr; }
    h->req_wpc = warps_per_cta; h->req_cps = ctas_per_sm;
    return configure_launch(h);
}

extern "C" int b200mpc_lmpc_solve(b200mpc_lmpc_t h, const double* x0, const double* u0, int dev) {
    HCHECK();
    if (!x0 || (!u0 && h->d.nu)) return fail(B200MPC_EINVAL, "null pointer");
    int rc = configure_launch(h);
    if (rc) return rc;
    const Dm& d = h->d;
    cudaMemcpyKind k = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (d.nx) CK(cudaMemcpyAsync(h->x0, x0, (size_t)h->batch * d.nx * sizeof(double), k, h->stream));
    if (d.nu) CK(cudaMemcpyAsync(h->u0, u0, (size_t)h->batch * d.nu * sizeof(double), k, h->stream));
    CK(cudaMemsetAsync(h->counter, 0, sizeof(int), h->stream));
    Prob pr;
#define A_(name) pr.name.p = h->name.p; pr.name.stride = h->name.per_instance ? (long long)h->name.count : 0
    A_(A); A_(B); A_(C); A_(Bd); A_(Dd); A_(OW); A_(UW); A_(DUW); A_(XMin); A_(XMax); A_(YMin); A_(YMax); A_(UMin); A_(UMax);
    A_(SMin); A_(SMax); A_(SX); A_(SU); A_(yRef); A_(uRef); A_(duRef); A_(uMeas);
#undef A_
    pr.x0 = h->x0; pr.u0 = h->u0;
    pr.warm_x = h->sol_x; pr.warm_y = h->sol_y;
    pr.warm = (h->enable_warm_start && h->has_prev) ? 1 : 0;   // LOptimizer.hpp:268-281
    Out o;
    o.cmd = h->cmd; o.cost = h->cost; o.status = h->status; o.solver_status = h->solver_status; o.feasible = h->feasible;
    o.iters = h->iters; o.rho_updates = h->rho_updates; o.polish = h->polish;
    o.seq_state = h->seq_state; o.seq_input = h->seq_input; o.seq_output = h->seq_output;
    o.sol_x = h->sol_x; o.sol_y = h->sol_y; o.prev_cmd = h->prev_cmd; o.prof = h->prof;
    if ((rc = launch(h, pr, o))) return rc;
    h->launches += 1;
    h->has_prev = true;   // optimal_prev_x / optimal_prev_y now hold a solution (LOptimizer.hpp:295-296)
    h->has_iters = true;
    return B200MPC_OK;
}

// ---- closed loop on the device ---------------------------------------------------------------------
// x_next = Ap x + Bp cmd for one (instance, row) per thread; records the step's command / status / iteration count.
__global__ void plant_step_kernel(int batch, int nx, int nu, const double* Ap, long long sA, const double* Bp, long long sB,
                                  const double* x, const double* cmd, double* x_next, double* u_out, const int* status, const int* iters,
                                  int* status_out, int* iters_out) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)batch * nx) return;
    int b = (int)(t / nx), r = (int)(t - (long long)b * nx);
    const double* A = Ap + (long long)b * sA + (long long)r * nx;
    const double* Bm = Bp + (long long)b * sB + (long long)r * nu;
    double acc = 0;
    for (int k = 0; k < nx; ++k) acc = fma(A[k], x[(long long)b * nx + k], acc);
    for (int k = 0; k < nu; ++k) acc = fma(Bm[k], cmd[(long long)b * nu + k], acc);
    x_next[(long long)b * nx + r] = acc;
    if (r < nu) u_out[(long long)b * nu + r] = cmd[(long long)b * nu + r];
    for (int k = nx + r; k < nu; k += nx) u_out[(long long)b * nu + k] = cmd[(long long)b * nu + k];     // nu > nx
    if (r == 0) { if (status_out) status_out[b] = status[b]; if (iters_out) iters_out[b] = iters[b]; }
}

extern "C" int b200mpc_lmpc_closed_loop(b200mpc_lmpc_t h, const double* x0, const double* u0, int steps, const double* Ap, const double* Bp,
                                        int plant_per_instance, double* traj_x, double* traj_u, int32_t* traj_status, int32_t* traj_iters,
                                        int dev) {
    HCHECK();
    if (!x0 || !u0 || steps < 1 || !traj_x || !traj_u || ((Ap == nullptr) != (Bp == nullptr))) return fail(B200MPC_EINVAL, "bad arguments");
    const Dm& d = h->d;
    const size_t Bn = (size_t)h->batch, nX = Bn * d.nx, nU = Bn * d.nu;
    std::vector<void*> tmp;
    struct Free { std::vector<void*>& v; ~Free() { for (void* p : v) cudaFree(p); } } freer{tmp};
    auto dbuf = [&](void** p, size_t bytes) -> int { CK(cudaMalloc(p, bytes ? bytes : 8)); tmp.push_back(*p); return 0; };
    double *dX = traj_x, *dU = traj_u; int *dS = traj_status, *dI = traj_iters;
    double* dU0 = nullptr;
    const double *dA = h->A.p, *dB = h->B.p;
    long long sA = h->A.per_instance ? (long long)d.nx * d.nx : 0, sB = h->B.per_instance ? (long long)d.nx * d.nu : 0;
    int rc;
    if (!dev) {
        if ((rc = dbuf((void**)&dX, (steps + 1) * nX * 8)) || (rc = dbuf((void**)&dU, steps * nU * 8))) return rc;
        if (traj_status && (rc = dbuf((void**)&dS, steps * Bn * 4))) return rc;
        if (traj_iters && (rc = dbuf((void**)&dI, steps * Bn * 4))) return rc;
    }
    if ((rc = dbuf((void**)&dU0, nU * 8))) return rc;
    cudaMemcpyKind kin = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CK(cudaMemcpyAsync(dX, x0, nX * 8, kin, h->stream));
    CK(cudaMemcpyAsync(dU0, u0, nU * 8, kin, h->stream));
    if (Ap) {
        size_t na = (size_t)d.nx * d.nx * (plant_per_instance ? Bn : 1), nb = (size_t)d.nx * d.nu * (plant_per_instance ? Bn : 1);
        if (dev) { dA = Ap; dB = Bp; }
        else {
            double *pa, *pb;
            if ((rc = dbuf((void**)&pa, na * 8)) || (rc = dbuf((void**)&pb, nb * 8))) return rc;
            CK(cudaMemcpyAsync(pa, Ap, na * 8, cudaMemcpyHostToDevice, h->stream));
            CK(cudaMemcpyAsync(pb, Bp, nb * 8, cudaMemcpyHostToDevice, h->stream));
            dA = pa; dB = pb;
        }
        sA = plant_per_instance ? (long long)d.nx * d.nx : 0; sB = plant_per_instance ? (long long)d.nx * d.nu : 0;
    }
    const unsigned blocks = (unsigned)((nX + 255) / 256);
    for (int k = 0; k < steps; ++k) {
        const double* xk = dX + (size_t)k * nX;
        const double* uk = k == 0 ? dU0 : dU + (size_t)(k - 1) * nU;
        if ((rc = b200mpc_lmpc_solve(h, xk, uk, 1))) return rc;       // IOptimizer::run on the current state, previous command
        plant_step_kernel<<<blocks, 256, 0, h->stream>>>(h->batch, d.nx, d.nu, dA, sA, dB, sB, xk, h->cmd, dX + (size_t)(k + 1) * nX,
                                                          dU + (size_t)k * nU, h->status, h->iters, dS ? dS + (size_t)k * Bn : nullptr,
                                                          dI ? dI + (size_t)k * Bn : nullptr);
        CK(cudaGetLastError());
    }
    if (!dev) {
        CK(cudaMemcpyAsync(traj_x, dX, (steps + 1) * nX * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(traj_u, dU, steps * nU * 8, cudaMemcpyDeviceToHost, h->stream));
        if (traj_status) CK(cudaMemcpyAsync(traj_status, dS, steps * Bn * 4, cudaMemcpyDeviceToHost, h->stream));
        if (traj_iters) CK(cudaMemcpyAsync(traj_iters, dI, steps * Bn * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));      // temporaries are freed on return
    return B200MPC_OK;
}

// One plant step on the device with the controller's own model: x <- A x + B cmd (in place, device pointer), u_out <- cmd.
// The piece of the examples' control loop between two optimize() calls (examples/quadrotor_ex.cpp), asynchronous on the
// handle's stream: solve / advance pairs can be enqueued back to back without a host round trip.
extern "C" int b200mpc_lmpc_advance(b200mpc_lmpc_t h, double* x_dev, double* u_out_dev) {
    HCHECK();
    if (!x_dev || !u_out_dev) return fail(B200MPC_EINVAL, "null pointer");
    const Dm& d = h->d;
    const size_t nX = (size_t)h->batch * d.nx;
    if (!h->xnext) CK(cudaMalloc(&h->xnext, (nX ? nX : 1) * sizeof(double)));
    const long long sA = h->A.per_instance ? (long long)d.nx * d.nx : 0, sB = h->B.per_instance ? (long long)d.nx * d.nu : 0;
    plant_step_kernel<<<(unsigned)((nX + 255) / 256), 256, 0, h->stream>>>(h->batch, d.nx, d.nu, h->A.p, sA, h->B.p, sB, x_dev, h->cmd, h->xnext,
                                                                          u_out_dev, h->status, h->iters, nullptr, nullptr);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(x_dev, h->xnext, nX * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return B200MPC_OK;
}

template <class T>
static int fetch(b200mpc_lmpc* h, T* dst, const T* src, size_t n, int dev) {
    if (!dst || !n) return B200MPC_OK;
    CK(cudaMemcpyAsync(dst, src, n * sizeof(T), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    return B200MPC_OK;
}

extern "C" int b200mpc_lmpc_get_result(b200mpc_lmpc_t h, double* cmd, double* cost, int32_t* status, int32_t* solver_status,
                                       int32_t* is_feasible, int32_t* iterations, int32_t* rho_updates, int32_t* status_polish, int dev) {
    HCHECK();
    size_t Bn = (size_t)h->batch;
    int rc;
    if ((rc = fetch(h, cmd, h->cmd, Bn * h->d.nu, dev))) return rc;
    if ((rc = fetch(h, cost, h->cost, Bn, dev))) return rc;
    if ((rc = fetch(h, status, h->status, Bn, dev))) return rc;
    if ((rc = fetch(h, solver_status, h->solver_status, Bn, dev))) return rc;
    if ((rc = fetch(h, is_feasible, h->feasible, Bn, dev))) return rc;
    if ((rc = fetch(h, iterations, h->iters, Bn, dev))) return rc;
    if ((rc = fetch(h, rho_updates, h->rho_updates, Bn, dev))) return rc;
    if ((rc = fetch(h, status_polish, h->polish, Bn, dev))) return rc;
    if (!dev) CK(cudaStreamSynchronize(h->stream));
    return B200MPC_OK;
}

extern "C" int b200mpc_lmpc_get_sequence(b200mpc_lmpc_t h, double* state, double* input, double* output, int dev) {
    HCHECK();
    size_t Bn = (size_t)h->batch * (h->d.ph + 1);
    int rc;
    if ((rc = fetch(h, state, h->seq_state, Bn * h->d.nx, dev))) return rc;
    if ((rc = fetch(h, input, h->seq_input, Bn * h->d.nu, dev))) return rc;
    if ((rc = fetch(h, output, h->seq_output, Bn * h->d.ny, dev))) return rc;
    if (!dev) CK(cudaStreamSynchronize(h->stream));
    return B200MPC_OK;
}

extern "C" int b200mpc_lmpc_cmd_device_ptr(b200mpc_lmpc_t h, double** cmd_dev) {
    if (!h || !cmd_dev) return fail(B200MPC_EINVAL, "null argument");
    *cmd_dev = h->cmd;
    return B200MPC_OK;
}

extern "C" int b200mpc_lmpc_info(b200mpc_lmpc_t h, int* warp_slots, size_t* ws_bytes, long long* launches) {
    HCHECK();
    int rc = configure_launch(h);
    if (rc) return rc;
    if (warp_slots) *warp_slots = h->grid * h->warps_per_cta;
    if (ws_bytes) *ws_bytes = h->ws_stride * sizeof(double);
    if (launches) *launches = h->launches;
    return B200MPC_OK;
}

// Debug/profiling aid: per-instance cycle counters of the phases of the solve kernel
// [setup+scale, factorize, admm sweeps, info/termination, polish prep, polish factor, polish solve, unpack].
extern "C" int b200mpc_lmpc_profile(b200mpc_lmpc_t h, long long* out_host) {
    HCHECK();
    size_t n = (size_t)h->batch * 16;
    if (!h->prof) { CK(cudaMalloc(&h->prof, n * sizeof(long long))); CK(cudaMemset(h->prof, 0, n * sizeof(long long))); return B200MPC_OK; }
    if (out_host) { CK(cudaStreamSynchronize(h->stream)); CK(cudaMemcpy(out_host, h->prof, n * sizeof(long long), cudaMemcpyDeviceToHost)); }
    return B200MPC_OK;
}

extern "C" int b200mpc_sync(b200mpc_lmpc_t h) {
    HCHECK();
    CK(cudaStreamSynchronize(h->stream));
    return B200MPC_OK;
}

// ---- the exchange step: one all-gather of the command block over NCCL, inside the product ------------
// The batch shards across ranks with no data-path collective; the single exchange is cmd[batch*nu] of every rank ->
// cmd_all[nranks*batch*nu] on every rank, enqueued on the handle's stream right behind the solve kernel (no host sync, no
// staging copy: the send buffer is the kernel's own output block).  NCCL is dlopen'ed (the process's already-loaded copy when
// there is one, e.g. torch's), so the library still loads on a box without it.
#include <dlfcn.h>
namespace {
struct NcclUid { char internal[128]; };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUid*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) if ((lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL))) break;      // the copy already in the process
        if (!lib) for (const char* n : names) if ((lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!lib) { err = "cannot load libnccl.so.2 (needed for the multi-GPU command all-gather)"; return false; }
#define SYM(f) f = (decltype(f))dlsym(lib, "nccl" #f); if (!f) { err = "libnccl lacks nccl" #f; lib = nullptr; return false; }
        SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(AllGather) SYM(GetErrorString)
#undef SYM
        return true;
    }
};
NcclApi g_nccl;
int ncfail(int r, const char* what) { return fail(B200MPC_ECUDA, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error")); }
}  // namespace

struct b200mpc_comm {
    void* comm = nullptr;      // ncclComm_t
    int nranks = 1, rank = 0;
    bool owned = false;
};

extern "C" int b200mpc_comm_unique_id(void* id128) {
    if (!id128) return fail(B200MPC_EINVAL, "null pointer");
    std::string err;
    if (!g_nccl.load(err)) return fail(B200MPC_ESTATE, err);
    NcclUid id;
    int r = g_nccl.GetUniqueId(&id);
    if (r) return ncfail(r, "ncclGetUniqueId");
    memcpy(id128, &id, sizeof id);
    return B200MPC_OK;
}
extern "C" int b200mpc_comm_init_rank(int nranks, int rank, const void* id128, int device, b200mpc_comm_t* out) {
    if (!id128 || !out || nranks < 1 || rank < 0 || rank >= nranks) return fail(B200MPC_EINVAL, "bad arguments");
    if (b200mpc_device_count() <= 0) return fail(B200MPC_ENOGPU, "no CUDA device: b200mpc has no CPU fallback");
    std::string err;
    if (!g_nccl.load(err)) return fail(B200MPC_ESTATE, err);
    CK(cudaSetDevice(device));
    NcclUid id;
    memcpy(&id, id128, sizeof id);
    void* comm = nullptr;
    int r = g_nccl.CommInitRank(&comm, nranks, id, rank);
    if (r) return ncfail(r, "ncclCommInitRank");
    b200mpc_comm* c = new (std::nothrow) b200mpc_comm();
    if (!c) return fail(B200MPC_EINVAL, "out of host memory");
    c->comm = comm; c->nranks = nranks; c->rank = rank; c->owned = true;
    *out = c;
    return B200MPC_OK;
}
extern "C" int b200mpc_comm_init(void* nccl_comm, int nranks, int rank, b200mpc_comm_t* out) {
    if (!nccl_comm || !out || nranks < 1 || rank < 0 || rank >= nranks) return fail(B200MPC_EINVAL, "bad arguments");
    std::string err;
    if (!g_nccl.load(err)) return fail(B200MPC_ESTATE, err);
    b200mpc_comm* c = new (std::nothrow) b200mpc_comm();
    if (!c) return fail(B200MPC_EINVAL, "out of host memory");
    c->comm = nccl_comm; c->nranks = nranks; c->rank = rank; c->owned = false;
    *out = c;
    return B200MPC_OK;
}
extern "C" int b200mpc_comm_destroy(b200mpc_comm_t c) {
    if (!c) return B200MPC_OK;
    if (c->owned && c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
    return B200MPC_OK;
}
extern "C" int b200mpc_comm_size(b200mpc_comm_t c, int* nranks, int* rank) {
    if (!c) return fail(B200MPC_EINVAL, "null communicator");
    if (nranks) *nranks = c->nranks;
    if (rank) *rank = c->rank;
    return B200MPC_OK;
}
extern "C" int b200mpc_lmpc_allgather_cmd(b200mpc_lmpc_t h, b200mpc_comm_t c, double* cmd_all_dev) {
    HCHECK();
    if (!c || !cmd_all_dev) return fail(B200MPC_EINVAL, "null argument");
    const size_t count = (size_t)h->batch * h->d.nu;
    int r = g_nccl.AllGather(h->cmd, cmd_all_dev, count, /* ncclFloat64 */ 8, c->comm, h->stream);
    if (r) return ncfail(r, "ncclAllGather");
    h->launches += 1;
    return B200MPC_OK;
}
