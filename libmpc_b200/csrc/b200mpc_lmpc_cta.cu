// the CTA-per-controller LMPC engine: configuration, dispatch, and the kernels with run-time dimensions
#include "lmpc_cta_launch_impl.cuh"
#include <cstdlib>

namespace b200mpc {

template int cta_launch_t<Dm>(const CtaLaunchCfg&, const Dm&, const Params&, const Prob&, const Out&, int, double*, int*, int, const int*, double,
                              cudaStream_t);
extern template int cta_launch_t<DmQuadCta>(const CtaLaunchCfg&, const Dm&, const Params&, const Prob&, const Out&, int, double*, int*, int,
                                            const int*, double, cudaStream_t);

int cta_configure(const Dm& d, bool quad, int device, int num_sms, int batch, int req_threads, CtaLaunchCfg* cfg) {
    int max_smem = 0;
    CK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    const size_t stat = 64;                               // static shared memory of the kernel (the drawn index)
    const int generic_scratch = quad ? 0 : 1;           // the run-time-dimension path eliminates in shared memory
    CtaLayout L = cta_layout(d, 1, generic_scratch);
    if ((size_t)L.total * sizeof(double) + stat > (size_t)max_smem) L = cta_layout(d, 0, generic_scratch);
    if ((size_t)L.total * sizeof(double) + stat > (size_t)max_smem)
        return fail(B200MPC_EINVAL, "controller vectors exceed the shared memory of an SM");
    if (d.b + d.ne >= 32768) return fail(B200MPC_EINVAL, "stage too large for the elimination table");
    cfg->L = L;
    cfg->threads = req_threads == 384 ? 384 : 256;
    cfg->quad = quad ? 1 : 0;
    cfg->smem_bytes = (size_t)L.total * sizeof(double);
    cfg->grid = batch < num_sms ? batch : num_sms;       // one CTA per SM: the whole SM works on one controller
    return B200MPC_OK;
}

int cta_launch(const CtaLaunchCfg& cfg, const Dm& d, const Params& p, const Prob& pr, const Out& o, int batch, double* scratch,
               int* counter, int model_shared, const int* order, double time_limit, cudaStream_t stream) {
    if (cfg.quad) return cta_launch_t<DmQuadCta>(cfg, d, p, pr, o, batch, scratch, counter, model_shared, order, time_limit, stream);
    return cta_launch_t<Dm>(cfg, d, p, pr, o, batch, scratch, counter, model_shared, order, time_limit, stream);
}

}  // namespace b200mpc
