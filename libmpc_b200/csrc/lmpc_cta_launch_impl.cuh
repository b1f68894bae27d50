// lmpc_cta_launch_impl.cuh -- launch of one dimension policy of the CTA-per-controller engine (included by b200mpc_lmpc_cta_*.cu)
#pragma once
#include "lmpc_cta_launch.h"

namespace b200mpc {

template <class DM, int NT, bool FSH>
static int cta_launch_k(const CtaLaunchCfg& cfg, const DM& dm, const Params& p, const Prob& pr, const Out& o, int batch, double* scratch,
                        int* counter, int model_shared, const int* order, double time_limit, cudaStream_t stream) {
    auto kern = lmpc_cta_kernel<DM, NT, FSH>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes));
    kern<<<cfg.grid, NT, cfg.smem_bytes, stream>>>(dm, p, pr, o, cfg.L, batch, scratch, counter, model_shared, order, time_limit);
    CK(cudaGetLastError());
    return B200MPC_OK;
}

template <class DM>
int cta_launch_t(const CtaLaunchCfg& cfg, const Dm& d, const Params& p, const Prob& pr, const Out& o, int batch, double* scratch,
                 int* counter, int model_shared, const int* order, double time_limit, cudaStream_t stream) {
    DM dm; dm.from(d);
#define B200_CTA_CASE(NT_)                                                                                                         \
    if (cfg.threads == NT_) return cfg.L.fac_shared                                                                                \
        ? cta_launch_k<DM, NT_, true>(cfg, dm, p, pr, o, batch, scratch, counter, model_shared, order, time_limit, stream)         \
        : cta_launch_k<DM, NT_, false>(cfg, dm, p, pr, o, batch, scratch, counter, model_shared, order, time_limit, stream);
    B200_CTA_CASE(256)
    B200_CTA_CASE(384)
#undef B200_CTA_CASE
    return fail(B200MPC_EINVAL, "unsupported CTA size");
}

}  // namespace b200mpc
