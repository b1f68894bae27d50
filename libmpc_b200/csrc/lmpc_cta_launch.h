// lmpc_cta_launch.h -- host interface of the CTA-per-controller LMPC engine (lmpc_cta_kernels.cuh); the kernels are
// instantiated in b200mpc_lmpc_cta_*.cu so that they compile in parallel with the warp-per-controller engine.
#pragma once
#include "capi_common.h"
#include "lmpc_cta_kernels.cuh"

namespace b200mpc {

struct CtaLaunchCfg {
    CtaLayout L;
    int threads = 0, grid = 0, quad = 0;
    size_t smem_bytes = 0;
};

// Picks the residency (factor in shared memory when it fits) and the launch geometry; B200MPC_EINVAL when the controller's
// vectors do not fit shared memory (the caller then uses the warp-per-controller engine).
int cta_configure(const Dm& d, bool quad, int device, int num_sms, int batch, int req_threads, CtaLaunchCfg* cfg);
int cta_launch(const CtaLaunchCfg& cfg, const Dm& d, const Params& p, const Prob& pr, const Out& o, int batch, double* scratch,
               int* counter, int model_shared, const int* order, double time_limit, cudaStream_t stream);

template <class DM> int cta_launch_t(const CtaLaunchCfg& cfg, const Dm& d, const Params& p, const Prob& pr, const Out& o, int batch,
                                     double* scratch, int* counter, int model_shared, const int* order, double time_limit,
                                     cudaStream_t stream);
typedef SDm<12, 4, 4, 12> DmQuadCta;

}  // namespace b200mpc
