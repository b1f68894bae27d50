// the CTA-per-controller LMPC kernels with the quadrotor's compile-time dimensions (BASELINE configs[1], [4])
#include "lmpc_cta_launch_impl.cuh"
namespace b200mpc {
template int cta_launch_t<DmQuadCta>(const CtaLaunchCfg&, const Dm&, const Params&, const Prob&, const Out&, int, double*, int*, int, const int*,
                                     double, cudaStream_t);
}
