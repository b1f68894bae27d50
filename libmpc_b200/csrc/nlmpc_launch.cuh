// nlmpc_launch.cuh -- host-side launch policy of the NLMPC kernels; instantiated once per built-in system in its own
// translation unit (nlmpc_sys_*.cu) so the systems compile in parallel.
#pragma once
#include "capi_common.h"
#include "nlmpc_sqp.cuh"

namespace b200mpc {

template <class S>
int nl_eval_t(const NlEvalArgs& a, cudaStream_t stream) {
    int wpb = 4;
    size_t smem = (size_t)wpb * (a.ph + 1) * (S::nx + S::nu) * sizeof(double);
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int grid = (a.batch + wpb - 1) / wpb;
    if (grid > sms * 8) grid = sms * 8;
    CK(cudaFuncSetAttribute(nlmpc_eval_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nlmpc_eval_kernel<S><<<grid, wpb * 32, smem, stream>>>(a);
    CK(cudaGetLastError());
    return B200MPC_OK;
}

template <class S, bool GM, int NT>
int nl_launch_t(NlSolveArgs& a, size_t smem_per_group, int groups_per_cta, size_t mat_doubles, int sms, cudaStream_t stream,
                       std::vector<void*>& tofree) {
    auto kern = nlmpc_solve_kernel<S, GM, NT>;
    const size_t smem = smem_per_group * groups_per_cta;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT * groups_per_cta, smem));
    if (occ < 1) occ = 1;
    int grid = (a.batch + groups_per_cta - 1) / groups_per_cta;
    if (grid > sms * occ) grid = sms * occ;
    a.mat_ws = nullptr;
    if (GM) {
        void* ws = nullptr;
        CK(cudaMalloc(&ws, (size_t)grid * groups_per_cta * mat_doubles * sizeof(double)));
        tofree.push_back(ws);
        a.mat_ws = (double*)ws;
    }
    kern<<<grid, NT * groups_per_cta, smem, stream>>>(a);
    CK(cudaGetLastError());
    return B200MPC_OK;
}

// Launch policy: matrices in shared memory when they fit (a warp per controller for tiny problems, a 4-warp CTA otherwise),
// else an 8-warp CTA per controller with the matrices in a per-CTA HBM workspace.
template <class S>
int nl_solve_t(NlSolveArgs& a, cudaStream_t stream, std::vector<void*>& tofree) {
    int dev = 0, sms = 0, maxsm = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int n = a.ph * S::nx + a.ch * S::nu + 1, me = a.ph * S::nx;
    const int ni = S::nineq(a.ph);
    const size_t vecb = NlWs::vec_doubles(n, me, ni, a.ph, S::nx, S::nu) * sizeof(double);
    const size_t matb = NlWs::mat_doubles(n, me, ni, false) * sizeof(double);
    if (vecb + matb <= (size_t)maxsm) {
        if (n <= 32) return nl_launch_t<S, false, 32>(a, vecb + matb, 2 * (vecb + matb) <= (size_t)maxsm / 2 ? 2 : 1, 0, sms, stream, tofree);
        return nl_launch_t<S, false, 128>(a, vecb + matb, 1, 0, sms, stream, tofree);
    }
    if (vecb > (size_t)maxsm) return fail(B200MPC_EINVAL, "NLMPC problem too large: its vectors do not fit shared memory");
    return nl_launch_t<S, true, 256>(a, vecb, 1, NlWs::mat_doubles(n, me, ni, true), sms, stream, tofree);
}


#define B200MPC_INSTANTIATE_NL_SYSTEM(S)                                                         \
    template int nl_eval_t<S>(const NlEvalArgs&, cudaStream_t);                                   \
    template int nl_solve_t<S>(NlSolveArgs&, cudaStream_t, std::vector<void*>&);

}  // namespace b200mpc
