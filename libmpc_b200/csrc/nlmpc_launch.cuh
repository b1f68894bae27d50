// nlmpc_launch.cuh -- host-side launch policy of the NLMPC kernels; instantiated once per built-in system in its own
// translation unit (nlmpc_sys_*.cu) so the systems compile in parallel.
#pragma once
#include "capi_common.h"
#include "nlmpc_sqp.cuh"
#include "nlmpc_launch_choice.h"

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

template <class S, int MODE, int NT>
int nl_launch_t(NlSolveArgs& a, size_t smem_per_group, int groups_per_cta, size_t gmem_doubles, int sms, cudaStream_t stream,
                std::vector<void*>& tofree) {
    auto kern = nlmpc_solve_kernel<S, MODE, NT>;
    const size_t smem = smem_per_group * groups_per_cta;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT * groups_per_cta, smem));
    if (occ < 1) occ = 1;
    int grid = (a.batch + groups_per_cta - 1) / groups_per_cta;
    if (grid > sms * occ) grid = sms * occ;
    a.mat_ws = nullptr;
    if (MODE) {
        void* ws = nullptr;
        CK(cudaMallocAsync(&ws, (size_t)grid * groups_per_cta * gmem_doubles * sizeof(double), stream));
        tofree.push_back(ws);
        a.mat_ws = (double*)ws;
    }
    kern<<<grid, NT * groups_per_cta, smem, stream>>>(a);
    CK(cudaGetLastError());
    return B200MPC_OK;
}

// Launch policy (NlWs residency modes): everything in shared memory when it fits (a warp per controller for tiny problems,
// a 4-warp CTA otherwise); else the packed KKT factor in shared memory and B / J in a per-CTA HBM workspace (8 warps);
// else everything in the workspace.
template <class S, int NT>
int nl_launch_structured_t(NlSolveArgs& a, size_t smem, int sms, cudaStream_t stream) {
    auto kern = nlmpc_structured_kernel<S, NT>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    if (occ < 1) occ = 1;
    int grid = a.batch < sms * occ ? a.batch : sms * occ;
    a.mat_ws = nullptr;
    kern<<<grid, NT, smem, stream>>>(a);          // a.counter: allocated and zeroed by the caller (b200mpc_nlmpc.cu)
    CK(cudaGetLastError());
    return B200MPC_OK;
}

template <class S>
int nl_solve_t(NlSolveArgs& a, cudaStream_t stream, std::vector<void*>& tofree) {
    int dev = 0, sms = 0, maxsm = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    {
        constexpr int K = NlIneqPerStage<S>::value > 0 ? NlIneqPerStage<S>::value : 1;
        const size_t ssm = nls_doubles(a.ph, a.ch, S::nx, S::nu, K) * sizeof(double);
        int nt = 64;
        const int choice = nl_structured_choice(nls_supported<S>(a.ph, a.ch), ssm, maxsm, a.ph * S::nx + a.ch * S::nu + 1, &nt);
        if (choice < 0) return fail(B200MPC_EINVAL, "the stage-structured NLMPC solver does not apply to this system / horizon");
        if (choice > 0) {
            if (nt == 32) return nl_launch_structured_t<S, 32>(a, ssm, sms, stream);
            if (nt == 128) return nl_launch_structured_t<S, 128>(a, ssm, sms, stream);
            return nl_launch_structured_t<S, 64>(a, ssm, sms, stream);
        }
    }
    const int n = a.ph * S::nx + a.ch * S::nu + 1, me = a.ph * S::nx;
    const int ni = S::nineq(a.ph) + nl_neq<S>(a.ph);
    auto sm = [&](int mode) { return NlWs::smem_doubles(mode, n, me, ni, a.ph, S::nx, S::nu) * sizeof(double); };
    if (sm(0) <= (size_t)maxsm) {
        if (n <= 32) return nl_launch_t<S, 0, 32>(a, sm(0), 2 * sm(0) <= (size_t)maxsm / 2 ? 2 : 1, 0, sms, stream, tofree);
        return nl_launch_t<S, 0, 128>(a, sm(0), 1, 0, sms, stream, tofree);
    }
    if (sm(1) <= (size_t)maxsm) return nl_launch_t<S, 1, 256>(a, sm(1), 1, NlWs::gmem_doubles(1, n, me, ni), sms, stream, tofree);
    if (sm(2) > (size_t)maxsm) return fail(B200MPC_EINVAL, "NLMPC problem too large: its vectors do not fit shared memory");
    return nl_launch_t<S, 2, 256>(a, sm(2), 1, NlWs::gmem_doubles(2, n, me, ni), sms, stream, tofree);
}

template <class S>
int nl_plant_t(const NlPlantArgs& a, cudaStream_t stream) {
    int grid = (a.batch + 127) / 128;
    nlmpc_plant_kernel<S><<<grid, 128, 0, stream>>>(a);
    CK(cudaGetLastError());
    return B200MPC_OK;
}

#define B200MPC_INSTANTIATE_NL_SYSTEM(S)                                                         \
    template int nl_eval_t<S>(const NlEvalArgs&, cudaStream_t);                                   \
    template int nl_solve_t<S>(NlSolveArgs&, cudaStream_t, std::vector<void*>&);          \
    template int nl_plant_t<S>(const NlPlantArgs&, cudaStream_t);

}  // namespace b200mpc
