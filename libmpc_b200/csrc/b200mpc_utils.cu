// b200mpc_utils.cu -- set-up helpers upstream of the solve, batched on the device.
//
// mpc::discretization<nx,nu>(A, B, Ts, Ad, Bd)  (include/mpc/Utils.hpp:23-47): zero-order-hold c2d through the matrix
// exponential of the augmented matrix  M = [[A, B], [0, 0]] Ts ;  Ad = exp(M)[0:nx, 0:nx], Bd = exp(M)[0:nx, nx:nx+nu].
// The reference calls Eigen's MatrixBase::exp() once per controller; here one CTA exponentiates one (nx+nu) x (nx+nu)
// matrix in shared memory -- scaling and squaring around a degree-18 Taylor polynomial evaluated by Horner's rule
// (||M / 2^s||_1 <= 1/2, truncation 0.5^19 / 19! ~ 1e-23, far below FP64 rounding) -- so per-instance models for a whole
// batch are generated in one launch (examples/ugv_ex.cpp:36-57 does this on the host for its single controller).
#include "capi_common.h"

namespace b200mpc {

constexpr int kExpmThreads = 256;
constexpr int kTaylorDegree = 18;

// C = A * B (N x N, row-major, shared memory); all threads; ends with a barrier
__device__ __forceinline__ void mm(double* C, const double* A, const double* B, int N) {
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
        int r = e / N, c = e - r * N;
        double a0 = 0, a1 = 0;
        int k = 0;
        for (; k + 1 < N; k += 2) { a0 = fma(A[r * N + k], B[k * N + c], a0); a1 = fma(A[r * N + k + 1], B[(k + 1) * N + c], a1); }
        if (k < N) a0 = fma(A[r * N + k], B[k * N + c], a0);
        C[e] = a0 + a1;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kExpmThreads) c2d_kernel(int nx, int nu, int batch, const double* A, long long sA, const double* B,
                                                            long long sB, const double* Ts, long long sT, double* Ad, double* Bd) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double norm1;
    const int N = nx + nu;
    double* M = sm; double* E = M + N * N; double* T = E + N * N;
    for (int inst = blockIdx.x; inst < batch; inst += gridDim.x) {
        const double ts = Ts[inst * sT];
        const double* a = A + inst * sA; const double* b = B + inst * sB;
        for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
            int r = e / N, c = e - r * N;
            M[e] = r < nx ? ts * (c < nx ? a[r * nx + c] : b[r * nu + (c - nx)]) : 0.0;
        }
        __syncthreads();
        if (threadIdx.x == 0) {                        // 1-norm (max column sum): picks the scaling
            double mx = 0;
            for (int c = 0; c < N; ++c) { double s = 0; for (int r = 0; r < N; ++r) s += fabs(M[r * N + c]); mx = fmax(mx, s); }
            norm1 = mx;
        }
        __syncthreads();
        int s = 0;
        { double nrm = norm1; while (nrm > 0.5 && s < 60) { nrm *= 0.5; ++s; } }
        const double scale = ldexp(1.0, -s);
        for (int e = threadIdx.x; e < N * N; e += blockDim.x) { M[e] *= scale; E[e] = (e / N == e % N) ? 1.0 : 0.0; }
        __syncthreads();
        for (int k = kTaylorDegree; k >= 1; --k) {     // Horner: E <- I + (M / k) E
            mm(T, M, E, N);
            const double ik = 1.0 / k;
            for (int e = threadIdx.x; e < N * N; e += blockDim.x) E[e] = ((e / N == e % N) ? 1.0 : 0.0) + ik * T[e];
            __syncthreads();
        }
        for (int q = 0; q < s; ++q) {                   // undo the scaling: E <- E^2, s times
            mm(T, E, E, N);
            for (int e = threadIdx.x; e < N * N; e += blockDim.x) E[e] = T[e];
            __syncthreads();
        }
        for (int e = threadIdx.x; e < nx * N; e += blockDim.x) {
            int r = e / N, c = e - r * N;
            if (c < nx) Ad[(long long)inst * nx * nx + r * nx + c] = E[e];
            else Bd[(long long)inst * nx * nu + r * nu + (c - nx)] = E[e];
        }
        __syncthreads();
    }
}

}  // namespace b200mpc

using namespace b200mpc;

extern "C" int b200mpc_c2d(int nx, int nu, int batch, const double* A, const double* B, int model_per_instance, const double* Ts,
                           int ts_per_instance, double* Ad, double* Bd, int dev, void* stream_) {
    if (b200mpc_device_count() <= 0) return fail(B200MPC_ENOGPU, "no CUDA device: b200mpc has no CPU fallback");
    if (nx < 1 || nu < 0 || batch < 1 || !A || (!B && nu) || !Ts || !Ad || (!Bd && nu)) return fail(B200MPC_EINVAL, "bad arguments");
    const int N = nx + nu;
    const size_t smem = 3 * (size_t)N * N * sizeof(double);
    int cur = 0, maxsm = 0, sms = 0;
    CK(cudaGetDevice(&cur));
    CK(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, cur));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cur));
    if (smem > (size_t)maxsm) return fail(B200MPC_EINVAL, "nx + nu too large for the shared-memory matrix exponential");
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t nA = (size_t)nx * nx, nB = (size_t)nx * nu, cm = model_per_instance ? batch : 1, ct = ts_per_instance ? batch : 1;
    std::vector<void*> tmp;
    struct Free { std::vector<void*>& v; cudaStream_t s; ~Free() { for (void* p : v) cudaFreeAsync(p, s); } } freer{tmp, stream};
    const double *dA = A, *dB = B, *dT = Ts; double *dAd = Ad, *dBd = Bd;
    if (!dev) {
        auto up = [&](const double* h, size_t n, const double** d) -> int {
            double* p = nullptr;
            CK(cudaMallocAsync(&p, (n ? n : 1) * sizeof(double), stream)); tmp.push_back(p);
            if (n) CK(cudaMemcpyAsync(p, h, n * sizeof(double), cudaMemcpyHostToDevice, stream));
            *d = p; return 0;
        };
        int rc;
        if ((rc = up(A, nA * cm, &dA)) || (rc = up(B, nB * cm, &dB)) || (rc = up(Ts, ct, &dT))) return rc;
        CK(cudaMallocAsync(&dAd, nA * batch * sizeof(double), stream)); tmp.push_back(dAd);
        CK(cudaMallocAsync(&dBd, (nB ? nB : 1) * batch * sizeof(double), stream)); tmp.push_back(dBd);
    }
    CK(cudaFuncSetAttribute(c2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = batch < sms * 8 ? batch : sms * 8;
    c2d_kernel<<<grid, kExpmThreads, smem, stream>>>(nx, nu, batch, dA, model_per_instance ? (long long)nA : 0, dB,
                                                     model_per_instance ? (long long)nB : 0, dT, ts_per_instance ? 1 : 0, dAd, dBd);
    CK(cudaGetLastError());
    if (!dev) {
        CK(cudaMemcpyAsync(Ad, dAd, nA * batch * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (nB) CK(cudaMemcpyAsync(Bd, dBd, nB * batch * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }
    return B200MPC_OK;
}
