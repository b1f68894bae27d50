// Micro-benchmarks that pin the machine constants the LMPC kernel design leans on (B200, sm_100a):
// DFMA dependent-chain latency, DFMA throughput per SM, shared-memory load latency (LDS vs generic LD),
// DFMA fed from shared memory (the inner loop shape of the sweeps).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void dfma_latency(double* out, long long* cyc, int n) {
    double a = out[0], b = out[1], c = out[2];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void dfma_throughput(double* out, int n) {
    double a0 = out[0], a1 = out[1], a2 = out[2], a3 = out[3], a4 = out[4], a5 = out[5], a6 = out[6], a7 = out[7];
    double b = out[8], c = out[9];
    for (int i = 0; i < n; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void ffma_throughput(float* out, int n) {
    float a0 = out[0], a1 = out[1], a2 = out[2], a3 = out[3], a4 = out[4], a5 = out[5], a6 = out[6], a7 = out[7];
    float b = out[8], c = out[9];
    for (int i = 0; i < n; ++i) {
        a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
        a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void lds_latency(int* out, long long* cyc, int n) {
    __shared__ int buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = (i * 7 + 1) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { p = buf[p]; p = buf[p]; p = buf[p]; p = buf[p]; }
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void dot_smem(double* out, long long* cyc, int n) {
    __shared__ double A[32 * 21];
    __shared__ double x[32];
    for (int i = threadIdx.x; i < 32 * 21; i += blockDim.x) A[i] = 1.0 / (1 + i);
    if (threadIdx.x < 32) x[threadIdx.x] = 0.5;
    __syncthreads();
    const double* row = A + threadIdx.x * 21;
    double acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < n; ++it) {
        double a0 = 0, a1 = 0;
#pragma unroll
        for (int q = 0; q < 20; q += 2) { a0 = fma(row[q], x[q], a0); a1 = fma(row[q + 1], x[q + 1], a1); }
        acc += a0 + a1;
        __syncwarp();
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* d; long long* c; CK(cudaMalloc(&d, 1 << 24)); CK(cudaMalloc(&c, 64)); CK(cudaMemset(d, 0, 1 << 24));
    long long hc; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    int n = 10000;
    dfma_latency<<<1, 32>>>(d, c, n); CK(cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost));
    printf("DFMA dependent latency: %.2f cycles\n", (double)hc / (4.0 * n));
    for (int wps = 4; wps <= 32; wps *= 2) {
        int blocks = 148 * 4, threads = wps * 32 / 4 * 1; if (threads < 32) threads = 32;
        blocks = 148; threads = wps * 32;
        dfma_throughput<<<blocks, threads>>>(d, 100);
        cudaEventRecord(e0); dfma_throughput<<<blocks, threads>>>(d, 20000); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * 20000.0 * blocks * threads;
        printf("DFMA throughput, %2d warps/SM: %.2f TFLOP/s\n", wps, fl / (ms * 1e-3) / 1e12);
    }
    {
        int blocks = 148, threads = 1024;
        ffma_throughput<<<blocks, threads>>>((float*)d, 100);
        cudaEventRecord(e0); ffma_throughput<<<blocks, threads>>>((float*)d, 20000); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA throughput, 32 warps/SM: %.2f TFLOP/s\n", 2.0 * 8 * 20000.0 * blocks * threads / (ms * 1e-3) / 1e12);
    }
    lds_latency<<<1, 32>>>((int*)d, c, n); CK(cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost));
    printf("LDS dependent latency: %.2f cycles\n", (double)hc / (4.0 * n));
    dot_smem<<<1, 32>>>(d, c, n); CK(cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost));
    printf("20-term smem dot product (1 warp, unrolled, 2 acc) + syncwarp: %.1f cycles\n", (double)hc / n);
    return 0;
}
