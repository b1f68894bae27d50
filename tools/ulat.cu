// Dependent-issue latencies on one warp (cycles per operation, clock64 around a chain of N dependent operations): what bounds the
// stage recurrences and the pivot loop of the LMPC / NLMPC kernels.   nvcc -arch=sm_100a -O3 -o ulat tools/ulat.cu && ./ulat
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 512;
__global__ void lat(double* out, long long* cyc, double seed, int* chase) {
    __shared__ int sidx[64];
    __shared__ double sval[64];
    const int lane = threadIdx.x;
    sidx[lane] = chase[lane]; sidx[lane + 32] = chase[lane + 32]; sval[lane] = seed + lane; sval[lane + 32] = seed - lane;
    __syncwarp();
    double x = seed + lane * 1e-3, y = 1.0 + 1e-9 * lane, acc = 0;
    long long t0, t1;
#define TIME(slot, body) t0 = clock64(); _Pragma("unroll 16") for (int i = 0; i < N; ++i) { body; } t1 = clock64(); if (lane == 0) cyc[slot] = t1 - t0; acc += x;
    TIME(0, x = fma(x, y, 1e-9))                               // DFMA
    TIME(1, x = x + y)                                         // DADD
    TIME(2, x = x * y)                                         // DMUL
    TIME(3, x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31))  // SHFL of a double (2 x SHFL.32)
    TIME(4, x = 1.0 / x + 1.5)                                 // IEEE division (+ add)
    TIME(5, x = rsqrt(fabs(x) + 1.0))                          // rsqrt
    TIME(6, x = sqrt(fabs(x) + 1.0))                           // sqrt
    int j = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) j = sidx[j];                   // dependent LDS.32
    t1 = clock64(); if (lane == 0) cyc[7] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { sval[lane] = x; __syncwarp(); x = sval[(lane + 1) & 31] + 1e-9; __syncwarp(); }   // STS -> barrier -> LDS -> DADD
    t1 = clock64(); if (lane == 0) cyc[8] = t1 - t0;
    float f = (float)seed + lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) f = fmaf(f, 1.0000001f, 1e-9f);
    t1 = clock64(); if (lane == 0) cyc[9] = t1 - t0;
    out[lane] = acc + x + j + f;
}
int main() {
    double* out; long long* cyc; int* chase; int h[64];
    for (int i = 0; i < 64; ++i) h[i] = (i * 17 + 5) % 64;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 16 * 8); cudaMalloc(&chase, 64 * 4);
    cudaMemcpy(chase, h, sizeof h, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) lat<<<1, 32>>>(out, cyc, 1.000001, chase);
    long long c[16]; cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    const char* names[] = {"DFMA", "DADD", "DMUL", "SHFL.f64", "1.0/x+add", "rsqrt", "sqrt", "LDS chase", "STS+sync+LDS+DADD", "FFMA"};
    printf("{");
    for (int k = 0; k < 10; ++k) printf("\"%s\": %.1f%s", names[k], (double)c[k] / N, k < 9 ? ", " : "");
    printf("}\n");
    return 0;
}
