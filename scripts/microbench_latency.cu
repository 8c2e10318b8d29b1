// Dependent-chain latencies on one warp (cycles per op), B200.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k(double* out, long long* cyc, double seed) {
    __shared__ double sm[64];
    double a = seed + threadIdx.x * 1e-3, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    int idx = 0;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = fma(a, b, c);
    t1 = clock64(); if (threadIdx.x == 0) cyc[idx] = t1 - t0; idx++;
    // DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = a * b;
    t1 = clock64(); if (threadIdx.x == 0) cyc[idx] = t1 - t0; idx++;
    // DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = a + c;
    t1 = clock64(); if (threadIdx.x == 0) cyc[idx] = t1 - t0; idx++;
    // reciprocal chain 1.0/a
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) a = 1.0 / (a + 1.5);
    t1 = clock64(); if (threadIdx.x == 0) cyc[idx] = t1 - t0; idx++;
    // __drcp_rn chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) a = __drcp_rn(a + 1.5);
    t1 = clock64(); if (threadIdx.x == 0) cyc[idx] = t1 - t0; idx++;
    // division chain a = c2 / a
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) a = b / (a + 1.5);
    t1 = clock64(); if (threadIdx.x == 0) cyc[idx] = t1 - t0; idx++;
    // sqrt chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) a = sqrt(a + 1.5);
    t1 = clock64(); if (threadIdx.x == 0) cyc[idx] = t1 - t0; idx++;
    // smem round trip: STS -> syncwarp -> LDS (other lane's value)
    sm[threadIdx.x] = a; __syncwarp();
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) { sm[threadIdx.x] = a; __syncwarp(); a = sm[(threadIdx.x + 1) & 31]; __syncwarp(); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[idx] = t1 - t0; idx++;
    // shuffle chain (double = 2 shuffles)
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31);
    t1 = clock64(); if (threadIdx.x == 0) cyc[idx] = t1 - t0; idx++;
    // FFMA chain (float) for reference
    float f = (float)a, g = 1.0001f, h = 1e-6f;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) f = fmaf(f, g, h);
    t1 = clock64(); if (threadIdx.x == 0) cyc[idx] = t1 - t0; idx++;
    // independent DFMA x8 (throughput for one warp)
    double a0 = a, a1 = a + 1, a2 = a + 2, a3 = a + 3, a4 = a + 4, a5 = a + 5, a6 = a + 6, a7 = a + 7;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) { a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[idx] = t1 - t0; idx++;
    out[threadIdx.x] = a + f + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 64 * 8); cudaMalloc(&cyc, 16 * 8);
    const char* names[] = {"DFMA dep", "DMUL dep", "DADD dep", "1.0/x dep (+add)", "__drcp_rn dep (+add)", "b/x dep (+add)", "sqrt dep (+add)", "STS+syncwarp+LDS+syncwarp", "shfl double dep", "FFMA dep", "8 indep DFMA (per 8)"};
    for (int rep = 0; rep < 2; rep++) {
        k<<<1, 32>>>(out, cyc, 1.0); cudaDeviceSynchronize();
        long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        if (rep == 1) for (int i = 0; i < 11; i++) printf("%-28s %8.1f cycles/iter\n", names[i], (double)h[i] / N);
    }
    // multi-warp DFMA throughput per SM: 1,2,4,8,16 warps with 8 independent chains
    return 0;
}
