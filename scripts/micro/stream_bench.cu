// Micro-benchmark: how fast can ONE CTA stream a private region of HBM, as a function of the method and of how many CTAs do
// so at once?  (What bounds the L-BFGS direction update: 256..1024 CTAs, each reading its own 2 x 259 KB history twice.)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_bench stream_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../fast_racing_b200/csrc/bulk_copy.cuh"
using namespace frx;

#define THREADS 256
// method 0: plain 8-byte loads, 8 independent per thread per trip
__global__ void __launch_bounds__(THREADS) k_ldg(const double* __restrict__ base, size_t per_cta, double* out, long long* cyc) {
    const double* p = base + (size_t)blockIdx.x * per_cta;
    const long long t0 = clock64();
    double acc = 0.0;
    for (size_t i = threadIdx.x; i + 7 * THREADS < per_cta; i += 8 * THREADS) {
        double v[8];
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = p[i + e * THREADS];
#pragma unroll
        for (int e = 0; e < 8; e++) acc += v[e];
    }
    out[blockIdx.x * THREADS + threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}
// method 1: 16-byte loads, 8 independent per thread per trip
__global__ void __launch_bounds__(THREADS) k_ldg128(const double* __restrict__ base, size_t per_cta, double* out, long long* cyc) {
    const double2* p = (const double2*)(base + (size_t)blockIdx.x * per_cta);
    const size_t n2 = per_cta / 2;
    const long long t0 = clock64();
    double acc = 0.0;
    for (size_t i = threadIdx.x; i + 7 * THREADS < n2; i += 8 * THREADS) {
        double2 v[8];
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = p[i + e * THREADS];
#pragma unroll
        for (int e = 0; e < 8; e++) acc += v[e].x + v[e].y;
    }
    out[blockIdx.x * THREADS + threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}
// method 2: bulk copies into a ring of NST stages of CH doubles
template <int NST, int CHMAX>
__global__ void __launch_bounds__(THREADS) k_bulk(const double* __restrict__ base, size_t per_cta, double* out, long long* cyc, int CH = CHMAX,
                                                  int off = 0) {
    extern __shared__ __align__(128) double ring[];
    __shared__ unsigned long long full[NST];
    const double* p = base + (size_t)blockIdx.x * per_cta + off;
    const int Q = (int)((per_cta - off) / CH);
    const int tid = threadIdx.x;
    if (tid == 0)
        for (int i = 0; i < NST; i++) mbar_init(full + i, 1);
    __syncthreads();
    const long long t0 = clock64();
    if (tid == 0)
        for (int q = 0; q < NST && q < Q; q++) {
            mbar_expect_tx(full + q, CH * 8);
            bulk_g2s(ring + (size_t)q * CHMAX, p + (size_t)q * CH, CH * 8, full + q);
        }
    double acc = 0.0;
    for (int q = 0; q < Q; q++) {
        const int stg = q % NST;
        mbar_wait(full + stg, (q / NST) & 1);
        for (int i = tid; i < CH; i += THREADS) acc += ring[(size_t)stg * CHMAX + i];
        __syncthreads();
        if (tid == 0 && q + NST < Q) {
            mbar_expect_tx(full + stg, CH * 8);
            bulk_g2s(ring + (size_t)stg * CHMAX, p + (size_t)(q + NST) * CH, CH * 8, full + stg);
        }
    }
    out[blockIdx.x * THREADS + tid] = acc;
    __syncthreads();
    if (tid == 0) cyc[blockIdx.x] = clock64() - t0;
}

template <class F>
static void run(const char* name, F launch, int ctas, size_t per_cta, long long* cyc_d) {
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    launch();  // warm
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    std::vector<long long> c(ctas);
    cudaMemcpy(c.data(), cyc_d, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
    double mean = 0;
    long long mx = 0;
    for (auto v : c) mean += (double)v / ctas, mx = v > mx ? v : mx;
    const cudaError_t e = cudaGetLastError();
    printf("%-22s ctas %5d  per-CTA %7.0f KB  kernel %8.1f us  %7.0f GB/s   per-CTA cycles mean %9.0f max %9lld  %s\n", name, ctas,
           per_cta * 8 / 1024.0, ms * 1e3, ctas * per_cta * 8 / (ms * 1e6), mean, mx, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    const size_t per_cta = 128 * 253 * 2;  // doubles: S and Y of one C4 trajectory (518 KB)
    const int max_ctas = 2048;
    double *buf, *out;
    long long* cyc;
    cudaMalloc(&buf, sizeof(double) * per_cta * max_ctas);
    cudaMemset(buf, 0, sizeof(double) * per_cta * max_ctas);
    cudaMalloc(&out, sizeof(double) * THREADS * max_ctas);
    cudaMalloc(&cyc, sizeof(long long) * max_ctas);
    cudaFuncSetAttribute(k_bulk<4, 2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 2048 * 8);
    cudaFuncSetAttribute(k_bulk<8, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 1024 * 8);
    cudaFuncSetAttribute(k_bulk<2, 8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8192 * 8);
    cudaFuncSetAttribute(k_bulk<16, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 512 * 8);
    cudaFuncSetAttribute(k_bulk<6, 4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 4096 * 8);
    for (int ctas : {1, 256, 1024}) {
        run("ldg64 x8", [&] { k_ldg<<<ctas, THREADS>>>(buf, per_cta, out, cyc); }, ctas, per_cta, cyc);
        run("ldg128 x8", [&] { k_ldg128<<<ctas, THREADS>>>(buf, per_cta, out, cyc); }, ctas, per_cta, cyc);
        run("bulk 4 x 16 KB", [&] { k_bulk<4, 2048><<<ctas, THREADS, 4 * 2048 * 8>>>(buf, per_cta, out, cyc); }, ctas, per_cta, cyc);
        run("bulk 8 x 8 KB", [&] { k_bulk<8, 1024><<<ctas, THREADS, 8 * 1024 * 8>>>(buf, per_cta, out, cyc); }, ctas, per_cta, cyc);
        run("bulk 16 x 4 KB", [&] { k_bulk<16, 512><<<ctas, THREADS, 16 * 512 * 8>>>(buf, per_cta, out, cyc); }, ctas, per_cta, cyc);
        run("bulk 2 x 64 KB", [&] { k_bulk<2, 8192><<<ctas, THREADS, 2 * 8192 * 8>>>(buf, per_cta, out, cyc); }, ctas, per_cta, cyc);
        run("bulk 6 x 32 KB", [&] { k_bulk<6, 4096><<<ctas, THREADS, 6 * 4096 * 8>>>(buf, per_cta, out, cyc); }, ctas, per_cta, cyc);
        run("bulk 4x16192B", [&] { k_bulk<4, 2048><<<ctas, THREADS, 4 * 2048 * 8>>>(buf, per_cta, out, cyc, 2024, 0); }, ctas, per_cta, cyc);
        run("bulk 4x16192B +16B", [&] { k_bulk<4, 2048><<<ctas, THREADS, 4 * 2048 * 8>>>(buf, per_cta, out, cyc, 2024, 2); }, ctas, per_cta, cyc);
        run("bulk 4x16384B +16B", [&] { k_bulk<4, 2048><<<ctas, THREADS, 4 * 2048 * 8>>>(buf, per_cta, out, cyc, 2048, 2); }, ctas, per_cta, cyc);
        run("bulk 4x16384B +64B", [&] { k_bulk<4, 2048><<<ctas, THREADS, 4 * 2048 * 8>>>(buf, per_cta, out, cyc, 2048, 8); }, ctas, per_cta, cyc);
        printf("\n");
    }
    return 0;
}
