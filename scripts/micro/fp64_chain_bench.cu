// Micro-benchmark: K1/K3 run one dependent fp64 chain per trajectory on lanes 0..2 of one warp, seven CTAs per SM.  Does a
// chain slow down when other CTAs of the SM run chains too — i.e. is the fp64 unit a per-SM resource that a 3-lane warp
// instruction occupies like a full one?   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_chain_bench fp64_chain_bench.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(64) chain(double* out, int iters, int lanes, long long* cyc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0000001, c = 1e-7;
    const long long t0 = clock64();
    if (warp == 0 && lane < lanes) {
#pragma unroll 16
        for (int i = 0; i < iters; i++) a = fma(a, b, c);  // dependent chain
    }
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
    out[blockIdx.x * 64 + threadIdx.x] = a;
}
int main() {
    const int sms = 148, iters = 6000;
    double* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(double) * 64 * sms * 16);
    cudaMalloc(&cyc, sizeof(long long) * sms * 16);
    for (int lanes : {3, 32}) {
        for (int per_sm : {1, 2, 4, 7, 10, 14}) {
            const int grid = sms * per_sm;
            chain<<<grid, 64>>>(out, iters, lanes, cyc);
            cudaDeviceSynchronize();
            cudaEvent_t a, b;
            cudaEventCreate(&a), cudaEventCreate(&b);
            cudaEventRecord(a);
            chain<<<grid, 64>>>(out, iters, lanes, cyc);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            std::vector<long long> c(grid);
            cudaMemcpy(c.data(), cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
            double mean = 0;
            for (auto v : c) mean += (double)v / grid;
            printf("active lanes %2d, %2d CTAs/SM: kernel %7.1f us, per-CTA %8.0f cycles = %5.2f cycles per dependent DFMA\n", lanes, per_sm, ms * 1e3, mean,
                   mean / iters);
        }
    }
    return 0;
}
