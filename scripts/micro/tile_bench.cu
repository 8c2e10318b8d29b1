// Micro-benchmark: one CTA streams 256 rows x ld doubles (the S and Y history of one trajectory, row-major) as 8-row tiles:
//   (a) eight 1-D bulk copies per tile (one per row segment),  (b) one 2-D tensor-map copy per tile,
//   (c) one 1-D bulk copy of 2 whole rows (reference: the pass-2 pattern).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tile_bench tile_bench.cu
#include <cstdio>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../fast_racing_b200/csrc/bulk_copy.cuh"
using namespace frx;
#define THREADS 256
#define NST 3

__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* map, int c0, int r0, unsigned long long* bar) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst), a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(d),
                 "l"(map), "r"(c0), "r"(r0), "r"(a)
                 : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(THREADS) k_tiles(const double* __restrict__ base, const CUtensorMap* __restrict__ map, int ld, int W, int rows_per_cta,
                                                   double* out, long long* cyc) {
    extern __shared__ __align__(128) double ring[];
    __shared__ unsigned long long full[NST];
    const int tid = threadIdx.x;
    const int CT = (ld + W - 1) / W;
    const int Q = MODE == 2 ? rows_per_cta / 2 : (rows_per_cta / 8) * CT;
    const int CH = MODE == 2 ? 2 * ld : 8 * (W + 4);
    const size_t row0 = (size_t)blockIdx.x * rows_per_cta;
    if (tid == 0)
        for (int i = 0; i < NST; i++) mbar_init(full + i, 1);
    __syncthreads();
    auto issue = [&](int q) {
        const int stg = q % NST;
        double* dst = ring + (size_t)stg * CH;
        if (MODE == 2) {
            mbar_expect_tx(full + stg, 2u * ld * 8u);
            bulk_g2s(dst, base + (row0 + 2 * q) * ld, 2u * ld * 8u, full + stg);
        } else {
            const int ct = q % CT, rb = q / CT, c0 = ct * W, wt = min(W, ld - c0);
            if (MODE == 0) {
                mbar_expect_tx(full + stg, 8u * wt * 8u);
                for (int r = 0; r < 8; r++) bulk_g2s(dst + r * (W + 4), base + (row0 + rb * 8 + r) * ld + c0, wt * 8u, full + stg);
            } else {
                mbar_expect_tx(full + stg, 8u * W * 8u);  // the full box is counted, out-of-bounds columns are zero-filled
                tma2d(dst, map, c0, (int)(row0 + rb * 8), full + stg);
            }
        }
    };
    const long long t0 = clock64();
    if (tid == 0)
        for (int q = 0; q < NST && q < Q; q++) issue(q);
    double acc = 0.0;
    for (int q = 0; q < Q; q++) {
        const int stg = q % NST;
        mbar_wait(full + stg, (q / NST) & 1);
        for (int i = tid; i < (MODE == 2 ? 2 * ld : 8 * W); i += THREADS) acc += ring[(size_t)stg * CH + i];
        __syncthreads();
        if (tid == 0 && q + NST < Q) issue(q + NST);
    }
    out[blockIdx.x * THREADS + tid] = acc;
    __syncthreads();
    if (tid == 0) cyc[blockIdx.x] = clock64() - t0;
}

template <class F>
static void run(const char* name, F launch, int ctas, size_t bytes_per_cta, long long* cyc_d) {
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    launch();
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
    for (auto v : c) mean += (double)v / ctas;
    const cudaError_t e = cudaGetLastError();
    printf("%-28s ctas %5d  kernel %8.1f us  %7.0f GB/s   per-CTA cycles mean %9.0f  %s\n", name, ctas, ms * 1e3, ctas * bytes_per_cta / (ms * 1e6), mean,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    const int ld = 1192, rows = 256, max_ctas = 1024;
    double *buf, *out;
    long long* cyc;
    cudaMalloc(&buf, sizeof(double) * (size_t)ld * rows * max_ctas);
    cudaMemset(buf, 0, sizeof(double) * (size_t)ld * rows * max_ctas);
    cudaMalloc(&out, sizeof(double) * THREADS * max_ctas);
    cudaMalloc(&cyc, sizeof(long long) * max_ctas);
    typedef CUresult (*encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    encode_t encode = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qr) != cudaSuccess || !encode) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    for (int W : {244, 124}) {
        CUtensorMap hmap;
        const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)rows * max_ctas};
        const cuuint64_t gstr[1] = {(cuuint64_t)ld * 8};
        const cuuint32_t box[2] = {(cuuint32_t)W, 8};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = encode(&hmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("W = %d: encode -> %d\n", W, (int)r);
        CUtensorMap* dmap;
        cudaMalloc(&dmap, sizeof(CUtensorMap));
        cudaMemcpy(dmap, &hmap, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
        const size_t smem = sizeof(double) * NST * (size_t)(8 * (W + 4) > 2 * ld ? 8 * (W + 4) : 2 * ld);
        cudaFuncSetAttribute(k_tiles<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_tiles<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_tiles<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int ctas : {1, 256, 1024}) {
            const size_t bpc = (size_t)ld * rows * 8;
            run("8 x 1-D copies per tile", [&] { k_tiles<0><<<ctas, THREADS, smem>>>(buf, dmap, ld, W, rows, out, cyc); }, ctas, bpc, cyc);
            run("one 2-D tensor copy per tile", [&] { k_tiles<1><<<ctas, THREADS, smem>>>(buf, dmap, ld, W, rows, out, cyc); }, ctas, bpc, cyc);
            run("one 1-D copy of 2 rows", [&] { k_tiles<2><<<ctas, THREADS, smem>>>(buf, dmap, ld, W, rows, out, cyc); }, ctas, bpc, cyc);
        }
        printf("\n");
    }
    return 0;
}
