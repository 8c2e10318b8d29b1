// The three kernels of one batched cost+gradient evaluation (SE3GCOPTER::objectiveFunc,
// se3gcopter_cpu.hpp "CPU.hpp" :961-1000), hand-written for sm_100a, fp64 throughout:
//
//   K1 minco_forward_kernel   x=(tau,xi) -> T, q -> banded MINCO system -> LU -> c      (one 2-warp CTA / trajectory:
//                             pivot warp + feeder warp over a shared-memory ring of the band)
//        forwardT CPU.hpp:626, splitToFineT :930, forwardP :729, MINCO_S3::generate :425,
//        BandedSystem::factorizeLU/solve trajectory.hpp:655/692
//   K2 penalty_kernel         c -> per-piece penalty integral, dJ/dc, dJ/dT, jerk energy  (samples x pieces x batch)
//        MINCO_S3::addTimeIntPenalty CPU.hpp:188-408, getTrajJerkCost :507, addGradJbyT :65, addGradJbyC :80
//   K3 minco_backward_kernel  dJ/dc -> adjoint solve -> dJ/dT, dJ/dq -> dJ/d(tau,xi), cost (one 2-warp CTA / trajectory:
//                             sweep warp + cp.async producer warp streaming the factors)
//        BandedSystem::solveAdj trajectory.hpp:724, addPropCtoT CPU.hpp:104, addPropCtoP :154,
//        mergeToCoarseGradT :946, addLayerTGrad :816, addLayerPGrad :897, rho terms :988-989
//
// No tensor cores: the largest contraction on this path is 6x3 . 3x3.  See DESIGN.md for the
// roofline of each kernel.
#include <cuda_runtime.h>
#include <math.h>
#include <cstdlib>
#include <mutex>

#include "device_batch.cuh"
#include "bulk_copy.cuh"

#ifdef FRX_EXACT
#define FRX_IMPL frx_exact  // reference-order build: -DFRX_EXACT --fmad=false (see build.py)
#else
#define FRX_IMPL frx_fast
#endif

namespace FRX_IMPL {
using namespace frx;

// ------------------------------------------------------------------------------------------------
// Band storage used on the device: row-major, 16 doubles per row: slot (j - i + 6) for column j of
// row i (13 diagonals, bandwidth 6/6 as trajectory.hpp:645 with N-major order swapped for coalesced
// row access).  After the factorisation slots 0..5 hold the multipliers l_{i,i-6..i-1}, slot 6 the RECIPROCAL of
// the U diagonal, 7..12 the upper band of U, 13..15 the forward-substituted right-hand side.
// ------------------------------------------------------------------------------------------------
#define BW 16

__device__ __forceinline__ double c2_forward(double t) {  // CPU.hpp:639-641
    return t > 0.0 ? ((0.5 * t + 1.0) * t + 1.0) : 1.0 / ((0.5 * t - 1.0) * t + 1.0);
}

// 1/x without the slow-path branches of the compiler's division: hardware seed + two Newton steps (~1 ulp).
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// Same accuracy in three dependent operations instead of four: one cubically convergent step y (1 + e + e^2), e = 1 - x y
// (the hardware seed is good to ~2^-23, so e^3 is below the rounding of a double).  Used on the 2 x 2 pivot chain of K1.
__device__ __forceinline__ double fast_rcp3(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double p = fma(e, e, e);
    return fma(y, p, y);
}

// CTA-wide barrier usable from warp-specialised branches (both warps of the 64-thread CTA arrive at barrier 1).
// (barrier.sync, not bar.sync = barrier.sync.aligned: warp 0 reaches it from lanes that have not reconverged)
__device__ __forceinline__ void cta_bar64() { asm volatile("barrier.sync 1, 64;" ::: "memory"); }

// Phase clocks (debug builds only: FRX_NVCC_EXTRA=-DFRX_PHASE_CLK): lane 0 of each warp of one CTA records clock64()
// at phase boundaries; scripts/phase_clocks.py prints the differences.  Slot = kernel*32 + warp*16 + mark.
__device__ long long g_phase_clk[96];
// Per-CTA wall spans (debug builds): globaltimer (ns) and SM id at the first and last PHASE_MARK of K1 (kernel 0) / K3 (2),
// for every CTA of the grid: shows launch skew, waves and per-SM crowding (scripts/cta_spans.py).
__device__ unsigned long long g_cta_span[2][3][4096];
#ifdef FRX_PHASE_CLK
__device__ __forceinline__ void cta_span_mark(int kernel, int which) {
    if (threadIdx.x == 0 && blockIdx.x < 4096) {
        unsigned long long t;
        unsigned sm;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        g_cta_span[kernel ? 1 : 0][which][blockIdx.x] = t;
        if (which == 0) g_cta_span[kernel ? 1 : 0][2][blockIdx.x] = sm;
    }
}
#define CTA_SPAN(kernel, which) cta_span_mark(kernel, which)
#define PHASE_MARK(kernel, mark)                                                                          \
    do {                                                                                                  \
        if (blockIdx.x == gridDim.x / 2 && (threadIdx.x & 31) == 0)                                       \
            g_phase_clk[(kernel)*32 + (threadIdx.x >> 5) * 16 + (mark)] = clock64();                      \
    } while (0)
#define PHASE_ACC_DECL long long acc_t0_ = clock64(), acc_work_ = 0, acc_wait_ = 0
#define PHASE_ACC_WORK()                   \
    do {                                   \
        const long long n_ = clock64();    \
        acc_work_ += n_ - acc_t0_;         \
        acc_t0_ = n_;                      \
    } while (0)
#define PHASE_ACC_WAIT()                   \
    do {                                   \
        const long long n_ = clock64();    \
        acc_wait_ += n_ - acc_t0_;         \
        acc_t0_ = n_;                      \
    } while (0)
#define PHASE_ACC_STORE(kernel)                                                       \
    do {                                                                              \
        if (blockIdx.x == gridDim.x / 2 && (threadIdx.x & 31) == 0) {                 \
            g_phase_clk[(kernel)*32 + (threadIdx.x >> 5) * 16 + 8] = acc_work_;       \
            g_phase_clk[(kernel)*32 + (threadIdx.x >> 5) * 16 + 9] = acc_wait_;       \
        }                                                                             \
    } while (0)
#else
#define PHASE_ACC_DECL
#define PHASE_ACC_WORK() \
    do {                 \
    } while (0)
#define PHASE_ACC_WAIT() \
    do {                 \
    } while (0)
#define PHASE_ACC_STORE(kernel) \
    do {                        \
    } while (0)
#define PHASE_MARK(kernel, mark) \
    do {                         \
    } while (0)
#define CTA_SPAN(kernel, which) \
    do {                        \
    } while (0)
#endif

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int NPEND>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(NPEND));
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#ifdef FRX_EXACT  // reference-order build only: the banded system in the reference's own elimination order
// K1 ---------------------------------------------------------------------------------------------
// One trajectory per CTA of two warps.
//   warp 0  runs the dependent chain: the 6M-1 pivots of the banded LU (trajectory.hpp:655-687) with the
//           forward substitution of the three right-hand sides fused in (:695-705), on a 32-row ring of the
//           band in shared memory; then the back substitution (:706-717).
//   warp 1  keeps the chain fed: it fills the rows of the junction two pieces ahead from a constant entry
//           table (MINCO_S3::generate, CPU.hpp:437-499), and retires finished rows: LU factors to global
//           memory for the adjoint solve of K3, (1/u_ii, U row, y) into a compact shared store for the back
//           substitution.  The warps meet at one barrier per piece (6 pivots).
// Only ~(4 + 0.5 M) KB of shared memory per trajectory, so a 1024-trajectory batch is resident in one wave.
//
// Ring row layout (16 doubles): [0..5] multipliers l_{i,i-6..i-1} | [6] 1/u_ii once final (the diagonal
// before) | [7..12] u_{i,i+1..i+6} | [13..15] right-hand side (x,y,z).  Rows are retired to bv.lu in the
// same layout.
#define RING 32
#define UST 10  // reference-order build: u_ii, u_{i,i+1..i+6}, y_{x,y,z} (division kept in the back substitution)

// Entry tables of one junction block (rows 6i+3 .. 6i+8, CPU.hpp:446-476): value = coef * T^pow at
// (row offset, band slot = column - row + 6).
__device__ const int JR[31] = {0, 0, 0, 0, 1, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5};
__device__ const int JS[31] = {6, 7, 8, 12, 6, 7, 12, 1, 2, 3, 4, 5, 6, 0, 1, 2, 3, 4, 5, 6, 0, 1, 2, 3, 4, 6, 0, 1, 2, 3, 6};
__device__ const double JC[31] = {6.0, 24.0, 60.0, -6.0, 24.0, 120.0, -24.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0,
                              1.0, 1.0, -1.0, 1.0, 2.0, 3.0, 4.0, 5.0, -1.0, 2.0, 6.0, 12.0, 20.0, -2.0};
__device__ const int JP[31] = {0, 1, 2, 0, 0, 1, 0, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5, 0, 0, 1, 2, 3, 4, 0, 0, 1, 2, 3, 0};
// Tail block (rows 6M-3 .. 6M-1, CPU.hpp:481-495), T of the last piece.
__device__ const int TR[15] = {0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2};
__device__ const int TS[15] = {3, 4, 5, 6, 7, 8, 3, 4, 5, 6, 7, 3, 4, 5, 6};
__device__ const double TC[15] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 2.0, 3.0, 4.0, 5.0, 2.0, 6.0, 12.0, 20.0};
__device__ const int TP[15] = {0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 0, 1, 2, 3};

__device__ __forceinline__ double pick_pow(int p, double T1, double T2, double T3, double T4, double T5) {
    return p == 0 ? 1.0 : (p == 1 ? T1 : (p == 2 ? T2 : (p == 3 ? T3 : (p == 4 ? T4 : T5))));
}

// Per-lane entry of the fill tables (lane e writes entry e of a junction / tail block).
struct FillLane {
    int jr, js, jp, tr, ts, tp;
    double jc, tc;
};

// Fill the rows of block `g` into the ring, by one full warp.  g = -1: head rows 0..2; 0 <= g <= M-2: junction g
// (duration Tg, waypoint qv on lanes 0..2); g = M-1: tail rows (duration Tg = T of the last piece); g = M: six zero
// rows after the matrix (so that the pivot loop needs no bounds predicates).
__device__ __forceinline__ void fill_block(double* ring, const BatchView& bv, int b, int M, int g, int lane, const FillLane& fl,
                                           double Tg, double qv) {
    const int N = 6 * M;
    if (g < -1 || g > M) return;
    const int row0 = (g < 0) ? 0 : ((g >= M - 1) ? (g == M ? N : N - 3) : 6 * g + 3);
    const int nrows = (g < 0 || g == M - 1) ? 3 : 6;
    for (int e = lane; e < nrows * BW; e += 32) ring[((row0 + e / BW) & (RING - 1)) * BW + (e % BW)] = 0.0;
    __syncwarp();
    if (g < 0) {
        const double* ht = bv.headtail + 18 * (size_t)b;
        if (lane < 3) ring[(lane & (RING - 1)) * BW + 6] = (lane == 2) ? 2.0 : 1.0;  // A(0,0)=1, A(1,1)=1, A(2,2)=2
        if (lane < 9) ring[((lane / 3) & (RING - 1)) * BW + 13 + lane % 3] = ht[lane];  // rows 0,1,2 <- head p, v, a
    } else if (g < M - 1) {
        const double T1 = Tg, T2 = T1 * T1, T3 = T2 * T1, T4 = T2 * T2, T5 = T4 * T1;
        if (lane < 31) ring[((row0 + fl.jr) & (RING - 1)) * BW + fl.js] = fl.jc * pick_pow(fl.jp, T1, T2, T3, T4, T5);
        if (lane < 3) ring[((row0 + 2) & (RING - 1)) * BW + 13 + lane] = qv;  // waypoint row 6g+5
    } else if (g == M - 1) {
        const double T1 = Tg, T2 = T1 * T1, T3 = T2 * T1, T4 = T2 * T2, T5 = T4 * T1;
        const double* ht = bv.headtail + 18 * (size_t)b + 9;
        if (lane < 15) ring[((row0 + fl.tr) & (RING - 1)) * BW + fl.ts] = fl.tc * pick_pow(fl.tp, T1, T2, T3, T4, T5);
        if (lane < 9) ring[((row0 + lane / 3) & (RING - 1)) * BW + 13 + lane % 3] = ht[lane];
    }
    __syncwarp();
}

// Retire rows [row0, row0+cnt): factors to global (same 16-double layout), compact store for back substitution.
__device__ __forceinline__ void retire_rows(const double* ring, double* lug, double* ust, int row0, int cnt, int lane) {
    for (int e = lane; e < cnt * BW; e += 32) lug[(size_t)row0 * BW + e] = ring[((row0 + e / BW) & (RING - 1)) * BW + (e % BW)];
    for (int e = lane; e < cnt * UST; e += 32) {
        const int rr = e / UST, q = e - UST * rr;
        const double* rw = ring + ((row0 + rr) & (RING - 1)) * BW;
        ust[(size_t)(row0 + rr) * UST + q] = rw[6 + q];  // diagonal, upper band, y: unscaled
    }
}

// smem: ring[RING][16] | scratch[32] | ust[N][UST] (head aliased by Tc[coarseN] during the prologue, tail by the T/q staging)
__global__ void __launch_bounds__(64) minco_forward_kernel(BatchView bv, const double* __restrict__ xin) {
    extern __shared__ double sm[];
    const int b = bv.slot_list ? bv.slot_list[blockIdx.x] : bv.sub_traj0 + blockIdx.x;
    if (b < 0) return;  // empty solver slot
    if (!bv.slot_list && bv.active && !bv.active[b]) return;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int p0 = bv.piece_off[b];
    const int M = bv.piece_off[b + 1] - p0;
    const int N = 6 * M;
    const int co0 = bv.coarse_off[b];
    const int coarseN = bv.coarse_off[b + 1] - co0;
    const DevParams par = bv.par[b];
    const double* x = xin + bv.x_off[b];
    const int dimT = bv.dimT[b];
    double* ring = sm;
    double* scratch = ring + RING * BW;
    double* ust = scratch + 32;
    double* Tc = ust;  // coarse durations: dead before the first row is retired into ust
    PHASE_MARK(0, 0);

    // waypoint index tables (free dimension, xi offset, vertex offset) staged next to Tc: one round trip instead of one per use
    int* wpi = reinterpret_cast<int*>(ust + M + 1);
    for (int i = tid; i < M - 1; i += 64) {
        wpi[i] = bv.wp_vert_cnt[p0 + i] - 1;
        wpi[M + i] = bv.wp_xi_off[p0 + i];
        wpi[2 * M + i] = bv.wp_vert_off[p0 + i];
    }
    // ---- forwardT (CPU.hpp:626-676) ----
    if (par.softT) {
        for (int i = tid; i < coarseN; i += 64) Tc[i] = par.c2 ? c2_forward(x[i]) : exp(x[i]);
    } else if (tid == 0) {  // fixed total time: sequential sums as in the reference (rare path)
        const int Ms1 = dimT;
        for (int i = 0; i < Ms1; i++) Tc[i] = par.c2 ? c2_forward(x[i]) : exp(x[i]);
        Tc[Ms1] = 0.0;
        double s = 0.0;
        for (int i = 0; i <= Ms1; i++) s += Tc[i];
        const double den = 1.0 + s;
        for (int i = 0; i <= Ms1; i++) Tc[i] /= den;
        s = 0.0;
        for (int i = 0; i <= Ms1; i++) s += Tc[i];
        Tc[Ms1] = 1.0 - s;
        for (int i = 0; i <= Ms1; i++) Tc[i] *= par.sumT;
    }
    __syncthreads();
    // ---- splitToFineT (CPU.hpp:930-944) + forwardP (CPU.hpp:729-747) ----
    // T_g and q_g are also staged in the tail of `ust` for the ring fills: entry g is consumed at phase g-3, long before
    // the retired rows (6*UST doubles per phase from the front) reach the tail.
    double* stage = ust + (size_t)(6 * UST - 4) * M;
    for (int i = tid; i < M; i += 64) {
        const int pl = bv.piece_poly[p0 + i];
        const double Tv = Tc[pl] / bv.intervals[co0 + pl];
        bv.T[p0 + i] = Tv;
        stage[4 * i] = Tv;
    }
    for (int i = tid; i < M - 1; i += 64) {  // one thread per waypoint, sums in the reference's order
        const int k = bv.wp_vert_cnt[p0 + i] - 1;
        const double* V = bv.verts + 3 * (size_t)bv.wp_vert_off[p0 + i];
        const double* xi = x + dimT + bv.wp_xi_off[p0 + i];
        double sn = 0.0;
#pragma unroll 8
        for (int a = 0; a < k; a++) sn += xi[a] * xi[a];
        const double sc = 2.0 / (1.0 + sn);
        double qx = 0.0, qy = 0.0, qz = 0.0;
#pragma unroll 4
        for (int a = 0; a < k; a++) {
            const double r = sc * xi[a];
            const double rr = r * r;
            qx += V[3 * (a + 1) + 0] * rr;
            qy += V[3 * (a + 1) + 1] * rr;
            qz += V[3 * (a + 1) + 2] * rr;
        }
        const double qv[3] = {qx + V[0], qy + V[1], qz + V[2]};
        for (int d = 0; d < 3; d++) bv.q[3 * (size_t)(p0 + i) + d] = stage[4 * i + 1 + d] = qv[d];
    }
    __syncthreads();  // the fills read the staged T and q
    PHASE_MARK(0, 1);

    double* lug = bv.lu + (size_t)BW * 6 * p0;
    FillLane fl;
    {
        const int e = lane < 31 ? lane : 0, t = lane < 15 ? lane : 0;
        fl.jr = JR[e], fl.js = JS[e], fl.jp = JP[e], fl.jc = JC[e];
        fl.tr = TR[t], fl.ts = TS[t], fl.tp = TP[t], fl.tc = TC[t];
    }
    auto loadT = [&](int g) { return (g >= 0 && g < M) ? stage[4 * g] : 1.0; };
    auto loadq = [&](int g) { return (lane < 3 && g >= 0 && g < M - 1) ? stage[4 * g + 1 + lane] : 0.0; };
    // ---- ring prologue: head rows, block 0, block 1 ----
    if (warp == 0) {
        fill_block(ring, bv, b, M, -1, lane, fl, 1.0, 0.0);
        fill_block(ring, bv, b, M, 0, lane, fl, loadT(0), loadq(0));
    } else {
        fill_block(ring, bv, b, M, 1, lane, fl, loadT(1), loadq(1));
    }
    __syncthreads();
    __syncthreads();

    PHASE_MARK(0, 2);
    // ---- phases: piece j = pivots 6j .. 6j+5 ----
    if (warp == 0) {
        const int r = lane / 3, cg = lane - 3 * r;
        constexpr unsigned RB = RING * BW * 8;  // ring size in bytes (power of two)
        char* const rb = reinterpret_cast<char*>(ring);
        unsigned ok = 0;                               // byte offset of pivot row k
        unsigned oi = ((1 + r) * BW * 8) & (RB - 1);   // byte offset of row k+1+r
        const unsigned cA = (5 - r) * 8, cB = (6 + cg - r) * 8, cC = (13 + cg) * 8, cK = cg * 8;
        double* pend = scratch + lane;
        double l_pend = 0.0;
        const bool act = lane < 18;
        for (int j = 0; j < M; j++) {
            const int kend = min(6 * j + 6, N - 1);
            for (int k = 6 * j; k < kend; k++) {
                if (act) {
                    const double* rk = reinterpret_cast<const double*>(rb + ok);
                    const double* rk2 = reinterpret_cast<const double*>(rb + ok + cK);
                    double* pa = reinterpret_cast<double*>(rb + oi + cA);
                    double* pb = reinterpret_cast<double*>(rb + oi + cB);
                    double* pc = reinterpret_cast<double*>(rb + oi + cC);
                    const double l = *pa / rk[6];  // a_ik / u_kk as the reference (trajectory.hpp:667)
                    const double v0 = pb[0] - l * rk2[7];
                    const double v1 = pb[3] - l * rk2[10];
                    const double v2 = *pc - l * rk2[13];
                    pb[0] = v0;
                    pb[3] = v1;
                    *pc = v2;
                    *pend = l_pend;  // multiplier of the PREVIOUS step (nobody reads column k-1 any more)
                    pend = (cg == 0) ? pa : scratch + lane;
                    l_pend = l;
                }
                ok = (ok + BW * 8) & (RB - 1);
                oi = (oi + BW * 8) & (RB - 1);
                __syncwarp();
            }
            if (j == M - 1 && act) *pend = l_pend;
            cta_bar64();
        }
    } else {
        // The feeder's per-phase work is straight-line code on per-lane constants (K1 is issue-bound with seven CTAs per SM):
        //   fill     slot e = lane + 32 t (t < 3) of the 6 x 16 junction block holds  fc[t] * T^fp[t]  (0 for non-entries),
        //            except the waypoint row's right-hand side (row 2, columns 13..15 = lanes 13..15 at t = 1);
        //   retire   48 sixteen-byte chunks of factors to global, and UST compact entries per row for the back substitution.
        double fc[3];
        int fp[3];
#pragma unroll
        for (int t = 0; t < 3; t++) {
            const int e = lane + 32 * t;
            fc[t] = 0.0, fp[t] = 0;
            for (int q = 0; q < 31; q++)
                if (JR[q] * BW + JS[q] == e) fc[t] = JC[q], fp[t] = JP[q];
        }
        const int ue1 = lane + 32;
        const int ur0 = lane / UST, uq0 = lane - UST * ur0, ur1 = ue1 / UST, uq1 = ue1 - UST * ur1;
        const bool u1on = ue1 < 6 * UST;
        auto loadq13 = [&](int g) { return (lane >= 13 && lane < 16 && g >= 0 && g < M - 1) ? stage[4 * g + 1 + (lane - 13)] : 0.0; };
        double Tn = loadT(2), qn = loadq13(2);  // operands of the fill of phase 0
        PHASE_ACC_DECL;
        for (int j = 0; j < M; j++) {
            const double Tc_ = Tn, qc_ = qn;
            Tn = loadT(j + 3), qn = loadq13(j + 3);  // prefetch for the next phase
            const int g = j + 2;
            if (g < M - 1) {  // junction block g: rows 6g+3 .. 6g+8
                const int row0 = 6 * g + 3;
                const double T1 = Tc_, T2 = T1 * T1, T3 = T2 * T1, T4 = T2 * T2, T5 = T4 * T1;
#pragma unroll
                for (int t = 0; t < 3; t++) {
                    const int e = lane + 32 * t;
                    double v = (fc[t] != 0.0) ? fc[t] * pick_pow(fp[t], T1, T2, T3, T4, T5) : 0.0;  // structural zeros stay exact
                    if (t == 1 && lane >= 13 && lane < 16) v = qc_;  // waypoint row 6g+5, right-hand side
                    ring[((row0 + (e >> 4)) & (RING - 1)) * BW + (e & 15)] = v;
                }
            } else {
                fill_block(ring, bv, b, M, g, lane, fl, Tc_, 0.0);  // tail rows / zero rows / nothing: once each
            }
            if (j >= 1) {
                const int row0 = 6 * (j - 1);
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    const int c = lane + 32 * t;  // 16-byte chunk of the 6 x 16 block
                    if (c < 48)
                        *reinterpret_cast<double2*>(lug + (size_t)row0 * BW + 2 * c) =
                            *reinterpret_cast<const double2*>(ring + ((row0 + (c >> 3)) & (RING - 1)) * BW + 2 * (c & 7));
                }
                {
                    const double* rw = ring + ((row0 + ur0) & (RING - 1)) * BW;
                    ust[(size_t)(row0 + ur0) * UST + uq0] = rw[6 + uq0];
                }
                if (u1on) {
                    const double* rw = ring + ((row0 + ur1) & (RING - 1)) * BW;
                    ust[(size_t)(row0 + ur1) * UST + uq1] = rw[6 + uq1];
                }
            }
            PHASE_ACC_WORK();
            cta_bar64();
            PHASE_ACC_WAIT();
        }
        PHASE_ACC_STORE(0);
    }
    PHASE_MARK(0, 3);
    // ---- retire the last six rows ----
    if (warp == 1) retire_rows(ring, lug, ust, 6 * (M - 1), 6, lane);
    __syncthreads();
    PHASE_MARK(0, 4);

    // ---- back substitution U x = y (trajectory.hpp:706-717), row-oriented: lane d < 3 of warp 0 owns
    //      right-hand side d and keeps the last six results in registers ----
    if (warp == 0 && lane < 3) {
        const int d = lane;
        double* cg_ = bv.c + 18 * (size_t)p0;
        double w[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (int base = N - 6; base >= 0; base -= 6) {
#pragma unroll
            for (int u = 5; u >= 0; u--) {
                const int i = base + u;
                const double* Ui = ust + (size_t)i * UST;
                double acc = Ui[7 + d];
#pragma unroll
                for (int t = 6; t >= 1; t--) acc -= Ui[t] * w[(u + t) % 6];
                acc = acc / Ui[0];
                w[u] = acc;
                cg_[3 * i + d] = acc;
            }
        }
    }
    PHASE_MARK(0, 5);
}
#endif  // FRX_EXACT (banded kernels)

// K2 ---------------------------------------------------------------------------------------------
// Per-sample work of MINCO_S3::addTimeIntPenalty (CPU.hpp:188-408), restructured around ADJOINT
// 3-vectors: every penalty's dJ/dc is  beta0 (x) g0 + beta1 (x) g1 + beta2 (x) g2 + beta3 (x) g3  with
// g_d = adjoint of (pos, vel, acc, jer) at the sample, and the time gradient is
// alpha (g0.vel + g1.acc + g2.jer + g3.sna) + the direct v^3/K terms.  The Jacobians of the body
// frame w.r.t. the acceleration (dxB, dyB, dzB of CPU.hpp:268-278) are never materialised: with
//   dzB = (I - z z^T)/f,  dyB = (I - y y^T) C (I - z z^T)/(cn f),  C z = (0, z_z, -z_y),  x = y x z
// their transposes are applied matrix-free:  dxB^T a = dyB^T (z x a) + dzB^T (a x y),
//   dzB^T v = (v - z (z.v))/f ,   dyB^T v = (I - z z^T) C^T (I - y y^T) v / (cn f).
// Same mathematics as the reference, different association (inside the stated tolerance).
//
// smem: cf[ppc][60] derivative coefficient tables | fc[ppc][6*max_faces] | part[20][threads]
#define PEN_THREADS 64
#define CF 60  // per piece: 18 pos + 15 vel + 12 acc + 9 jerk + 6 snap coefficients

#ifndef FRX_PEN_MINBLOCKS
#define FRX_PEN_MINBLOCKS 8  // CTAs per SM the register allocation aims at (8 x 64 threads x 128 registers = the whole file)
#endif
__global__ void __launch_bounds__(PEN_THREADS, FRX_PEN_MINBLOCKS) penalty_kernel(BatchView bv, PenaltyLaunch pl) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    int gp_base, npieces;
    if (bv.slot_list) {  // slot mode: CTA = (solver slot, group of ppc pieces of that slot's trajectory)
        const int groups = (bv.max_pieces + pl.ppc - 1) / pl.ppc;
        const int slot = blockIdx.x / groups, grp = blockIdx.x - slot * groups;
        const int b = bv.slot_list[slot];
        if (b < 0) return;
        gp_base = bv.piece_off[b] + grp * pl.ppc;
        npieces = min(pl.ppc, bv.piece_off[b + 1] - gp_base);
        if (npieces <= 0) return;
    } else {
        gp_base = bv.sub_piece0 + blockIdx.x * pl.ppc;
        npieces = min(pl.ppc, bv.sub_piece1 - gp_base);
    }
    double* cf = sm;
    double* fc = cf + (size_t)pl.ppc * CF;
    double* part = fc + (size_t)pl.ppc * 6 * bv.max_faces;
    float4* ff = reinterpret_cast<float4*>(part + 20 * PEN_THREADS);  // [ppc][max_faces] fp32 (n, d_lo) for the prefilter
    double* Ts = reinterpret_cast<double*>(ff + (size_t)pl.ppc * bv.max_faces);  // [ppc] piece durations

    // per-thread operands, requested before the staging so that their latency overlaps it
    const int pl_piece = tid / pl.tpp;
    const int chunk = tid - pl_piece * pl.tpp;
    bool mine = pl_piece < npieces;
    int my_traj = 0, my_F = 0, my_K = 1;
    if (mine) {
        const int gp = gp_base + pl_piece;
        my_traj = bv.piece_traj[gp];
        my_F = bv.piece_face_cnt[gp];
        my_K = bv.cons[gp];
    }
    if (tid < npieces) Ts[tid] = bv.T[gp_base + tid];
    // ---- stage: face lists (16-byte async copies, every piece of the CTA in flight at once) and polynomial
    //      coefficient tables (c and its first four derivatives, per axis) ----
    {
        const int per = 3 * bv.max_faces;  // 16-byte chunks per piece slot (a face row = 48 B)
        for (int i = tid; i < npieces * per; i += PEN_THREADS) {
            const int p = i / per, e = i - p * per;
            const int gp = gp_base + p;
            const int cnt = bv.piece_face_cnt[gp], off = bv.piece_face_off[gp];
            if (e < 3 * cnt) cp_async16(fc + (size_t)p * 6 * bv.max_faces + 2 * e, bv.faces + 6 * (size_t)off + 2 * e);
            if (e < cnt) cp_async16(ff + (size_t)p * bv.max_faces + e, bv.faces_f32 + 4 * ((size_t)off + e));
        }
        cp_async_commit();
    }
    for (int i = tid; i < npieces * 18; i += PEN_THREADS) {
        const int p = i / 18, e = i - 18 * p, k = e / 3, d = e - 3 * k;
        const double v = bv.c[18 * (size_t)gp_base + i];
        double* t = cf + (size_t)p * CF;
        t[3 * k + d] = v;                                                                     // pos:  c_k
        if (k >= 1) t[18 + 3 * (k - 1) + d] = v * (double)k;                                  // vel:  k c_k
        if (k >= 2) t[33 + 3 * (k - 2) + d] = v * (double)(k * (k - 1));                      // acc
        if (k >= 3) t[45 + 3 * (k - 3) + d] = v * (double)(k * (k - 1) * (k - 2));            // jerk
        if (k >= 4) t[54 + 3 * (k - 4) + d] = v * (double)(k * (k - 1) * (k - 2) * (k - 3));  // snap
    }
#pragma unroll
    for (int v = 0; v < 20; v++) part[v * PEN_THREADS + tid] = 0.0;
    cp_async_wait<0>();
    __syncthreads();

    double* acc = part + tid;  // acc[v * PEN_THREADS]

    if (mine && !(bv.active && !bv.active[my_traj])) {
        const DevParams* __restrict__ P = bv.par + my_traj;
        const double* t = cf + (size_t)pl_piece * CF;
        const double* faces = fc + (size_t)pl_piece * 6 * bv.max_faces;
        const float4* faces32 = ff + (size_t)pl_piece * bv.max_faces;
        const int F = my_F;
        const int K = my_K;
        const int S = K + 1;
        const double step = Ts[pl_piece] / K;
        const double invK = 1.0 / K;
        const int j0 = chunk * pl.spt;
        const int j1 = min(j0 + pl.spt, S);
        // running-sum abscissa exactly as the reference: s1 = 0; s1 += step (CPU.hpp:246,400)
        double s1 = 0.0;
        for (int j = 0; j < j0 && j < S; j++) s1 += step;
        const double ex = P->ell[0], ey = P->ell[1], ez = P->ell[2];
        const double emax = fmax(ex, fmax(ey, ez));
        const float reach32 = (float)(P->margin + emax + 1.0e-9) + 2.0e-6f;  // fp32 prefilter: margin + ellipsoid reach, rounded up
        const double chi0 = P->chi[0], chi1 = P->chi[1], chi2 = P->chi[2], chi3 = P->chi[3];
        const double margin = P->margin, gacc = P->gacc;
        const double vmax_sqr = P->vmax_sqr, thrmin_sqr = P->thrmin_sqr, thrmax_sqr = P->thrmax_sqr, bdrmax_sqr = P->bdrmax_sqr;

        for (int j = j0; j < j1; j++) {
            const double alpha = invK * j;
            const double omg = (j == 0 || j == S - 1) ? 0.5 : 1.0;
            const double w = omg * step;
            // Ordered to keep the live register set small: position + acceleration + frame first, then the face loop
            // (only pos, the frame and the SFC accumulators are live), then velocity / jerk and the V, T, B penalties.
            double pos[3], ac[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                pos[d] = ((((t[15 + d] * s1 + t[12 + d]) * s1 + t[9 + d]) * s1 + t[6 + d]) * s1 + t[3 + d]) * s1 + t[d];
                ac[d] = ((t[42 + d] * s1 + t[39 + d]) * s1 + t[36 + d]) * s1 + t[33 + d];
            }
            // body frame (CPU.hpp:266-279)
            const double hx = ac[0], hy = ac[1], hz = ac[2] + gacc;
            const double hsq = hx * hx + hy * hy + hz * hz;
            const double fThr = sqrt(hsq);
            const double inv_f = 1.0 / fThr;
            const double zx = hx * inv_f, zy = hy * inv_f, zz = hz * inv_f;
            const double cn = sqrt(zz * zz + zy * zy);
            const double inv_cn = 1.0 / cn;
            const double yy = zz * inv_cn, yz = -zy * inv_cn;                 // yB = (0, yy, yz)
            const double xx = yy * zz - yz * zy, xy = yz * zx, xz = -yy * zx;  // xB = yB x zB

            double g0[3] = {0, 0, 0};
            double uy[3] = {0, 0, 0}, uz[3] = {0, 0, 0};  // arguments of dyB^T and dzB^T
            double cst = 0.0, direct = 0.0;
            bool any = false, jac = false;

            // ---- safe-flight-corridor faces (CPU.hpp:310-345) ----
            if (chi0 != 0.0) {  // chi0 == 0 adds exact zeros in the reference
                double ax[3] = {0, 0, 0}, ay[3] = {0, 0, 0}, az[3] = {0, 0, 0};  // sum coef * e_m * n
                bool sfc_on = false;
                // Two passes per group of 32 faces: (1) every lane tests all faces (cheap, convergent) and records the
                // near ones in a bit mask; (2) each lane walks its own near faces, so a warp iterates max-over-lanes
                // near-face counts instead of once per face that is near for ANY lane.  Ascending face order is kept.
                for (int k0 = 0; k0 < F; k0 += 32) {
                    const int kn = min(32, F - k0);
                    unsigned near = 0u;
                    {   // conservative fp32 test on the (otherwise idle) fp32 pipe: t >= n.pos + margin + emax + 1e-9 and
                        // d_lo <= n.p in exact arithmetic, so t < d_lo proves the fp64 early-out  nd + margin + emax < -1e-9;
                        // everything else goes to the fp64 body, whose own  sd > 0  test decides (results unchanged).
                        const float px = (float)pos[0], py = (float)pos[1], pz = (float)pos[2];
                        const float bias = reach32 + 1.0e-6f * (fabsf(px) + fabsf(py) + fabsf(pz));
                        for (int k = 0; k < kn; k++) {
                            const float4 f = faces32[k0 + k];
                            const float t = fmaf(f.x, px, fmaf(f.y, py, fmaf(f.z, pz, bias)));
                            if (!(t < f.w)) near |= 1u << k;
                        }
                    }
                    while (near) {
                        const double* fk = faces + 6 * (k0 + __ffs(near) - 1);
                        near &= near - 1u;
                        const double2 f01 = *reinterpret_cast<const double2*>(fk);
                        const double2 f23 = *reinterpret_cast<const double2*>(fk + 2);
                        const double2 f45 = *reinterpret_cast<const double2*>(fk + 4);
                        const double nx = f01.x, ny = f01.y, nz = f23.x;
                        const double nd = nx * (pos[0] - f23.y) + ny * (pos[1] - f45.x) + nz * (pos[2] - f45.y);
                        const double ux = (xx * nx + xy * ny + xz * nz) * ex;
                        const double uyv = (yy * ny + yz * nz) * ey;
                        const double uzv = (zx * nx + zy * ny + zz * nz) * ez;
                        const double eNorm = sqrt(ux * ux + uyv * uyv + uzv * uzv);
                        const double sd = nd + eNorm + margin;
                        if (sd > 0.0) {
                            const double sd2 = sd * sd, sd3 = sd * sd2;
                            const double coef = w * chi0 * 3.0 * sd2;
                            const double ie = 1.0 / eNorm;
                            const double e0 = coef * (ux * ie * ex), e1 = coef * (uyv * ie * ey), e2 = coef * (uzv * ie * ez);
                            g0[0] += coef * nx, g0[1] += coef * ny, g0[2] += coef * nz;
                            ax[0] += e0 * nx, ax[1] += e0 * ny, ax[2] += e0 * nz;
                            ay[0] += e1 * nx, ay[1] += e1 * ny, ay[2] += e1 * nz;
                            az[0] += e2 * nx, az[1] += e2 * ny, az[2] += e2 * nz;
                            cst += w * chi0 * sd3;
                            direct += omg * chi0 * sd3 * invK;
                            sfc_on = true;
                        }
                    }
                }
                if (sfc_on) {
                    // dxB^T ax + dyB^T ay + dzB^T az = dyB^T (ay + z x ax) + dzB^T (az + ax x y)
                    uy[0] = ay[0] + (zy * ax[2] - zz * ax[1]);
                    uy[1] = ay[1] + (zz * ax[0] - zx * ax[2]);
                    uy[2] = ay[2] + (zx * ax[1] - zy * ax[0]);
                    uz[0] = az[0] + (ax[1] * yz - ax[2] * yy);
                    uz[1] = az[1] + (-ax[0] * yz);
                    uz[2] = az[2] + (ax[0] * yy);
                    any = true, jac = true;
                }
            }
            double vel[3], jer[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                vel[d] = (((t[30 + d] * s1 + t[27 + d]) * s1 + t[24 + d]) * s1 + t[21 + d]) * s1 + t[18 + d];
                jer[d] = (t[51 + d] * s1 + t[48 + d]) * s1 + t[45 + d];
            }
            double g1[3] = {0, 0, 0}, g2[3] = {0, 0, 0}, g3[3] = {0, 0, 0};
            // ---- velocity (CPU.hpp:347-359) ----
            const double violaVel = vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2] - vmax_sqr;
            if (violaVel > 0.0) {
                const double v2 = violaVel * violaVel, v3 = v2 * violaVel;
                const double coef = w * chi1 * (3.0 * v2) * 2.0;
                g1[0] = coef * vel[0], g1[1] = coef * vel[1], g1[2] = coef * vel[2];
                cst += w * chi1 * v3;
                direct += omg * chi1 * v3 * invK;
                any = true;
            }
            // ---- thrust (CPU.hpp:361-385) ----
            const double sqrMagThr = fThr * fThr;
            const double violaThrl = thrmin_sqr - sqrMagThr;
            const double violaThrh = sqrMagThr - thrmax_sqr;
            if (violaThrl > 0.0 || violaThrh > 0.0) {
                const double v = violaThrl > 0.0 ? violaThrl : violaThrh;
                const double v2 = v * v, v3 = v2 * v;
                const double coef = (violaThrl > 0.0 ? -1.0 : 1.0) * (w * chi2 * (3.0 * v2) * 2.0);
                g2[0] = coef * hx, g2[1] = coef * hy, g2[2] = coef * hz;
                cst += w * chi2 * v3;
                direct += omg * chi2 * v3 * invK;
                any = true;
            }
            // ---- body rate (CPU.hpp:285-304, 387-398) ----
            {
                const double r0 = xx * jer[0] + xy * jer[1] + xz * jer[2];
                const double r1 = yy * jer[1] + yz * jer[2];
                const double r2 = zx * jer[0] + zy * jer[1] + zz * jer[2];
                const double bd0 = r0 * inv_f, bd1 = r1 * inv_f;
                const double violaBdr = bd1 * bd1 + bd0 * bd0 - bdrmax_sqr;
                if (violaBdr > 0.0) {
                    const double v2 = violaBdr * violaBdr, v3 = v2 * violaBdr;
                    const double coef = w * chi3 * (3.0 * v2);
                    const double t0 = 2.0 * bd0, t1 = 2.0 * bd1;
                    const double wx = t0 * xx, wy = t0 * xy + t1 * yy, wz = t0 * xz + t1 * yz;  // wv = t0 xB + t1 yB
                    // dSqrMagBdr = -hfac h - (1/f) [ dyB^T(r0 (z x wv) + r1 wv) + dzB^T(r0 (wv x y) + r2 wv) ]
                    const double hfac = (t0 * r0 + t1 * r1) * inv_f * inv_f * inv_f;
                    const double k0 = -coef * inv_f;
                    const double zwx = zy * wz - zz * wy, zwy = zz * wx - zx * wz, zwz = zx * wy - zy * wx;  // z x wv
                    const double wyx = wy * yz - wz * yy, wyy = -wx * yz, wyz = wx * yy;                      // wv x y
                    uy[0] += k0 * (r0 * zwx + r1 * wx), uy[1] += k0 * (r0 * zwy + r1 * wy), uy[2] += k0 * (r0 * zwz + r1 * wz);
                    uz[0] += k0 * (r0 * wyx + r2 * wx), uz[1] += k0 * (r0 * wyy + r2 * wy), uz[2] += k0 * (r0 * wyz + r2 * wz);
                    const double kh = -coef * hfac;
                    g2[0] += kh * hx, g2[1] += kh * hy, g2[2] += kh * hz;
                    const double kj = coef * inv_f;  // dJerSqrMagBdr = wv / f (CPU.hpp:297-299)
                    g3[0] = kj * wx, g3[1] = kj * wy, g3[2] = kj * wz;
                    cst += w * chi3 * v3;
                    direct += omg * chi3 * v3 * invK;
                    any = true, jac = true;
                }
            }
            if (jac) {
                // g2 += dzB^T uz + dyB^T uy
                const double zu = zx * uz[0] + zy * uz[1] + zz * uz[2];
                g2[0] += (uz[0] - zx * zu) * inv_f, g2[1] += (uz[1] - zy * zu) * inv_f, g2[2] += (uz[2] - zz * zu) * inv_f;
                const double yu = yy * uy[1] + yz * uy[2];
                const double t1y = uy[1] - yy * yu, t1z = uy[2] - yz * yu;  // (I - y y^T) uy; its x component is unused by C^T
                const double cy = -t1z, cz = t1y;                            // C^T t1 = (0, -t1z, t1y)
                const double zc = zy * cy + zz * cz;
                const double kk = inv_cn * inv_f;
                g2[0] += (-zx * zc) * kk, g2[1] += (cy - zy * zc) * kk, g2[2] += (cz - zz * zc) * kk;
            }
            if (any) {
                const double s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2, s5 = s4 * s1;
                // dJ/dc += beta0 (x) g0 + beta1 (x) g1 + beta2 (x) g2 + beta3 (x) g3
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    acc[(0 + d) * PEN_THREADS] += g0[d];
                    acc[(3 + d) * PEN_THREADS] += s1 * g0[d] + g1[d];
                    acc[(6 + d) * PEN_THREADS] += s2 * g0[d] + (2.0 * s1) * g1[d] + 2.0 * g2[d];
                    acc[(9 + d) * PEN_THREADS] += s3 * g0[d] + (3.0 * s2) * g1[d] + (6.0 * s1) * g2[d] + 6.0 * g3[d];
                    acc[(12 + d) * PEN_THREADS] += s4 * g0[d] + (4.0 * s3) * g1[d] + (12.0 * s2) * g2[d] + (24.0 * s1) * g3[d];
                    acc[(15 + d) * PEN_THREADS] += s5 * g0[d] + (5.0 * s4) * g1[d] + (20.0 * s3) * g2[d] + (60.0 * s2) * g3[d];
                }
                const double sn0 = t[57] * s1 + t[54], sn1 = t[58] * s1 + t[55], sn2 = t[59] * s1 + t[56];
                acc[18 * PEN_THREADS] += alpha * ((g0[0] * vel[0] + g0[1] * vel[1] + g0[2] * vel[2]) +
                                                  (g1[0] * ac[0] + g1[1] * ac[1] + g1[2] * ac[2]) +
                                                  (g2[0] * jer[0] + g2[1] * jer[1] + g2[2] * jer[2]) +
                                                  (g3[0] * sn0 + g3[1] * sn1 + g3[2] * sn2)) +
                                         direct;
                acc[19 * PEN_THREADS] += cst;
            }
            s1 += step;
        }
    }
    __syncthreads();
    // ---- fixed-order reduction of the per-thread partials into per-piece sums ----
    for (int task = tid; task < npieces * 20; task += PEN_THREADS) {
        const int p = task / 20, v = task - p * 20;
        const int gp = gp_base + p;
        const double* src = part + (size_t)v * PEN_THREADS + p * pl.tpp;
        double s = 0.0;
        for (int q = 0; q < pl.tpp; q++) s += src[q];
        const double* c = cf + (size_t)p * CF;
        const double T1 = Ts[p], T2 = T1 * T1, T3 = T2 * T1, T4 = T2 * T2, T5 = T4 * T1;
        if (v < 18) {
            // energy gradient w.r.t. c rows 3..5 (CPU.hpp:80-95), added first as in the reference
            const int row = v / 3, d = v - 3 * row;
            double e = 0.0;
            const double c3 = c[9 + d], c4 = c[12 + d], c5 = c[15 + d];
            if (row == 5) e = 240.0 * c3 * T3 + 720.0 * c4 * T4 + 1440.0 * c5 * T5;
            if (row == 4) e = 144.0 * c3 * T2 + 384.0 * c4 * T3 + 720.0 * c5 * T4;
            if (row == 3) e = 72.0 * c3 * T1 + 144.0 * c4 * T2 + 240.0 * c5 * T3;
            bv.gdC[18 * (size_t)gp + v] = e + s;
        } else {
            const double q33 = c[9] * c[9] + c[10] * c[10] + c[11] * c[11];
            const double q44 = c[12] * c[12] + c[13] * c[13] + c[14] * c[14];
            const double q55 = c[15] * c[15] + c[16] * c[16] + c[17] * c[17];
            const double q43 = c[12] * c[9] + c[13] * c[10] + c[14] * c[11];
            const double q53 = c[15] * c[9] + c[16] * c[10] + c[17] * c[11];
            const double q54 = c[15] * c[12] + c[16] * c[13] + c[17] * c[14];
            if (v == 18) {  // CPU.hpp:65-77
                const double e = 36.0 * q33 + 288.0 * q43 * T1 + 576.0 * q44 * T2 + 720.0 * q53 * T2 + 2880.0 * q54 * T3 +
                                 3600.0 * q55 * T4;
                bv.gdT[gp] = e + s;
            } else {  // CPU.hpp:507-520
                bv.energy[gp] = 36.0 * q33 * T1 + 144.0 * q43 * T2 + 192.0 * q44 * T3 + 240.0 * q53 * T3 + 720.0 * q54 * T4 +
                                720.0 * q55 * T5;
                bv.pen[gp] = s;
            }
        }
    }
}

#ifdef FRX_EXACT
// K2 (reference-order build) -----------------------------------------------------------------------
// One THREAD per trajectory, every loop sequential, every expression in the reference's order
// (MINCO_S3::addTimeIntPenalty CPU.hpp:188-408 with the Jacobians dxB/dyB/dzB materialised, getTrajJerkCost :507,
// addGradJbyT :65, addGradJbyC :80): the accumulators gdC / gdT / pena receive their terms in exactly the sequence
// the reference adds them, so the result is bit-identical to the CPU code.  Slow by construction; validation only.
__device__ __forceinline__ double ex_dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ double ex_sq3(const double* a) { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; }
__device__ __forceinline__ void ex_cross3(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ void ex_normalizeFDF(const double* x, double* xNor, double* G) {  // CPU.hpp:163-185, G row-major
    const double a = x[0], b = x[1], c = x[2];
    const double aSqr = a * a, bSqr = b * b, cSqr = c * c;
    const double ab = a * b, bc = b * c, ca = c * a;
    const double xSqrNorm = aSqr + bSqr + cSqr;
    const double xNorm = sqrt(xSqrNorm);
    const double den = xSqrNorm * xNorm;
    xNor[0] = a / xNorm, xNor[1] = b / xNorm, xNor[2] = c / xNorm;
    G[0] = bSqr + cSqr, G[1] = -ab, G[2] = -ca;
    G[3] = -ab, G[4] = aSqr + cSqr, G[5] = -bc;
    G[6] = -ca, G[7] = -bc, G[8] = aSqr + bSqr;
    for (int i = 0; i < 9; i++) G[i] /= den;
}

__global__ void __launch_bounds__(32) penalty_exact_kernel(BatchView bv) {
    int b = bv.sub_traj0 + blockIdx.x * 32 + threadIdx.x;
    if (bv.slot_list) {
        if (b >= bv.n_slots) return;
        b = bv.slot_list[b];
        if (b < 0) return;
    } else {
        if (b >= bv.sub_traj1) return;
        if (bv.active && !bv.active[b]) return;
    }
    const int p0 = bv.piece_off[b];
    const int N = bv.piece_off[b + 1] - p0;
    const DevParams P = bv.par[b];
    const double vMaxSqr = P.vmax_sqr, thrAccMinSqr = P.thrmin_sqr, thrAccMaxSqr = P.thrmax_sqr, bdrMaxSqr = P.bdrmax_sqr;
    const double* ell = P.ell;
    const double* ci = P.chi;
    double pena = 0.0;
    for (int i = 0; i < N; i++) {
        const int gp = p0 + i;
        const double* c = bv.c + 18 * (size_t)gp;  // 6x3 block, c[3*k+d]
        double* gc = bv.gdC + 18 * (size_t)gp;
        const double T1 = bv.T[gp], T2 = T1 * T1, T3 = T2 * T1, T4 = T2 * T2, T5 = T4 * T1;
        const double *c3 = c + 9, *c4 = c + 12, *c5 = c + 15;
        // energy, its time gradient and its coefficient gradient first (evalTrajCostGrad order, CPU.hpp:542-544)
        bv.energy[gp] = 36.0 * ex_sq3(c3) * T1 + 144.0 * ex_dot3(c4, c3) * T2 + 192.0 * ex_sq3(c4) * T3 + 240.0 * ex_dot3(c5, c3) * T3 +
                        720.0 * ex_dot3(c5, c4) * T4 + 720.0 * ex_sq3(c5) * T5;
        double gdT = 0.0;
        gdT += 36.0 * ex_sq3(c3) + 288.0 * ex_dot3(c4, c3) * T1 + 576.0 * ex_sq3(c4) * T2 + 720.0 * ex_dot3(c5, c3) * T2 +
               2880.0 * ex_dot3(c5, c4) * T3 + 3600.0 * ex_sq3(c5) * T4;
        for (int q = 0; q < 9; q++) gc[q] = 0.0;
        for (int cc = 0; cc < 3; cc++) {
            double g5 = 0.0, g4 = 0.0, g3 = 0.0;
            g5 += 240.0 * c3[cc] * T3 + 720.0 * c4[cc] * T4 + 1440.0 * c5[cc] * T5;
            g4 += 144.0 * c3[cc] * T2 + 384.0 * c4[cc] * T3 + 720.0 * c5[cc] * T4;
            g3 += 72.0 * c3[cc] * T1 + 144.0 * c4[cc] * T2 + 240.0 * c5[cc] * T3;
            gc[15 + cc] = g5, gc[12 + cc] = g4, gc[9 + cc] = g3;
        }
        const int cons = bv.cons[gp];
        const double step = T1 / cons;
        double s1 = 0.0;
        const int innerLoop = cons + 1;
        const double* faces = bv.faces + 6 * (size_t)bv.piece_face_off[gp];
        const int K = bv.piece_face_cnt[gp];
        for (int j = 0; j < innerLoop; j++) {
            const double s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2, s5 = s4 * s1;
            const double beta0[6] = {1.0, s1, s2, s3, s4, s5};
            const double beta1[6] = {0.0, 1.0, 2.0 * s1, 3.0 * s2, 4.0 * s3, 5.0 * s4};
            const double beta2[6] = {0.0, 0.0, 2.0, 6.0 * s1, 12.0 * s2, 20.0 * s3};
            const double beta3[6] = {0.0, 0.0, 0.0, 6.0, 24.0 * s1, 60.0 * s2};
            const double beta4[6] = {0.0, 0.0, 0.0, 0.0, 24.0, 120.0 * s1};
            const double alpha = 1.0 / cons * j;
            double pos[3], vel[3], acc[3], jer[3], sna[3];
            for (int d = 0; d < 3; d++) {
                double p = 0, v = 0, a = 0, jj = 0, sn = 0;
                for (int k = 0; k < 6; k++) {
                    p += c[3 * k + d] * beta0[k];
                    v += c[3 * k + d] * beta1[k];
                    a += c[3 * k + d] * beta2[k];
                    jj += c[3 * k + d] * beta3[k];
                    sn += c[3 * k + d] * beta4[k];
                }
                pos[d] = p, vel[d] = v, acc[d] = a, jer[d] = jj, sna[d] = sn;
            }
            const double h[3] = {acc[0], acc[1], acc[2] + P.gacc};
            double zB[3], dzB[9];
            ex_normalizeFDF(h, zB, dzB);
            const double czB[3] = {0.0, zB[2], -zB[1]};
            const double cdzB[9] = {0, 0, 0, dzB[6], dzB[7], dzB[8], -dzB[3], -dzB[4], -dzB[5]};
            double yB[3], dnczB[9];
            ex_normalizeFDF(czB, yB, dnczB);
            double xB[3];
            ex_cross3(yB, zB, xB);
            double dyB[9];
            for (int r = 0; r < 3; r++)
                for (int cc = 0; cc < 3; cc++)
                    dyB[3 * r + cc] = dnczB[3 * r + 0] * cdzB[0 + cc] + dnczB[3 * r + 1] * cdzB[3 + cc] + dnczB[3 * r + 2] * cdzB[6 + cc];
            double dxB[9];
            for (int m = 0; m < 3; m++) {
                const double dy[3] = {dyB[m], dyB[3 + m], dyB[6 + m]};
                const double dz[3] = {dzB[m], dzB[3 + m], dzB[6 + m]};
                double t0[3], t1[3];
                ex_cross3(dy, zB, t0);
                ex_cross3(yB, dz, t1);
                for (int r = 0; r < 3; r++) dxB[3 * r + m] = t0[r] + t1[r];
            }
            const double rotM[9] = {xB[0], yB[0], zB[0], xB[1], yB[1], zB[1], xB[2], yB[2], zB[2]};
            double gradSdTxyz[9];
            for (int r = 0; r < 3; r++) {
                gradSdTxyz[3 * r + 0] = dxB[3 * r] * jer[0] + dxB[3 * r + 1] * jer[1] + dxB[3 * r + 2] * jer[2];
                gradSdTxyz[3 * r + 1] = dyB[3 * r] * jer[0] + dyB[3 * r + 1] * jer[1] + dyB[3 * r + 2] * jer[2];
                gradSdTxyz[3 * r + 2] = dzB[3 * r] * jer[0] + dzB[3 * r + 1] * jer[1] + dzB[3 * r + 2] * jer[2];
            }
            const double fThr = sqrt(ex_sq3(h));
            const double dfThr[3] = {h[0] / fThr, h[1] / fThr, h[2] / fThr};
            const double sqrMagThr = fThr * fThr;
            const double dSqrMagThr[3] = {2 * h[0], 2 * h[1], 2 * h[2]};
            double rotTrDotJer[3];
            for (int r = 0; r < 3; r++) rotTrDotJer[r] = rotM[0 + r] * jer[0] + rotM[3 + r] * jer[1] + rotM[6 + r] * jer[2];
            const double bdr[3] = {rotTrDotJer[0] / fThr, rotTrDotJer[1] / fThr, rotTrDotJer[2] / fThr};
            const double xyBdr[3] = {-bdr[1], bdr[0], 0.0};
            const double sqrMagBdr = ex_sq3(xyBdr);
            double comb[9], dBdr[9];
            for (int q = 0; q < 9; q++) comb[q] = dxB[q] * rotTrDotJer[0] + dyB[q] * rotTrDotJer[1] + dzB[q] * rotTrDotJer[2];
            for (int r = 0; r < 3; r++)
                for (int cc = 0; cc < 3; cc++) {
                    const double first = (-rotTrDotJer[r]) * dfThr[cc] / (fThr * fThr);
                    const double prod = rotM[0 + r] * comb[0 + cc] + rotM[3 + r] * comb[3 + cc] + rotM[6 + r] * comb[6 + cc];
                    dBdr[3 * r + cc] = first - prod / fThr;
                }
            double dSqrMagBdr[3], dJerSqrMagBdr[3];
            for (int cc = 0; cc < 3; cc++)
                dSqrMagBdr[cc] = 2.0 * xyBdr[0] * (-dBdr[3 + cc]) + 2.0 * xyBdr[1] * dBdr[0 + cc] + 2.0 * xyBdr[2] * 0.0;
            for (int cc = 0; cc < 3; cc++) {
                const double r0 = rotM[3 * cc + 0] / fThr;
                const double r1 = rotM[3 * cc + 1] / fThr;
                dJerSqrMagBdr[cc] = 2.0 * xyBdr[0] * (-r1) + 2.0 * xyBdr[1] * r0 + 2.0 * xyBdr[2] * 0.0;
            }
            const double violaVel = ex_sq3(vel) - vMaxSqr;
            const double violaThrl = thrAccMinSqr - sqrMagThr;
            const double violaThrh = sqrMagThr - thrAccMaxSqr;
            const double violaBdr = sqrMagBdr - bdrMaxSqr;
            const double omg = (j == 0 || j == innerLoop - 1) ? 0.5 : 1.0;
            for (int k = 0; k < K; k++) {
                const double* on = faces + 6 * k;
                const double* pt = on + 3;
                double gradSdT[3];
                for (int m = 0; m < 3; m++) gradSdT[m] = gradSdTxyz[0 + m] * on[0] + gradSdTxyz[3 + m] * on[1] + gradSdTxyz[6 + m] * on[2];
                const double outerNormaldVel = ex_dot3(on, vel);
                double eNormGd[3];
                for (int m = 0; m < 3; m++) eNormGd[m] = (rotM[0 + m] * on[0] + rotM[3 + m] * on[1] + rotM[6 + m] * on[2]) * ell[m];
                const double eNorm = sqrt(ex_sq3(eNormGd));
                for (int m = 0; m < 3; m++) eNormGd[m] /= eNorm;
                const double dp[3] = {pos[0] - pt[0], pos[1] - pt[1], pos[2] - pt[2]};
                double signedDist = ex_dot3(on, dp) + eNorm;
                for (int m = 0; m < 3; m++) eNormGd[m] *= ell[m];
                signedDist += P.margin;
                if (signedDist > 0) {
                    const double signedDistSqr = signedDist * signedDist;
                    const double signedDistCub = signedDist * signedDistSqr;
                    const double fac = omg * step * ci[0] * 3.0 * signedDistSqr;
                    for (int r = 0; r < 6; r++)
                        for (int cc = 0; cc < 3; cc++) {
                            const double gx = (beta2[r] * on[0]) * dxB[0 + cc] + (beta2[r] * on[1]) * dxB[3 + cc] + (beta2[r] * on[2]) * dxB[6 + cc];
                            const double gy = (beta2[r] * on[0]) * dyB[0 + cc] + (beta2[r] * on[1]) * dyB[3 + cc] + (beta2[r] * on[2]) * dyB[6 + cc];
                            const double gz = (beta2[r] * on[0]) * dzB[0 + cc] + (beta2[r] * on[1]) * dzB[3 + cc] + (beta2[r] * on[2]) * dzB[6 + cc];
                            const double gradSdC = beta0[r] * on[cc] + gx * eNormGd[0] + gy * eNormGd[1] + gz * eNormGd[2];
                            gc[3 * r + cc] += fac * gradSdC;
                        }
                    const double gradSignedDt = alpha * (outerNormaldVel + gradSdT[0] * eNormGd[0] + gradSdT[1] * eNormGd[1] + gradSdT[2] * eNormGd[2]);
                    gdT += omg * ci[0] * (3.0 * signedDistSqr * gradSignedDt * step + signedDistCub / cons);
                    pena += omg * step * ci[0] * signedDistCub;
                }
            }
            if (violaVel > 0.0) {
                double pd = violaVel * violaVel;
                const double pn = pd * violaVel;
                pd *= 3.0;
                const double gradViolaVt = (2.0 * alpha * vel[0]) * acc[0] + (2.0 * alpha * vel[1]) * acc[1] + (2.0 * alpha * vel[2]) * acc[2];
                const double fac = omg * step * ci[1] * pd;
                for (int r = 0; r < 6; r++)
                    for (int cc = 0; cc < 3; cc++) gc[3 * r + cc] += fac * (2.0 * beta1[r] * vel[cc]);
                gdT += omg * (ci[1] * pd * gradViolaVt * step + ci[1] * pn / cons);
                pena += omg * step * ci[1] * pn;
            }
            if (violaThrl > 0.0) {
                double pd = violaThrl * violaThrl;
                const double pn = pd * violaThrl;
                pd *= 3.0;
                const double gt = (-alpha * dSqrMagThr[0]) * jer[0] + (-alpha * dSqrMagThr[1]) * jer[1] + (-alpha * dSqrMagThr[2]) * jer[2];
                const double fac = omg * step * ci[2] * pd;
                for (int r = 0; r < 6; r++)
                    for (int cc = 0; cc < 3; cc++) gc[3 * r + cc] += fac * (-beta2[r] * dSqrMagThr[cc]);
                gdT += omg * (ci[2] * pd * gt * step + ci[2] * pn / cons);
                pena += omg * step * ci[2] * pn;
            }
            if (violaThrh > 0.0) {
                double pd = violaThrh * violaThrh;
                const double pn = pd * violaThrh;
                pd *= 3.0;
                const double gt = (alpha * dSqrMagThr[0]) * jer[0] + (alpha * dSqrMagThr[1]) * jer[1] + (alpha * dSqrMagThr[2]) * jer[2];
                const double fac = omg * step * ci[2] * pd;
                for (int r = 0; r < 6; r++)
                    for (int cc = 0; cc < 3; cc++) gc[3 * r + cc] += fac * (beta2[r] * dSqrMagThr[cc]);
                gdT += omg * (ci[2] * pd * gt * step + ci[2] * pn / cons);
                pena += omg * step * ci[2] * pn;
            }
            if (violaBdr > 0.0) {
                double pd = violaBdr * violaBdr;
                const double pn = pd * violaBdr;
                pd *= 3.0;
                const double gt = alpha * (ex_dot3(dSqrMagBdr, jer) + ex_dot3(dJerSqrMagBdr, sna));
                const double fac = omg * step * ci[3] * pd;
                for (int r = 0; r < 6; r++)
                    for (int cc = 0; cc < 3; cc++) gc[3 * r + cc] += fac * (beta2[r] * dSqrMagBdr[cc] + beta3[r] * dJerSqrMagBdr[cc]);
                gdT += omg * (ci[3] * pd * gt * step + ci[3] * pn / cons);
                pena += omg * step * ci[3] * pn;
            }
            s1 += step;
        }
        bv.gdT[gp] = gdT;
        bv.pen[gp] = 0.0;
    }
    bv.pen[p0] = pena;  // the single running accumulator of the reference (CPU.hpp:203,406): K3 adds N-1 exact zeros to it
}
#endif  // FRX_EXACT

#ifdef FRX_EXACT  // reference-order build only: the banded system in the reference's own elimination order
// K3 ---------------------------------------------------------------------------------------------
// One trajectory per CTA of two warps.  The LU factors written by K1 (bv.lu, 16 doubles per row) are
// streamed through a ring of 6-row blocks in shared memory: warp 1 issues cp.async copies LOOKAHEAD
// blocks ahead of the sweep, warp 0 (lane d < 3 = right-hand side d) runs the two triangular sweeps of
// A^T lambda = dJ/dc (trajectory.hpp:724-751) row by row with the last six results in registers.
// Because a block is exactly one piece (6 rows), every shared-memory address of the sweeps is
// "block base + compile-time offset".  The sweeps are software-pipelined by hand: the part of row u+2
// that does not depend on row u+1 is issued before row u+1 is finished, so the dependent chain per row is
// one FMA + one MUL.  While the copies fly, warp 1 also sums the cost (sequential order of the reference),
// stages the interval table and pulls this trajectory's x / vertex ranges towards L2 for the epilogue.
// smem: ring[NBLK][6][16] | Bs[N][3] (adjoint RHS -> lambda) | gT[M] | gq[3*M] | gTc[coarseN] | em[2*M] | ivs[coarseN]
#define NBLK 8
#define LOOKAHEAD (NBLK - 2)  // blocks in flight ahead of the sweep
#define BLKD (6 * BW)         // doubles per block

// Issue the copy of block `blk` (rows 6*blk .. 6*blk+5) into its ring slot; blocks outside [0, M) are skipped.
__device__ __forceinline__ void issue_block(double* ring, const double* lug, int blk, int M, int lane) {
    if (blk >= 0 && blk < M) {
        double* dst = ring + (size_t)(blk % NBLK) * BLKD;
        const double* src = lug + (size_t)blk * BLKD;
        for (int e = lane; e < BLKD / 2; e += 32) cp_async16(dst + 2 * e, src + 2 * e);  // 48 x 16 B
    }
}

__global__ void __launch_bounds__(64)
    minco_backward_kernel(BatchView bv, const double* __restrict__ xin, double* __restrict__ cost_out,
                          double* __restrict__ grad_out) {
    extern __shared__ double sm[];
    const int b = bv.slot_list ? bv.slot_list[blockIdx.x] : bv.sub_traj0 + blockIdx.x;
    if (b < 0) return;  // empty solver slot
    if (!bv.slot_list && bv.active && !bv.active[b]) return;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int p0 = bv.piece_off[b];
    const int M = bv.piece_off[b + 1] - p0;
    const int N = 6 * M;
    const int co0 = bv.coarse_off[b];
    const int coarseN = bv.coarse_off[b + 1] - co0;
    const DevParams par = bv.par[b];
    const double* x = xin + bv.x_off[b];
    double* grad = grad_out + bv.x_off[b];
    const int dimT = bv.dimT[b];
    const int nx = bv.x_off[b + 1] - bv.x_off[b];
    double* ring = sm;
    double* Bs = ring + NBLK * BLKD;
    double* gT = Bs + (size_t)3 * N;
    double* gq = gT + M;
    double* gTc = gq + 3 * (size_t)M;
    double* em = gTc + coarseN;
    double* tfac = em + 2 * (size_t)M;  // d(T)/d(tau) of the time map, per polytope
    int* ivs = reinterpret_cast<int*>(tfac + coarseN);
    int* ioff = ivs + coarseN;  // first piece of each polytope (prefix sum of intervals)
    int* wpi = ioff + coarseN;  // per waypoint: free dimension k, xi offset, vertex offset (3 x M), for addLayerPGrad
    PHASE_MARK(2, 0);

    const double* lug = bv.lu + (size_t)BW * 6 * p0;
    const double* cg_ = bv.c + 18 * (size_t)p0;
    // epilogue operands of this thread's first piece, requested now so that their latency is hidden
    double pc[18], pT = 0.0, pgdT = 0.0;
    if (tid < M) {
#pragma unroll
        for (int q = 0; q < 18; q++) pc[q] = cg_[18 * tid + q];
        pT = bv.T[p0 + tid];
        pgdT = bv.gdT[p0 + tid];
    }
    // ---- prologue ----
    if (warp == 1) {
        for (int e = lane; e < BLKD; e += 32) ring[(size_t)(NBLK - 1) * BLKD + e] = 0.0;  // block -1 (rows -6..-1) = zeros
        for (int blk = 0; blk < LOOKAHEAD; blk++) issue_block(ring, lug, blk, M, lane);
        for (int i = lane; i < M - 1; i += 32) {
            wpi[i] = bv.wp_vert_cnt[p0 + i] - 1;
            wpi[M + i] = bv.wp_xi_off[p0 + i];
            wpi[2 * M + i] = bv.wp_vert_off[p0 + i];
        }
        cp_async_commit();
        // while the copies fly: cost = sum energy + sum penalty + rho * sum T (CPU.hpp:542,406,988).  The operands are
        // gathered by all lanes; the three sums are sequential (reference order), one lane each.
        for (int i = lane; i < M; i += 32) {
            em[i] = bv.energy[p0 + i];
            em[M + i] = bv.pen[p0 + i];
        }
        for (int i = lane; i < coarseN; i += 32) ivs[i] = bv.intervals[co0 + i];
        if (par.softT) {
            for (int i = lane; i < coarseN; i += 32) {  // time map and its derivative (CPU.hpp:639-641, 830-838)
                const double t = x[i];
                double Tv, dv;
                if (par.c2) {
                    Tv = c2_forward(t);
                    if (t > 0)
                        dv = t + 1.0;
                    else {
                        const double den = (0.5 * t - 1.0) * t + 1.0;
                        dv = (1.0 - t) / (den * den);
                    }
                } else {
                    Tv = exp(t);
                    dv = Tv;
                }
                gTc[i] = Tv;
                tfac[i] = dv;
            }
        } else {
            for (int i = lane; i < M; i += 32) gT[i] = bv.T[p0 + i];  // fixed total time: rho = 0 anyway
        }
        __syncwarp();
        double part = 0.0;
        if (lane == 0) {
            for (int i = 0; i < M; i++) part += em[i];
        } else if (lane == 1) {
            for (int i = 0; i < M; i++) part += em[M + i];
        } else if (lane == 2) {
            if (par.softT) {
                for (int i = 0; i < coarseN; i++) part += gTc[i];
            } else {
                for (int i = 0; i < M; i++) part += gT[i];
            }
        } else if (lane == 3) {
            int o = 0;
            for (int i = 0; i < coarseN; i++) {
                ioff[i] = o;
                o += ivs[i];
            }
        }
        const double pn = __shfl_sync(0xffffffffu, part, 1), sT = __shfl_sync(0xffffffffu, part, 2);
        if (lane == 0) cost_out[b] = (part + pn) + par.rho * sT;
        // pull x and the vertex range of this trajectory towards L2 for addLayerPGrad
        for (int o = lane * 16; o < nx; o += 32 * 16) prefetch_l2(x + o);
        if (M > 1) {
            const double* v0 = bv.verts + 3 * (size_t)bv.wp_vert_off[p0];
            const double* v1 = bv.verts + 3 * ((size_t)bv.wp_vert_off[p0 + M - 2] + bv.wp_vert_cnt[p0 + M - 2]);
            for (const double* q = v0 + lane * 16; q < v1; q += 32 * 16) prefetch_l2(q);
        }
        cp_async_wait<0>();
    } else {
        const double* gdC = bv.gdC + 18 * (size_t)p0;
        for (int e = lane; e < 3 * N / 2; e += 32) cp_async16(Bs + 2 * e, gdC + 2 * e);  // 3N doubles, 16-byte aligned both sides
        cp_async_commit();
        cp_async_wait<0>();
    }
    PHASE_MARK(2, 1);
    __syncthreads();
    PHASE_MARK(2, 2);

    // ---- sweep 1: U^T z = b, ascending rows.  Phase p = rows 6p..6p+5 reads blocks p-1 and p. ----
    // coefficient u_{i-s,i} = row (i-s), slot 6+s; update order = ascending j = i-s, i.e. s = 6..1.
#define UCOEF(u, s) (((u) - (s) >= 0) ? cur[((u) - (s)) * BW + 6 + (s)] : prv[((u) - (s) + 6) * BW + 6 + (s)])
#define PRE1(u) (((((bs[3 * (u)] - UCOEF(u, 6) * w[((u) + 0) % 6]) - UCOEF(u, 5) * w[((u) + 1) % 6]) - UCOEF(u, 4) * w[((u) + 2) % 6]) - \
                  UCOEF(u, 3) * w[((u) + 3) % 6]) - UCOEF(u, 2) * w[((u) + 4) % 6])
#define FIN1(u, pre) (((pre)-UCOEF(u, 1) * w[((u) + 5) % 6]) / cur[(u) * BW + 6])
    {
        double w[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // z_{i-6..i-1} of this lane's right-hand side: z_j at w[j mod 6]
        for (int p = 0; p < M; p++) {
            if (warp == 1) {
                issue_block(ring, lug, p + LOOKAHEAD, M, lane);  // into the slot of block p-2 (dead)
                cp_async_commit();
                cp_async_wait<LOOKAHEAD - 1>();  // block p+1 has landed
            } else if (lane < 3) {
                const double* cur = ring + (size_t)(p % NBLK) * BLKD;
                const double* prv = ring + (size_t)((p + NBLK - 1) % NBLK) * BLKD;
                double* bs = Bs + 18 * (size_t)p + lane;
                // pipelined: PRE(u+2) (needs z_u) is issued before FIN(u+1)
                double pre0 = PRE1(0);
                double pre1 = PRE1(1);
                const double z0 = FIN1(0, pre0);
                w[0] = z0;
                double pre2 = PRE1(2);
                const double z1 = FIN1(1, pre1);
                w[1] = z1;
                double pre3 = PRE1(3);
                const double z2 = FIN1(2, pre2);
                w[2] = z2;
                double pre4 = PRE1(4);
                const double z3 = FIN1(3, pre3);
                w[3] = z3;
                double pre5 = PRE1(5);
                const double z4 = FIN1(4, pre4);
                w[4] = z4;
                const double z5 = FIN1(5, pre5);
                w[5] = z5;
                bs[0] = z0, bs[3] = z1, bs[6] = z2, bs[9] = z3, bs[12] = z4, bs[15] = z5;
            }
            __syncthreads();
        }
    }
#undef UCOEF
#undef PRE1
#undef FIN1
    PHASE_MARK(2, 3);
    // ---- transition: the ring now holds blocks M-NBLK..M-1; block M (rows >= N) must read as zeros ----
    if (warp == 1) {
        cp_async_wait<0>();
        for (int e = lane; e < BLKD; e += 32) ring[(size_t)(M % NBLK) * BLKD + e] = 0.0;
    }
    __syncthreads();
    // ---- sweep 2: L^T lambda = z, descending rows.  Phase p = rows 6p+5..6p reads blocks p and p+1. ----
    // coefficient l_{i+t,i} = row (i+t), slot 6-t; update order = descending j = i+t, i.e. t = 6..1.
#define LCOEF(u, t) (((u) + (t) < 6) ? cur[((u) + (t)) * BW + 6 - (t)] : nxt[((u) + (t) - 6) * BW + 6 - (t)])
#define PRE2(u) (((((bs[3 * (u)] - LCOEF(u, 6) * w[((u) + 6) % 6]) - LCOEF(u, 5) * w[((u) + 5) % 6]) - LCOEF(u, 4) * w[((u) + 4) % 6]) - \
                  LCOEF(u, 3) * w[((u) + 3) % 6]) - LCOEF(u, 2) * w[((u) + 2) % 6])
#define FIN2(u, pre) ((pre)-LCOEF(u, 1) * w[((u) + 1) % 6])
    {
        double w[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // lambda_{i+1..i+6}: lambda_j at w[j mod 6]
        for (int p = M - 1; p >= 0; p--) {
            if (warp == 1) {
                issue_block(ring, lug, p - LOOKAHEAD, M, lane);  // into the slot of block p+2 (dead: the sweep reads p and p+1)
                cp_async_commit();
                cp_async_wait<LOOKAHEAD - 1>();  // block p-1 has landed
            } else if (lane < 3) {
                const double* cur = ring + (size_t)(p % NBLK) * BLKD;
                const double* nxt = ring + (size_t)((p + 1) % NBLK) * BLKD;
                double* bs = Bs + 18 * (size_t)p + lane;
                double pre5 = PRE2(5);
                double pre4 = PRE2(4);
                const double l5 = FIN2(5, pre5);
                w[5] = l5;
                double pre3 = PRE2(3);
                const double l4 = FIN2(4, pre4);
                w[4] = l4;
                double pre2 = PRE2(2);
                const double l3 = FIN2(3, pre3);
                w[3] = l3;
                double pre1 = PRE2(1);
                const double l2 = FIN2(2, pre2);
                w[2] = l2;
                double pre0 = PRE2(0);
                const double l1 = FIN2(1, pre1);
                w[1] = l1;
                const double l0 = FIN2(0, pre0);
                w[0] = l0;
                bs[0] = l0, bs[3] = l1, bs[6] = l2, bs[9] = l3, bs[12] = l4, bs[15] = l5;
            }
            __syncthreads();
        }
    }
#undef LCOEF
#undef PRE2
#undef FIN2
    PHASE_MARK(2, 4);
    // ---- addPropCtoT (CPU.hpp:104-151), addPropCtoP (:154-161), rho (:989) ----
    for (int i = tid; i < M; i += 64) {
        if (i != tid) {  // pieces beyond the first 64: not prefetched
#pragma unroll
            for (int q = 0; q < 18; q++) pc[q] = cg_[18 * i + q];
            pT = bv.T[p0 + i];
            pgdT = bv.gdT[p0 + i];
        }
        const double T1 = pT, T2 = T1 * T1, T3 = T2 * T1, T4 = T2 * T2;
        double s = 0.0;
        const int rows = (i < M - 1) ? 6 : 3;
        const int base = (i < M - 1) ? (6 * i + 3) : (6 * M - 3);
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double c1 = pc[3 + d], c2 = pc[6 + d], c3 = pc[9 + d], c4 = pc[12 + d], c5 = pc[15 + d];
            const double nVel = -(c1 + 2.0 * T1 * c2 + 3.0 * T2 * c3 + 4.0 * T3 * c4 + 5.0 * T4 * c5);
            const double nAcc = -(2.0 * c2 + 6.0 * T1 * c3 + 12.0 * T2 * c4 + 20.0 * T3 * c5);
            const double nJer = -(6.0 * c3 + 24.0 * T1 * c4 + 60.0 * T2 * c5);
            if (rows == 6) {
                const double nSnp = -(24.0 * c4 + 120.0 * T1 * c5);
                const double nCrk = -120.0 * c5;
                const double B1[6] = {nSnp, nCrk, nVel, nVel, nAcc, nJer};
#pragma unroll
                for (int r = 0; r < 6; r++) s += B1[r] * Bs[3 * (base + r) + d];
            } else {
                const double B2[3] = {nVel, nAcc, nJer};
#pragma unroll
                for (int r = 0; r < 3; r++) s += B2[r] * Bs[3 * (base + r) + d];
            }
        }
        gT[i] = pgdT + s + par.rho;
        if (i < M - 1)
            for (int d = 0; d < 3; d++) gq[3 * i + d] = Bs[3 * (6 * i + 5) + d];
    }
    __syncthreads();
    PHASE_MARK(2, 5);
    // ---- mergeToCoarseGradT (CPU.hpp:946-959): mean over the pieces of each polytope; one thread per polytope ----
    for (int i = tid; i < coarseN; i += 64) {
        const int iv = ivs[i], offset = ioff[i];
        double sgt = 0.0;
        for (int a = 0; a < iv; a++) sgt += gT[offset + a];
        const double gm = sgt / iv;
        // ---- addLayerTGrad (CPU.hpp:816-894), free total time: the factor was computed in the prologue ----
        if (par.softT) {
            if (i < dimT) grad[i] = gm * tfac[i];
        } else
            gTc[i] = gm;
    }
    if (!par.softT) {
        __syncthreads();
        if (tid == 0) {
            const int Ms1 = dimT;
            const double gTail = par.sumT * gTc[Ms1];
            double expTauSum = 0.0, gFreeDotExpTau = 0.0;
            for (int i = 0; i < Ms1; i++) {
                const double t = x[i];
                const double gFree = par.sumT * gTc[i];
                double eTau;
                if (par.c2) {
                    if (t > 0)
                        eTau = (0.5 * t + 1.0) * t + 1.0;
                    else
                        eTau = 1.0 / ((0.5 * t - 1.0) * t + 1.0);
                } else
                    eTau = exp(t);
                expTauSum += eTau;
                gFreeDotExpTau += eTau * gFree;
            }
            const double den = expTauSum + 1.0;
            for (int i = 0; i < Ms1; i++) {
                const double t = x[i];
                const double gFree = par.sumT * gTc[i];
                double dE;
                if (par.c2) {
                    if (t > 0)
                        dE = t + 1.0;
                    else {
                        const double ds = (0.5 * t - 1.0) * t + 1.0;
                        dE = (1.0 - t) / (ds * ds);
                    }
                } else
                    dE = exp(t);
                grad[i] = (gFree - gTail) * dE / den - (gFreeDotExpTau - gTail * expTauSum) * dE / (den * den);
            }
        }
    }
    PHASE_MARK(2, 6);
    // ---- addLayerPGrad (CPU.hpp:897-928) ----
    for (int i = tid; i < M - 1; i += 64) {
        const int k = bv.wp_vert_cnt[p0 + i] - 1;
        const double* V = bv.verts + 3 * (size_t)bv.wp_vert_off[p0 + i];
        const double* xi = x + dimT + bv.wp_xi_off[p0 + i];
        double* gxi = grad + dimT + bv.wp_xi_off[p0 + i];
        double qn = 0.0;
#pragma unroll 8
        for (int a = 0; a < k; a++) qn += xi[a] * xi[a];
        const double q1 = qn + 1.0, q1s = q1 * q1;
        const double gx = gq[3 * i], gy = gq[3 * i + 1], gz = gq[3 * i + 2];
        double gdot = 0.0;
#pragma unroll 4
        for (int a = 0; a < k; a++) {
            const double r = 2.0 / q1 * xi[a];
            const double gd = (V[3 * (a + 1)] * gx + V[3 * (a + 1) + 1] * gy + V[3 * (a + 1) + 2] * gz) * r * 2.0;
            gdot += gd * xi[a];
        }
#pragma unroll 4
        for (int a = 0; a < k; a++) {
            const double r = 2.0 / q1 * xi[a];
            const double gd = (V[3 * (a + 1)] * gx + V[3 * (a + 1) + 1] * gy + V[3 * (a + 1) + 2] * gz) * r * 2.0;
            gxi[a] = gd * 2.0 / q1 - xi[a] * 4.0 * gdot / q1s;
        }
    }
    PHASE_MARK(2, 7);
}
#endif  // FRX_EXACT (banded kernels)

#ifndef FRX_EXACT
// =================================================================================================
// Fast build: the MINCO map (q,T) -> c and its adjoint through the REDUCED system.
//
// MINCO_S3::generate (CPU.hpp:425-505) solves a 6M x 6M banded system whose rows are: head / tail states, the waypoints,
// C^2 continuity of position at the junctions, and continuity of jerk and snap.  A quintic piece is fixed by the states
// (p, v, a) at its two ends, so with the node states as unknowns every row except jerk / snap continuity holds by
// construction, the node positions are the waypoints, and what remains are the node velocities and accelerations
// y_j = (v_j, a_j), j = 1..M-1.  Jerk / snap continuity at node j are the stationarity conditions of the jerk energy
// w.r.t. (a_j, v_j) (d/da_end of 1/2 int jerk^2 = jerk(T), d/dv_end = -snap(T), d/dp_end = crackle(T)), hence
//
//        S(T) y = r(q, T),        S symmetric positive definite, block tridiagonal, 2 x 2 blocks,
//        D_j = [192 (1/Ta^3 + 1/Tb^3),  36 (1/Tb^2 - 1/Ta^2);  .,  9 (1/Ta + 1/Tb)],   Ta = T_{j-1}, Tb = T_j,
//        L_j = [168/Ta^3, 24/Ta^2;  -24/Ta^2, -3/Ta]  (couples node j to node j-1),   U_j = L_{j+1}^T,
//
// the same c as the reference's banded LU up to rounding (the 6M sequential pivots of trajectory.hpp:655-719 become M-1
// sequential 2 x 2 block pivots; no pivoting is needed because S is SPD).  The back-propagation of CPU.hpp:97-161
// (A^T lambda = dJ/dc, addPropCtoT, addPropCtoP) becomes: dJ/d(node states) = H(T_i)^T dJ/dc_i per piece (H = the
// quintic Hermite map), one more solve S mu = dJ/dy with the SAME factors, and closed forms for the rest: with the
// "adjoint quintic" ct_i = H(T_i)(0, mu_i, 0, mu_{i+1}),
//        dJ/dq_j  = Pbar_j + crackle(ct_j) - crackle(ct_{j-1}),
//        dJ/dT_i += -(P1bar.v_{i+1} + V1bar.a_{i+1} + A1bar.jerk_i(T_i))                      [dc/dT at fixed end states]
//                   + jerk_c jerk_ct - snap_ct.a_{i+1} - snap_c.muA_{i+1} + crackle_ct.v_{i+1} + crackle_c.muV_{i+1}  at t = T_i.
// Derivation and a numpy check against a dense restatement of the reference system: scripts/proto/reduced_minco.py.
// =================================================================================================
// Parallel cyclic reduction of the reduced MINCO system  A_j y_{j-1} + D_j y_j + A_{j+1}^T y_{j+1} = r_j  (nodes j = 1..N, 2 x 2
// blocks, symmetric positive definite; three right-hand sides = the axes), one thread per node (N <= PCR_MAX_NODES = the CTA):
// at stride s every node eliminates its neighbours j -+ s at once,
//   alpha = -A_j Dinv_{j-s},  gamma = -A_{j+s}^T Dinv_{j+s},
//   D_j += alpha A_j^T + gamma A_{j+s},   A_j := alpha A_{j-s},   r_j += alpha r_{j-s} + gamma r_{j+s}
// (the reduced matrices stay symmetric, so the upper coupling is never stored: C_j = A_{j+s}^T), and after ceil(log2 N) strides
// y_j = Dinv_j r_j.  Six dependent rounds of ~70 fp64 instructions per thread for 64 pieces, against the 63-node chain of the
// block LDL^T (three active lanes, ~320 cycles per node) — scripts/proto/pcr_reduced.py is the numpy emulation of this slot
// algebra (1e-13 of a dense solve with durations spread over two decades).  Blocks: D (d11 d12 d22), A (row-major; A_1 is taken
// as zero: the head state is folded into r_1), R[6j + d] / R[6j + 3 + d] (velocity / acceleration rows, axis d).
#define PCR_MAX_NODES 64
#define PCR_STATE 13  // doubles per node that the neighbours read: inverse block (3), lower coupling (4), right-hand sides (6)
// In: D (stride DS), A (stride AS), R (6 per node) indexed by node j = 1..N.  W: scratch of PCR_STATE * N doubles, may overlap
// D / A / R (they are read into registers first): the state the neighbours read lives there component-major
// (W[c * N + j - 1]: consecutive lanes hit consecutive banks; node-major records of 4 or 6 doubles are 8- and
// 4-way bank conflicts for 8-byte loads).  Out: Yout[6j + d] / Yout[6j + 3 + d], may overlap W.  Every thread of the
// 64-thread CTA calls it.
template <int DS, int AS>  // strides of D and A in doubles per node (K1: separate arrays 3 / 4; K3: both inside the 8-double node record)
__device__ __forceinline__ void pcr_solve(const double* D, const double* A, const double* R, double* W, double* Yout, int N, int tid) {
    const int j = tid + 1;
    const bool on = j <= N;
    double d11 = 1.0, d12 = 0.0, d22 = 1.0, a11 = 0.0, a12 = 0.0, a21 = 0.0, a22 = 0.0;
    double r[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double i11 = 0.0, i12 = 0.0, i22 = 0.0;
    if (on) {
        d11 = D[DS * j], d12 = D[DS * j + 1], d22 = D[DS * j + 2];
        if (j > 1) a11 = A[AS * j], a12 = A[AS * j + 1], a21 = A[AS * j + 2], a22 = A[AS * j + 3];  // A_1 := 0
#pragma unroll
        for (int q = 0; q < 6; q++) r[q] = R[6 * j + q];
        const double id = fast_rcp3(fma(d11, d22, -(d12 * d12)));
        i11 = d22 * id, i12 = -(d12 * id), i22 = d11 * id;
    }
    __syncthreads();  // the inputs are in registers: W may overwrite them
    const int NS = N;
    double* w = W + tid;  // component c of this node: w[c * NS]
    auto publish = [&]() {
        w[0] = i11, w[NS] = i12, w[2 * NS] = i22;
        w[3 * NS] = a11, w[4 * NS] = a12, w[5 * NS] = a21, w[6 * NS] = a22;
#pragma unroll
        for (int q = 0; q < 6; q++) w[(7 + q) * NS] = r[q];
    };
    if (on) publish();
    __syncthreads();
    for (int s = 1; s < N; s <<= 1) {
        double n11 = d11, n12 = d12, n22 = d22, b11 = 0.0, b12 = 0.0, b21 = 0.0, b22 = 0.0;
        if (on) {
            if (j - s >= 1) {
                const double* lo = w - s;
                const double l11 = lo[0], l12 = lo[NS], l22 = lo[2 * NS];
                const double al11 = -fma(a11, l11, a12 * l12), al12 = -fma(a11, l12, a12 * l22);
                const double al21 = -fma(a21, l11, a22 * l12), al22 = -fma(a21, l12, a22 * l22);
                n11 = fma(al11, a11, fma(al12, a12, n11));
                n12 = fma(al11, a21, fma(al12, a22, n12));
                n22 = fma(al21, a21, fma(al22, a22, n22));
                const double c11 = lo[3 * NS], c12 = lo[4 * NS], c21 = lo[5 * NS], c22 = lo[6 * NS];
                b11 = fma(al11, c11, al12 * c21), b12 = fma(al11, c12, al12 * c22);
                b21 = fma(al21, c11, al22 * c21), b22 = fma(al21, c12, al22 * c22);
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double rv = lo[(7 + d) * NS], ra = lo[(10 + d) * NS];
                    r[d] = fma(al11, rv, fma(al12, ra, r[d]));
                    r[3 + d] = fma(al21, rv, fma(al22, ra, r[3 + d]));
                }
            }
            if (j + s <= N) {
                const double* hi = w + s;
                const double h11 = hi[3 * NS], h12 = hi[4 * NS], h21 = hi[5 * NS], h22 = hi[6 * NS];
                const double u11 = hi[0], u12 = hi[NS], u22 = hi[2 * NS];
                const double g11 = -fma(h11, u11, h21 * u12), g12 = -fma(h11, u12, h21 * u22);  // gamma = -A_hi^T Dinv_hi
                const double g21 = -fma(h12, u11, h22 * u12), g22 = -fma(h12, u12, h22 * u22);
                n11 = fma(g11, h11, fma(g12, h21, n11));
                n12 = fma(g11, h12, fma(g12, h22, n12));
                n22 = fma(g21, h12, fma(g22, h22, n22));
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double rv = hi[(7 + d) * NS], ra = hi[(10 + d) * NS];
                    r[d] = fma(g11, rv, fma(g12, ra, r[d]));
                    r[3 + d] = fma(g21, rv, fma(g22, ra, r[3 + d]));
                }
            }
        }
        __syncthreads();  // every node has read its neighbours
        if (on) {
            d11 = n11, d12 = n12, d22 = n22, a11 = b11, a12 = b12, a21 = b21, a22 = b22;
            const double id = fast_rcp3(fma(d11, d22, -(d12 * d12)));
            i11 = d22 * id, i12 = -(d12 * id), i22 = d11 * id;
            if ((s << 1) < N) publish();  // nobody reads the state after the last round
        }
        __syncthreads();
    }
    if (on) {
#pragma unroll
        for (int d = 0; d < 3; d++) {
            Yout[6 * j + d] = fma(i11, r[d], i12 * r[3 + d]);
            Yout[6 * j + 3 + d] = fma(i12, r[d], i22 * r[3 + d]);
        }
    }
}

// shared memory of minco_backward_kernel in doubles (layout in its header comment); the last slot is the mbarrier
__host__ __device__ inline size_t backward_smem_doubles(int M, int coarseN) {
    return (size_t)53 * M + 2 * (size_t)coarseN + 1 + (2 * (size_t)coarseN + 3 * (size_t)M + 1) / 2 + 3;
}
#define NODE_F 8  // doubles per node kept in bv.lu for K3.  Up to PCR_MAX_NODES + 1 pieces: [0..2] D_j (d11, d12, d22), [3..6] L_j —
                  // the node blocks themselves (K3 runs its own cyclic reduction); longer: [0..2] inverse pivot block, [3..6] G_j = L_j Di_{j-1}

// K1 (fast) -----------------------------------------------------------------------------------------
// One trajectory per CTA of two warps.  forwardT / splitToFineT / forwardP (CPU.hpp:626, :930, :729) by all threads, the
// node blocks by one thread per node, the block LDL^T + forward substitution + back substitution by lanes 0..2 of
// warp 0 (one right-hand side = one axis each; the 2 x 2 factorisation is recomputed by each of the three lanes instead
// of being broadcast), the Hermite expansion to c by one thread per (piece, axis).
// smem (doubles): Tc[coarseN+1] | Ts[M] | iT[M] | P[3(M+1)] | Y[6(M+1)] | nD[3M] | nL[4M] | nR[6M] | nG[4M] | wpi[3M ints]
__global__ void __launch_bounds__(64) minco_forward_kernel(BatchView bv, const double* __restrict__ xin) {
    extern __shared__ double sm[];
    const int b = bv.slot_list ? bv.slot_list[blockIdx.x] : bv.sub_traj0 + blockIdx.x;
    if (b < 0) return;  // empty solver slot
    if (!bv.slot_list && bv.active && !bv.active[b]) return;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int p0 = bv.piece_off[b];
    const int M = bv.piece_off[b + 1] - p0;
    const int co0 = bv.coarse_off[b];
    const int coarseN = bv.coarse_off[b + 1] - co0;
    const DevParams par = bv.par[b];
    const double* x = xin + bv.x_off[b];
    const int dimT = bv.dimT[b];
    double* Tc = sm;
    double* Ts = Tc + coarseN + 1;
    double* iT = Ts + M;
    double* P = iT + M;
    double* Y = P + 3 * (M + 1);
    double* nD = Y + 6 * (M + 1);
    double* nL = nD + 3 * M;
    double* nR = nL + 4 * M;
    double* nG = nR + 6 * M;
    int* wpi = reinterpret_cast<int*>(nG + 4 * M);
    CTA_SPAN(0, 0);
    PHASE_MARK(0, 0);

    // this trajectory's x and vertex ranges -> L2, one bulk prefetch each, issued before anything depends on them
    if (tid == 32) bulk_prefetch_l2(x, sizeof(double) * (size_t)(bv.x_off[b + 1] - bv.x_off[b]));
    if (tid == 33 && M > 1) {
        const double* v0 = bv.verts + 3 * (size_t)bv.wp_vert_off[p0];
        const double* v1 = bv.verts + 3 * ((size_t)bv.wp_vert_off[p0 + M - 2] + bv.wp_vert_cnt[p0 + M - 2]);
        bulk_prefetch_l2(v0, (size_t)((const char*)v1 - (const char*)v0));
    }
    // waypoint index tables (free dimension, xi offset, vertex offset): one round trip instead of one per use
    for (int i = tid; i < M - 1; i += 64) {
        wpi[i] = bv.wp_vert_cnt[p0 + i] - 1;
        wpi[M + i] = bv.wp_xi_off[p0 + i];
        wpi[2 * M + i] = bv.wp_vert_off[p0 + i];
    }
    // head / tail states (column-major 3 x 3 [p v a] each): node 0 and node M
    if (tid < 18) {
        const double v = bv.headtail[18 * (size_t)b + tid];
        const int e = tid < 9 ? tid : tid - 9, node = tid < 9 ? 0 : M;
        if (e < 3)
            P[3 * node + e] = v;
        else
            Y[6 * node + (e - 3)] = v;
    }
    // ---- forwardT (CPU.hpp:626-676) ----
    if (par.softT) {
        for (int i = tid; i < coarseN; i += 64) Tc[i] = par.c2 ? c2_forward(x[i]) : exp(x[i]);
    } else if (tid == 0) {  // fixed total time: sequential sums as in the reference (rare path)
        const int Ms1 = dimT;
        for (int i = 0; i < Ms1; i++) Tc[i] = par.c2 ? c2_forward(x[i]) : exp(x[i]);
        Tc[Ms1] = 0.0;
        double s = 0.0;
        for (int i = 0; i <= Ms1; i++) s += Tc[i];
        const double den = 1.0 + s;
        for (int i = 0; i <= Ms1; i++) Tc[i] /= den;
        s = 0.0;
        for (int i = 0; i <= Ms1; i++) s += Tc[i];
        Tc[Ms1] = 1.0 - s;
        for (int i = 0; i <= Ms1; i++) Tc[i] *= par.sumT;
    }
    __syncthreads();
    // ---- splitToFineT (CPU.hpp:930-944) + forwardP (CPU.hpp:729-747) ----
    for (int i = tid; i < M; i += 64) {
        const int pl = bv.piece_poly[p0 + i];
        const double Tv = Tc[pl] / bv.intervals[co0 + pl];
        bv.T[p0 + i] = Tv;
        Ts[i] = Tv;
        iT[i] = 1.0 / Tv;
    }
    {   // four lanes per waypoint (sector-sized coalesced reads of xi and V, many loads in flight), one pass:
        //   q = v0 + sc^2 * sum_a V_{a+1} xi_a^2,  sc = 2 / (1 + |xi|^2)
        const int grp = tid >> 2, l4 = tid & 3;
        const unsigned gmask = 0xFu << (lane & ~3);
        for (int i = grp; i < M - 1; i += 16) {
            const int k = wpi[i];
            const double* V = bv.verts + 3 * (size_t)wpi[2 * M + i];
            const double* xi = x + dimT + wpi[M + i];
            double sn = 0.0, sx = 0.0, sy = 0.0, sz = 0.0;
#pragma unroll 4
            for (int a = l4; a < k; a += 4) {
                const double xa = xi[a], x2 = xa * xa;
                sn += x2;
                sx += V[3 * (a + 1) + 0] * x2;
                sy += V[3 * (a + 1) + 1] * x2;
                sz += V[3 * (a + 1) + 2] * x2;
            }
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                sn += __shfl_xor_sync(gmask, sn, o);
                sx += __shfl_xor_sync(gmask, sx, o);
                sy += __shfl_xor_sync(gmask, sy, o);
                sz += __shfl_xor_sync(gmask, sz, o);
            }
            const double sc = 2.0 / (1.0 + sn), sc2 = sc * sc;
            const double sd_ = l4 == 0 ? sx : (l4 == 1 ? sy : sz);
            bv.wpw[4 * (size_t)(p0 + i) + l4] = (l4 == 3) ? sn : sd_;  // reused by addLayerPGrad in K3
            if (l4 < 3) {
                const double qd = V[l4] + sc2 * sd_;
                bv.q[3 * (size_t)(p0 + i) + l4] = qd;
                P[3 * (i + 1) + l4] = qd;
            }
        }
    }
    __syncthreads();
    PHASE_MARK(0, 1);

    // ---- node blocks of S and the right-hand sides (one thread per node j = 1..M-1) ----
    for (int j = 1 + tid; j < M; j += 64) {
        const double ia = iT[j - 1], ib = iT[j];
        const double ia2 = ia * ia, ib2 = ib * ib, ia3 = ia2 * ia, ib3 = ib2 * ib, ia4 = ia2 * ia2, ib4 = ib2 * ib2;
        nD[3 * j + 0] = 192.0 * (ia3 + ib3);
        nD[3 * j + 1] = 36.0 * (ib2 - ia2);
        nD[3 * j + 2] = 9.0 * (ia + ib);
        const double l11 = 168.0 * ia3, l12 = 24.0 * ia2, l21 = -24.0 * ia2, l22 = -3.0 * ia;
        nL[4 * j + 0] = l11, nL[4 * j + 1] = l12, nL[4 * j + 2] = l21, nL[4 * j + 3] = l22;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double da = P[3 * j + d] - P[3 * (j - 1) + d], db = P[3 * (j + 1) + d] - P[3 * j + d];
            double r0 = 360.0 * (da * ia4 + db * ib4);
            double r1 = 60.0 * (db * ib3 - da * ia3);
            if (j == 1) {  // known neighbour: head velocity / acceleration
                const double hv = Y[d], ha = Y[3 + d];
                r0 -= l11 * hv + l12 * ha;
                r1 -= l21 * hv + l22 * ha;
            }
            if (j == M - 1) {  // known neighbour: tail velocity / acceleration through U_{M-1} = L_M^T with Ta := T_{M-1}
                const double tv = Y[6 * M + d], ta = Y[6 * M + 3 + d];
                r0 -= (168.0 * ib3) * tv + (-24.0 * ib2) * ta;
                r1 -= (24.0 * ib2) * tv + (-3.0 * ib) * ta;
            }
            nR[6 * j + d] = r0;
            nR[6 * j + 3 + d] = r1;
        }
    }
    __syncthreads();
    PHASE_MARK(0, 2);

    double* fac = bv.lu + (size_t)NODE_F * p0;
    if (M - 1 <= PCR_MAX_NODES) {
        // ---- the reduced system by parallel cyclic reduction (one thread per node).  K3 solves the adjoint system the same way:
        //      it gets the node blocks themselves (8 doubles per node: D | L | unused) instead of LDL^T factors. ----
        for (int j = 1 + tid; j < M; j += 64) {
            double* f = fac + NODE_F * j;
            f[0] = nD[3 * j], f[1] = nD[3 * j + 1], f[2] = nD[3 * j + 2];
            f[3] = nL[4 * j], f[4] = nL[4 * j + 1], f[5] = nL[4 * j + 2], f[6] = nL[4 * j + 3];
        }
        PHASE_MARK(0, 3);
        pcr_solve<3, 4>(nD, nL, nR, nD, Y, M - 1, tid);  // scratch: the node-block arrays themselves (17 M doubles >= 13 (M - 1))
    } else
    // ---- long trajectories: block LDL^T with the forward substitution fused in, then the back substitution: lanes 0..2 of
    //      warp 0 ----
    if (warp == 0 && lane < 3 && M > 1) {
        const int d = lane;
        double d11 = nD[3], d12 = nD[4], d22 = nD[5];  // pivot block of node 1
        double z0 = nR[6 + d], z1 = nR[9 + d];
        // operands of the next node, loaded one step ahead of the dependent chain
        double nl11 = 0, nl12 = 0, nl21 = 0, nl22 = 0, nd11 = 0, nd12 = 0, nd22 = 0, nr0 = 0, nr1 = 0;
        if (M > 2) {
            nl11 = nL[8], nl12 = nL[9], nl21 = nL[10], nl22 = nL[11];
            nd11 = nD[6], nd12 = nD[7], nd22 = nD[8];
            nr0 = nR[12 + d], nr1 = nR[15 + d];
        }
        for (int j = 1; j < M; j++) {
            // Dependent chain of the recurrence, kept as short as it gets: 1/det (7 levels) runs beside N = L adj(D) L^T
            // (4 levels), then ONE fma per entry:  D'_{j+1} = D_{j+1} - N / det.  Everything else hangs off the side.
            const double det = fma(d11, d22, -(d12 * d12));
            const double id = fast_rcp3(det);
            const double l11 = nl11, l12 = nl12, l21 = nl21, l22 = nl22;  // L_{j+1} (zeros after the last node)
            const double c11 = nd11, c12 = nd12, c22 = nd22, r0 = nr0, r1 = nr1;
            const double t11 = fma(l11, d22, -(l12 * d12)), t12 = fma(l12, d11, -(l11 * d12));  // L adj(D_j)
            const double t21 = fma(l21, d22, -(l22 * d12)), t22 = fma(l22, d11, -(l21 * d12));
            const double n11 = fma(t11, l11, t12 * l12), n12 = fma(t11, l21, t12 * l22), n22 = fma(t21, l21, t22 * l22);
            const double i11 = d22 * id, i12 = -(d12 * id), i22 = d11 * id;  // inverse pivot block
            d11 = fma(-n11, id, c11), d12 = fma(-n12, id, c12), d22 = fma(-n22, id, c22);
            if (j + 2 < M) {
                const int q = j + 2;
                nl11 = nL[4 * q], nl12 = nL[4 * q + 1], nl21 = nL[4 * q + 2], nl22 = nL[4 * q + 3];
                nd11 = nD[3 * q], nd12 = nD[3 * q + 1], nd22 = nD[3 * q + 2];
                nr0 = nR[6 * q + d], nr1 = nR[6 * q + 3 + d];
            }
            const double w0 = fma(i11, z0, i12 * z1), w1 = fma(i12, z0, i22 * z1);  // w_j = Di_j z_j
            nR[6 * j + d] = w0;
            nR[6 * j + 3 + d] = w1;
            z0 = fma(-l12, w1, fma(-l11, w0, r0));  // r_{j+1} - L_{j+1} w_j
            z1 = fma(-l22, w1, fma(-l21, w0, r1));
            // The three lanes hold bit-identical factors: all of them store (same value, same address) instead of
            // diverging on lane 0 — a divergent tail of a dozen instructions per node costs more than the chain itself.
            fac[NODE_F * j + 0] = i11, fac[NODE_F * j + 1] = i12, fac[NODE_F * j + 2] = i22;
            if (j + 1 < M) {
                const double g11 = t11 * id, g12 = t12 * id, g21 = t21 * id, g22 = t22 * id;  // G_{j+1} = L_{j+1} Di_j
                double* gs = nG + 4 * (j + 1);
                gs[0] = g11, gs[1] = g12, gs[2] = g21, gs[3] = g22;
                double* gg = fac + NODE_F * (j + 1) + 3;
                gg[0] = g11, gg[1] = g12, gg[2] = g21, gg[3] = g22;
            }
        }
        __syncwarp(0x7u);
        PHASE_MARK(0, 3);
        // y_j = w_j - G_{j+1}^T y_{j+1}
        double y0 = nR[6 * (M - 1) + d], y1 = nR[6 * (M - 1) + 3 + d];
        Y[6 * (M - 1) + d] = y0;
        Y[6 * (M - 1) + 3 + d] = y1;
        double g0 = 0, g1 = 0, g2 = 0, g3 = 0, w0 = 0, w1 = 0;  // operands of node j, loaded one step ahead of the chain
        if (M > 2) {
            const double* g = nG + 4 * (M - 1);
            g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3];
            w0 = nR[6 * (M - 2) + d], w1 = nR[6 * (M - 2) + 3 + d];
        }
        for (int j = M - 2; j >= 1; j--) {
            const double t0 = fma(-g2, y1, fma(-g0, y0, w0));
            const double t1 = fma(-g3, y1, fma(-g1, y0, w1));
            if (j > 1) {
                const double* g = nG + 4 * j;
                g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3];
                w0 = nR[6 * (j - 1) + d], w1 = nR[6 * (j - 1) + 3 + d];
            }
            y0 = t0, y1 = t1;
            Y[6 * j + d] = y0;
            Y[6 * j + 3 + d] = y1;
        }
    }
    __syncthreads();
    PHASE_MARK(0, 4);

    // ---- Hermite expansion: c_i from the end states of piece i, one thread per (piece, axis) ----
    for (int task = tid; task < 3 * M; task += 64) {
        const int i = task / 3, d = task - 3 * i;
        const double t1 = iT[i], t2 = t1 * t1, t3 = t2 * t1, t4 = t2 * t2, t5 = t4 * t1;
        const double P0 = P[3 * i + d], P1 = P[3 * i + 3 + d];
        const double V0 = Y[6 * i + d], A0 = Y[6 * i + 3 + d], V1 = Y[6 * i + 6 + d], A1 = Y[6 * i + 9 + d];
        const double dP = P1 - P0;
        const double c3 = (10.0 * dP) * t3 - (4.0 * V1 + 6.0 * V0) * t2 - (3.0 * A0 - A1) * (0.5 * t1);
        const double c4 = (-15.0 * dP) * t4 + (7.0 * V1 + 8.0 * V0) * t3 + (3.0 * A0 - 2.0 * A1) * (0.5 * t2);
        const double c5 = (6.0 * dP) * t5 - (3.0 * (V1 + V0)) * t4 + (A1 - A0) * (0.5 * t3);
        double* cg = bv.c + 18 * (size_t)(p0 + i) + d;
        cg[0] = P0, cg[3] = V0, cg[6] = 0.5 * A0, cg[9] = c3, cg[12] = c4, cg[15] = c5;
    }
    CTA_SPAN(0, 1);
    PHASE_MARK(0, 5);
}

// K3 (fast) -----------------------------------------------------------------------------------------
// One trajectory per CTA of two warps.  dJ/dc, c and K1's factors arrive by three 1-D bulk copies (TMA unit, one mbarrier)
// while the index tables load; then: Hermite adjoint per piece (in place over dJ/dc), the node right-hand sides, the two
// substitution sweeps with K1's factors (lanes 0..2 of warp 0, a chain of two FMAs per node) while warp 1 sums the cost in
// the reference's order, the closed-form dJ/dT and dJ/dq, then mergeToCoarseGradT / addLayerTGrad / addLayerPGrad
// (CPU.hpp:946, :816, :897).
// smem (doubles): Ts[M] | iT[M] | SE[18M] (dJ/dc in, then per piece: start part (P,V,A) 9 | end part 9) | CS[18M] (c) |
//                 fc[8M] (after the solve: kt[3M] | gT[M] | gq[3M]) | Z[6M] | gTd[M] | gTc[coarseN+1] | tfac[coarseN] |
//                 ivs, ioff [coarseN ints each] | wpi[3M ints] | mbarrier
__global__ void __launch_bounds__(64)
    minco_backward_kernel(BatchView bv, const double* __restrict__ xin, double* __restrict__ cost_out,
                          double* __restrict__ grad_out) {
    extern __shared__ __align__(16) double sm[];
    const int b = bv.slot_list ? bv.slot_list[blockIdx.x] : bv.sub_traj0 + blockIdx.x;
    if (b < 0) return;  // empty solver slot
    if (!bv.slot_list && bv.active && !bv.active[b]) return;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int p0 = bv.piece_off[b];
    const int M = bv.piece_off[b + 1] - p0;
    const int co0 = bv.coarse_off[b];
    const int coarseN = bv.coarse_off[b + 1] - co0;
    const DevParams par = bv.par[b];
    const double* x = xin + bv.x_off[b];
    double* grad = grad_out + bv.x_off[b];
    const int dimT = bv.dimT[b];
    const int nx = bv.x_off[b + 1] - bv.x_off[b];
    double* Ts = sm;
    double* iT = Ts + M;
    double* SE = iT + M;
    double* CS = SE + 18 * (size_t)M;
    double* fc = CS + 18 * (size_t)M;
    double* Z = fc + NODE_F * (size_t)M;
    double* gTd = Z + 6 * (size_t)M;
    double* kt = fc;  // kt[3M] | gT[M] | gq[3M] reuse the factor block: written only after the sweeps are done with it
    double* gT = kt + 3 * (size_t)M;
    double* gq = gT + M;
    double* gTc = gTd + M;
    double* tfac = gTc + coarseN + 1;
    int* ivs = reinterpret_cast<int*>(tfac + coarseN);
    int* ioff = ivs + coarseN;
    int* wpi = ioff + coarseN;
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(sm + backward_smem_doubles(M, coarseN) - 1);
    CTA_SPAN(2, 0);
    PHASE_MARK(2, 0);

    const double* tailst = bv.headtail + 18 * (size_t)b + 9;  // tail p (0..2), v (3..5), a (6..8)
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (tid == 0) {
        const unsigned bc = 144u * (unsigned)M, bf = 8u * NODE_F * (unsigned)M;
        mbar_expect_tx(mbar, 2 * bc + bf);
        bulk_g2s(SE, bv.gdC + 18 * (size_t)p0, bc, mbar);
        bulk_g2s(CS, bv.c + 18 * (size_t)p0, bc, mbar);
        bulk_g2s(fc, bv.lu + (size_t)NODE_F * p0, bf, mbar);
    }
    // what the epilogue streams (x, the vertex range): one bulk L2 prefetch each, long before it is needed
    if (tid == 32) bulk_prefetch_l2(x, sizeof(double) * (size_t)nx);
    if (tid == 33 && M > 1) {
        const double* v0 = bv.verts + 3 * (size_t)bv.wp_vert_off[p0];
        const double* v1 = bv.verts + 3 * ((size_t)bv.wp_vert_off[p0 + M - 2] + bv.wp_vert_cnt[p0 + M - 2]);
        bulk_prefetch_l2(v0, (size_t)((const char*)v1 - (const char*)v0));
    }
    for (int i = tid; i < M; i += 64) {
        const double t = bv.T[p0 + i];
        Ts[i] = t, iT[i] = 1.0 / t;
    }
    for (int i = tid; i < M - 1; i += 64) {
        wpi[i] = bv.wp_vert_cnt[p0 + i] - 1;
        wpi[M + i] = bv.wp_xi_off[p0 + i];
        wpi[2 * M + i] = bv.wp_vert_off[p0 + i];
    }
    if (warp == 1) {
        // ---- cost = sum energy + sum penalty + rho * sum T (CPU.hpp:542,406,988) by warp 1 while the bulk copies
        //      fly: strided partial sums and a fixed shuffle tree (the reference sums sequentially; reference-order mode does
        //      too), time-map derivative, interval offsets ----
        double se = 0.0, sp = 0.0, sT = 0.0;
        for (int i = lane; i < M; i += 32) {
            se += bv.energy[p0 + i];
            sp += bv.pen[p0 + i];
        }
        for (int i = lane; i < coarseN; i += 32) ivs[i] = bv.intervals[co0 + i];
        if (par.softT) {
            for (int i = lane; i < coarseN; i += 32) {  // time map and its derivative (CPU.hpp:639-641, 830-838)
                const double t = x[i];
                double Tv, dv;
                if (par.c2) {
                    Tv = c2_forward(t);
                    if (t > 0)
                        dv = t + 1.0;
                    else {
                        const double den = (0.5 * t - 1.0) * t + 1.0;
                        dv = (1.0 - t) / (den * den);
                    }
                } else {
                    Tv = exp(t);
                    dv = Tv;
                }
                gTc[i] = Tv;
                tfac[i] = dv;
                sT += Tv;
            }
        } else {
            for (int i = lane; i < M; i += 32) sT += bv.T[p0 + i];  // fixed total time: rho = 0 anyway
        }
        __syncwarp();
        {
            int carry = 0;  // exclusive prefix sum of the interval counts, 32 at a time
            for (int base = 0; base < coarseN; base += 32) {
                const int i = base + lane;
                const int v = i < coarseN ? ivs[i] : 0;
                int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (i < coarseN) ioff[i] = carry + incl - v;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, o);
            sp += __shfl_xor_sync(0xffffffffu, sp, o);
            sT += __shfl_xor_sync(0xffffffffu, sT, o);
        }
        if (lane == 0) cost_out[b] = (se + sp) + par.rho * sT;
    }
    mbar_wait(mbar, 0);
    __syncthreads();
    PHASE_MARK(2, 1);
    // ---- Hermite adjoint per piece: (P0b, V0b, A0b | P1b, V1b, A1b) = H(T_i)^T dJ/dc_i, and dc/dT at fixed end states ----
    for (int i = tid; i < M; i += 64) {
        const double t = Ts[i];
        const double i1 = iT[i], i2 = i1 * i1, i3 = i2 * i1, i4 = i2 * i2, i5 = i4 * i1;
        const double* ci = CS + 18 * (size_t)i;
        double* gi = SE + 18 * (size_t)i;
        double gv[18];
#pragma unroll
        for (int q = 0; q < 18; q++) gv[q] = gi[q];
        double s = 0.0;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double g0 = gv[d], g1 = gv[3 + d], g2 = gv[6 + d], g3 = gv[9 + d], g4 = gv[12 + d], g5 = gv[15 + d];
            const double dP = (10.0 * i3) * g3 - (15.0 * i4) * g4 + (6.0 * i5) * g5;
            const double V0b = g1 - (6.0 * i2) * g3 + (8.0 * i3) * g4 - (3.0 * i4) * g5;
            const double V1b = (7.0 * i3) * g4 - (4.0 * i2) * g3 - (3.0 * i4) * g5;
            const double A0b = 0.5 * g2 - (1.5 * i1) * g3 + (1.5 * i2) * g4 - (0.5 * i3) * g5;
            const double A1b = (0.5 * i1) * g3 - i2 * g4 + (0.5 * i3) * g5;
            gi[d] = g0 - dP, gi[3 + d] = V0b, gi[6 + d] = A0b;       // start part
            gi[9 + d] = dP, gi[12 + d] = V1b, gi[15 + d] = A1b;      // end part
            const double V1 = (i + 1 < M) ? ci[18 + 3 + d] : tailst[3 + d];
            const double A1 = (i + 1 < M) ? 2.0 * ci[18 + 6 + d] : tailst[6 + d];
            const double J1 = 6.0 * ci[9 + d] + (24.0 * t) * ci[12 + d] + (60.0 * t * t) * ci[15 + d];
            s -= dP * V1 + V1b * A1 + A1b * J1;
        }
        gTd[i] = s;
    }
    __syncthreads();
    // ---- node right-hand sides: dJ/dy_j = end part of piece j-1 + start part of piece j ----
    for (int e = tid; e < 6 * (M - 1); e += 64) {
        const int j = 1 + e / 6, r = e - 6 * (j - 1);  // r = 0..2: velocity, 3..5: acceleration
        Z[6 * j + r] = SE[18 * (j - 1) + 12 + r] + SE[18 * j + 3 + r];
    }
    __syncthreads();
    PHASE_MARK(2, 2);
    const bool use_pcr = M - 1 <= PCR_MAX_NODES;
    // ---- S mu = dJ/dy: the same cyclic reduction as K1 on the node blocks K1 left in the factor slot (in place over Z) ----
    if (use_pcr) pcr_solve<NODE_F, NODE_F>(fc, fc + 3, Z, fc, Z, M - 1, tid);  // scratch: fc | Z (14 M doubles >= 13 (M - 1))
    if (warp == 0) {
        // ---- long trajectories: with K1's LDL^T factors: z_j = b_j - G_j z_{j-1},  w_j = Di_j z_j,  mu_j = w_j - G_{j+1}^T mu_{j+1} ----
        if (lane < 3 && M > 1 && !use_pcr) {
            const int d = lane;
            // both sweeps are chains of two FMAs per node; the operands of the next node are loaded into registers before
            // this node's results are stored (the stores would otherwise fence the loads: same shared array)
            double z0 = Z[6 + d], z1 = Z[9 + d];
            double f0 = fc[NODE_F], f1 = fc[NODE_F + 1], f2 = fc[NODE_F + 2];
            double g0 = 0, g1 = 0, g2 = 0, g3 = 0, b0 = 0, b1 = 0;
            if (M > 2) {
                const double* f = fc + 2 * NODE_F;
                g0 = f[3], g1 = f[4], g2 = f[5], g3 = f[6];
                b0 = Z[12 + d], b1 = Z[15 + d];
            }
            for (int j = 1; j < M; j++) {
                const double w0 = f0 * z0 + f1 * z1, w1 = f1 * z0 + f2 * z1;  // w_j = Di_j z_j
                const double t0 = fma(-g1, z1, fma(-g0, z0, b0)), t1 = fma(-g3, z1, fma(-g2, z0, b1));  // z_{j+1} = b_{j+1} - G_{j+1} z_j
                if (j + 1 < M) {
                    const double* f = fc + NODE_F * (j + 1);
                    f0 = f[0], f1 = f[1], f2 = f[2];
                }
                if (j + 2 < M) {
                    const double* f = fc + NODE_F * (j + 2);
                    g0 = f[3], g1 = f[4], g2 = f[5], g3 = f[6];
                    b0 = Z[6 * (j + 2) + d], b1 = Z[6 * (j + 2) + 3 + d];
                }
                Z[6 * j + d] = w0;
                Z[6 * j + 3 + d] = w1;
                z0 = t0, z1 = t1;
            }
            double m0 = Z[6 * (M - 1) + d], m1 = Z[6 * (M - 1) + 3 + d];
            if (M > 2) {
                const double* g = fc + NODE_F * (M - 1) + 3;
                g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3];
                b0 = Z[6 * (M - 2) + d], b1 = Z[6 * (M - 2) + 3 + d];
            }
            for (int j = M - 2; j >= 1; j--) {
                const double t0 = fma(-g2, m1, fma(-g0, m0, b0));  // mu_j = w_j - G_{j+1}^T mu_{j+1}
                const double t1 = fma(-g3, m1, fma(-g1, m0, b1));
                if (j > 1) {
                    const double* g = fc + NODE_F * j + 3;
                    g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3];
                    b0 = Z[6 * (j - 1) + d], b1 = Z[6 * (j - 1) + 3 + d];
                }
                m0 = t0, m1 = t1;
                Z[6 * j + d] = m0;
                Z[6 * j + 3 + d] = m1;
            }
        }
    }
    __syncthreads();
    PHASE_MARK(2, 4);
    // ---- dJ/dT (energy + penalty partial from K2, the two closed forms above, rho of CPU.hpp:989) and crackle(ct) ----
    for (int i = tid; i < M; i += 64) {
        const double t = Ts[i], i1 = iT[i], i2 = i1 * i1, i3 = i2 * i1, i4 = i2 * i2;
        const double* ci = CS + 18 * (size_t)i;
        double s = 0.0;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double mV0 = (i > 0) ? Z[6 * i + d] : 0.0, mA0 = (i > 0) ? Z[6 * i + 3 + d] : 0.0;
            const double mV1 = (i + 1 < M) ? Z[6 * i + 6 + d] : 0.0, mA1 = (i + 1 < M) ? Z[6 * i + 9 + d] : 0.0;
            const double ct3 = -((4.0 * mV1 + 6.0 * mV0) * i2) - (3.0 * mA0 - mA1) * (0.5 * i1);
            const double ct4 = (7.0 * mV1 + 8.0 * mV0) * i3 + (3.0 * mA0 - 2.0 * mA1) * (0.5 * i2);
            const double ct5 = (mA1 - mA0) * (0.5 * i3) - (3.0 * (mV1 + mV0)) * i4;
            const double c3 = ci[9 + d], c4 = ci[12 + d], c5 = ci[15 + d];
            const double jc = 6.0 * c3 + (24.0 * t) * c4 + (60.0 * t * t) * c5;
            const double sc = 24.0 * c4 + (120.0 * t) * c5;
            const double kc = 120.0 * c5;
            const double jt = 6.0 * ct3 + (24.0 * t) * ct4 + (60.0 * t * t) * ct5;
            const double st = 24.0 * ct4 + (120.0 * t) * ct5;
            const double ktv = 120.0 * ct5;
            const double V1 = (i + 1 < M) ? ci[18 + 3 + d] : tailst[3 + d];
            const double A1 = (i + 1 < M) ? 2.0 * ci[18 + 6 + d] : tailst[6 + d];
            s += jc * jt - st * A1 - sc * mA1 + ktv * V1 + kc * mV1;
            kt[3 * i + d] = ktv;
        }
        gT[i] = bv.gdT[p0 + i] + (gTd[i] + s) + par.rho;
    }
    __syncthreads();
    // ---- dJ/dq (addPropCtoP): waypoint w = node w+1 ----
    for (int e = tid; e < 3 * (M - 1); e += 64) {
        const int w = e / 3, d = e - 3 * w;
        gq[e] = (SE[18 * w + 9 + d] + SE[18 * (w + 1) + d]) + (kt[3 * (w + 1) + d] - kt[3 * w + d]);
    }
    __syncthreads();
    PHASE_MARK(2, 5);
    // ---- mergeToCoarseGradT (CPU.hpp:946-959): mean over the pieces of each polytope; one thread per polytope ----
    for (int i = tid; i < coarseN; i += 64) {
        const int iv = ivs[i], offset = ioff[i];
        double sgt = 0.0;
        for (int a = 0; a < iv; a++) sgt += gT[offset + a];
        const double gm = sgt / iv;
        // ---- addLayerTGrad (CPU.hpp:816-894), free total time: the factor was computed above ----
        if (par.softT) {
            if (i < dimT) grad[i] = gm * tfac[i];
        } else
            gTc[i] = gm;
    }
    if (!par.softT) {
        __syncthreads();
        if (tid == 0) {
            const int Ms1 = dimT;
            const double gTail = par.sumT * gTc[Ms1];
            double expTauSum = 0.0, gFreeDotExpTau = 0.0;
            for (int i = 0; i < Ms1; i++) {
                const double t = x[i];
                const double gFree = par.sumT * gTc[i];
                double eTau;
                if (par.c2) {
                    if (t > 0)
                        eTau = (0.5 * t + 1.0) * t + 1.0;
                    else
                        eTau = 1.0 / ((0.5 * t - 1.0) * t + 1.0);
                } else
                    eTau = exp(t);
                expTauSum += eTau;
                gFreeDotExpTau += eTau * gFree;
            }
            const double den = expTauSum + 1.0;
            for (int i = 0; i < Ms1; i++) {
                const double t = x[i];
                const double gFree = par.sumT * gTc[i];
                double dE;
                if (par.c2) {
                    if (t > 0)
                        dE = t + 1.0;
                    else {
                        const double ds = (0.5 * t - 1.0) * t + 1.0;
                        dE = (1.0 - t) / (ds * ds);
                    }
                } else
                    dE = exp(t);
                grad[i] = (gFree - gTail) * dE / den - (gFreeDotExpTau - gTail * expTauSum) * dE / (den * den);
            }
        }
    }
    PHASE_MARK(2, 6);
    // ---- addLayerPGrad (CPU.hpp:897-928) ----
    {   // four lanes per waypoint, ONE streaming pass: with s_a = V_{a+1}.gq, q1 = 1 + |xi|^2 and W = sum_a V_{a+1} xi_a^2
        // (both stored by K1):  gd_a = 4 s_a xi_a / q1,  gdot = 4 (W.gq) / q1,  d/dxi_a = xi_a (8 s_a / q1^2 - 16 (W.gq) / q1^3)
        const int grp = tid >> 2, l4 = tid & 3;
        for (int i = grp; i < M - 1; i += 16) {
            const int k = wpi[i];
            const double* V = bv.verts + 3 * ((size_t)wpi[2 * M + i] + 1);
            const double* xi = x + dimT + wpi[M + i];
            double* gxi = grad + dimT + wpi[M + i];
            const double* ww = bv.wpw + 4 * (size_t)(p0 + i);
            const double gx = gq[3 * i], gy = gq[3 * i + 1], gz = gq[3 * i + 2];
            const double wg = ww[0] * gx + ww[1] * gy + ww[2] * gz;
            const double iq = 1.0 / (1.0 + ww[3]), iq2 = iq * iq;
            const double c1 = 8.0 * iq2, c2 = 16.0 * wg * iq2 * iq;
            for (int a0 = l4; a0 < k; a0 += 20) {  // operands first (all loads in flight together), then the stores
                double xa[5], va[5], vb[5], vc[5];
#pragma unroll
                for (int u = 0; u < 5; u++) {
                    const int a = a0 + 4 * u;
                    if (a < k) xa[u] = __ldg(xi + a), va[u] = __ldg(V + 3 * a), vb[u] = __ldg(V + 3 * a + 1), vc[u] = __ldg(V + 3 * a + 2);
                }
#pragma unroll
                for (int u = 0; u < 5; u++) {
                    const int a = a0 + 4 * u;
                    if (a < k) gxi[a] = xa[u] * (c1 * (va[u] * gx + vb[u] * gy + vc[u] * gz) - c2);
                }
            }
        }
    }
    CTA_SPAN(2, 1);
    PHASE_MARK(2, 7);
}
#endif  // !FRX_EXACT (reduced-system kernels)

// ------------------------------------------------------------------------------------------------
// Host-side launchers
// ------------------------------------------------------------------------------------------------
#ifdef FRX_EXACT
size_t forward_smem_bytes(int M, int coarseN) {
    const size_t ust = (size_t)6 * M * UST, tc = (size_t)coarseN + 1;
    return sizeof(double) * ((size_t)RING * BW + 32 + (ust > tc ? ust : tc));
}
size_t backward_smem_bytes(int M, int coarseN) {
    return sizeof(double) * ((size_t)NBLK * BLKD + (size_t)18 * M + (size_t)4 * M + coarseN + 1 + (size_t)2 * M + coarseN + (coarseN + 1) + 1 +
                             (3 * (size_t)M + 1) / 2 + 2);
}
#else
size_t forward_smem_bytes(int M, int coarseN) {  // see the layout comment of minco_forward_kernel
    return sizeof(double) * ((size_t)coarseN + 1 + (size_t)28 * M + 9 + (3 * (size_t)M + 1) / 2 + 2);
}
size_t backward_smem_bytes(int M, int coarseN) { return sizeof(double) * backward_smem_doubles(M, coarseN); }
#endif

PenaltyLaunch plan_penalty(const BatchView& bv) {
    PenaltyLaunch pl;
    const int S = bv.max_samples;
    // samples per thread: aim for <= 16 threads per piece; among the candidates pick the split with the best combined
    // efficiency = (used sample slots / allotted slots) x (used lanes / CTA lanes)   (S = K+1 is odd, so something is wasted)
    int best_spt = (S + 15) / 16;
    double best_eff = -1.0;
    for (int spt = (S + 15) / 16; spt <= (S + 15) / 16 + 2; spt++) {
        const int tpp = (S + spt - 1) / spt;
        const int ppc = 64 / tpp > 0 ? 64 / tpp : 1;
        const double eff = ((double)S / (tpp * spt)) * ((double)(ppc * tpp) / 64.0);
        if (eff > best_eff + 1e-12) best_eff = eff, best_spt = spt;
    }
    pl.spt = best_spt;
    if (const char* e = getenv("FRX_PEN_SPT")) pl.spt = atoi(e) > 0 ? atoi(e) : pl.spt;  // tuning override
    pl.tpp = (S + pl.spt - 1) / pl.spt;
    pl.threads = 64;
    pl.ppc = pl.threads / pl.tpp;
    if (pl.ppc < 1) pl.ppc = 1;
    pl.grid = (bv.total_pieces + pl.ppc - 1) / pl.ppc;
    pl.smem = sizeof(double) * ((size_t)pl.ppc * 60 + (size_t)pl.ppc * 6 * bv.max_faces + (size_t)pl.threads * 20) +
              16 * (size_t)pl.ppc * bv.max_faces + sizeof(double) * pl.ppc;
    return pl;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per function and per device, and every launch must request no more than
// the current value: batches of different shapes coexist (frx.h: distinct batches are independent), so the limit only
// ever grows, under a lock (two threads may create batches concurrently).
cudaError_t configure_kernels(size_t fwd_smem, size_t bwd_smem, size_t pen_smem) {
    static std::mutex mu;
    static size_t cur[64][3];  // [device][kernel], zero-initialised
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(mu);
    const size_t want[3] = {fwd_smem, bwd_smem, pen_smem};
    const void* fn[3] = {(const void*)minco_forward_kernel, (const void*)minco_backward_kernel, (const void*)penalty_kernel};
    for (int k = 0; k < 3; k++) {
        if (want[k] <= cur[dev][k]) continue;
        e = cudaFuncSetAttribute(fn[k], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want[k]);
        if (e != cudaSuccess) return e;
        cur[dev][k] = want[k];
    }
    return cudaSuccess;
}

static void launch_penalty(const BatchView& bv, const PenaltyLaunch& pl, cudaStream_t st) {
#ifdef FRX_EXACT
    const int nt = bv.slot_list ? bv.n_slots : bv.sub_traj1 - bv.sub_traj0;
    penalty_exact_kernel<<<(nt + 31) / 32, 32, 0, st>>>(bv);
#else
    const int grid = bv.slot_list ? bv.n_slots * ((bv.max_pieces + pl.ppc - 1) / pl.ppc) : (bv.sub_piece1 - bv.sub_piece0 + pl.ppc - 1) / pl.ppc;
    penalty_kernel<<<grid, pl.threads, pl.smem, st>>>(bv, pl);
#endif
}

// One batched evaluation.  The banded kernels K1/K3 are latency-bound (a dependent chain per trajectory) and the
// penalty kernel K2 is throughput-bound, so the batch is cut into sub-batches that are software-pipelined over two
// streams: K1 of sub-batch i+1 runs under K2 of sub-batch i, K3 of sub-batch i under K2 of sub-batch i+1.
// `piece_off_host` = host copy of bv.piece_off.  aux/ev*: helper stream and events owned by the batch.
cudaError_t launch_eval_pipelined(const BatchView& bv0, const PenaltyLaunch& pl, size_t fwd_smem, size_t bwd_smem,
                                  const double* x_dev, double* cost_dev, double* grad_dev, cudaStream_t st, cudaStream_t aux,
                                  cudaEvent_t ev_fork, cudaEvent_t ev_k1, cudaEvent_t ev_join, const int* piece_off_host,
                                  int nsub) {
    const int B = bv0.B;
    if (nsub < 2 || B < 2 * nsub || !aux || bv0.slot_list) {  // solver slots: always one launch over the slots
        BatchView bv = bv0;
        bv.sub_traj0 = 0, bv.sub_traj1 = B, bv.sub_piece0 = 0, bv.sub_piece1 = bv.total_pieces;
        const int grid = bv.slot_list ? bv.n_slots : B;  // slot mode: one CTA per solver slot
        minco_forward_kernel<<<grid, 64, fwd_smem, st>>>(bv, x_dev);
        launch_penalty(bv, pl, st);
        minco_backward_kernel<<<grid, 64, bwd_smem, st>>>(bv, x_dev, cost_dev, grad_dev);
        return cudaGetLastError();
    }
    cudaEventRecord(ev_fork, st);
    cudaStreamWaitEvent(aux, ev_fork, 0);
    for (int i = 0; i < nsub; i++) {
        const int t0 = (int)((long long)B * i / nsub), t1 = (int)((long long)B * (i + 1) / nsub);
        BatchView bv = bv0;
        bv.sub_traj0 = t0, bv.sub_traj1 = t1;
        bv.sub_piece0 = piece_off_host[t0];
        bv.sub_piece1 = piece_off_host[t1];
        cudaStream_t s = (i & 1) ? aux : st;
        if (i == 1) cudaStreamWaitEvent(aux, ev_k1, 0);  // stagger: the second stream starts one K1 later
        minco_forward_kernel<<<t1 - t0, 64, fwd_smem, s>>>(bv, x_dev);
        if (i == 0) cudaEventRecord(ev_k1, st);
        launch_penalty(bv, pl, s);
        minco_backward_kernel<<<t1 - t0, 64, bwd_smem, s>>>(bv, x_dev, cost_dev, grad_dev);
    }
    cudaEventRecord(ev_join, aux);
    cudaStreamWaitEvent(st, ev_join, 0);
    return cudaGetLastError();
}

// The three kernels for the trajectory window [t0, t1) on one stream (used by the chunked host-buffer path, where the
// host<->device copies of one chunk overlap the kernels of the other).
cudaError_t launch_eval_window(const BatchView& bv0, const PenaltyLaunch& pl, size_t fwd_smem, size_t bwd_smem, const double* x_dev,
                               double* cost_dev, double* grad_dev, cudaStream_t st, int t0, int t1, const int* piece_off_host) {
    BatchView bv = bv0;
    bv.sub_traj0 = t0, bv.sub_traj1 = t1;
    bv.sub_piece0 = piece_off_host[t0];
    bv.sub_piece1 = piece_off_host[t1];
    minco_forward_kernel<<<t1 - t0, 64, fwd_smem, st>>>(bv, x_dev);
    launch_penalty(bv, pl, st);
    minco_backward_kernel<<<t1 - t0, 64, bwd_smem, st>>>(bv, x_dev, cost_dev, grad_dev);
    return cudaGetLastError();
}

cudaError_t launch_eval(const BatchView& bv0, const PenaltyLaunch& pl, size_t fwd_smem, size_t bwd_smem,
                        const double* x_dev, double* cost_dev, double* grad_dev, cudaStream_t st) {
    return launch_eval_pipelined(bv0, pl, fwd_smem, bwd_smem, x_dev, cost_dev, grad_dev, st, nullptr, nullptr, nullptr, nullptr,
                                 nullptr, 1);
}

// Per-kernel device times of `reps` evaluations: CUDA events on the launching stream around each
// of the three kernels.  ms[3] = average duration of K1, K2, K3.
cudaError_t profile_eval(const BatchView& bv, const PenaltyLaunch& pl, size_t fwd_smem, size_t bwd_smem,
                         const double* x_dev, double* cost_dev, double* grad_dev, cudaStream_t st, int reps, double* ms) {
    BatchView bvw = bv;
    bvw.sub_traj0 = 0, bvw.sub_traj1 = bv.B, bvw.sub_piece0 = 0, bvw.sub_piece1 = bv.total_pieces;
    cudaEvent_t ev[4];
    for (auto& e : ev) cudaEventCreate(&e);
    ms[0] = ms[1] = ms[2] = 0.0;
    cudaError_t err = cudaSuccess;
    for (int r = 0; r < reps && err == cudaSuccess; r++) {
        cudaEventRecord(ev[0], st);
        minco_forward_kernel<<<bv.B, 64, fwd_smem, st>>>(bvw, x_dev);
        cudaEventRecord(ev[1], st);
        launch_penalty(bvw, pl, st);
        cudaEventRecord(ev[2], st);
        minco_backward_kernel<<<bv.B, 64, bwd_smem, st>>>(bvw, x_dev, cost_dev, grad_dev);
        cudaEventRecord(ev[3], st);
        err = cudaStreamSynchronize(st);
        for (int k = 0; k < 3 && err == cudaSuccess; k++) {
            float t = 0.f;
            err = cudaEventElapsedTime(&t, ev[k], ev[k + 1]);
            ms[k] += t / reps;
        }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    return err;
}

// fp64 FMA peak micro-benchmark: 8 independent DFMA chains per thread.
__global__ void dfma_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c), a1 = fma(a1, m, c), a2 = fma(a2, m, c), a3 = fma(a3, m, c);
        a4 = fma(a4, m, c), a5 = fma(a5, m, c), a6 = fma(a6, m, c), a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
double measure_fp64_tflops() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    double* out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        dfma_peak_kernel<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 8.0 * iters * (double)blocks * threads;
        best = fmax(best, fl / (ms * 1e-3) * 1e-12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return best;
}

cudaError_t fetch_phase_clocks(long long* out96) { return cudaMemcpyFromSymbol(out96, g_phase_clk, sizeof(long long) * 96); }
cudaError_t fetch_cta_spans(unsigned long long* out) { return cudaMemcpyFromSymbol(out, g_cta_span, sizeof(unsigned long long) * 2 * 3 * 4096); }
}  // namespace FRX_IMPL
