// Batched device-side L-BFGS — see solver.cuh for the design.
#include "solver.cuh"
#include "bulk_copy.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>


#ifdef FRX_EXACT
#define FRX_IMPL frx_exact
#else
#define FRX_IMPL frx_fast
#endif

namespace FRX_IMPL {
using namespace frx;

#define UPD_THREADS 256

#ifdef FRX_EXACT
// Reference operation order: plain sequential sums (lbfgs.hpp:467-475), by one thread, broadcast through smem.
__device__ __forceinline__ double block_dot(const double* __restrict__ a, const double* __restrict__ b, int n, double* red) {
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; i++) s += a[i] * b[i];
        red[0] = s;
    }
    __syncthreads();
    return red[0];
}
__device__ __forceinline__ void block_dot2(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c,
                                           const double* __restrict__ e, int n, double* red, double& r0, double& r1) {
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int i = 0; i < n; i++) s0 += a[i] * b[i];
        for (int i = 0; i < n; i++) s1 += c[i] * e[i];
        red[0] = s0, red[1] = s1;
    }
    __syncthreads();
    r0 = red[0], r1 = red[1];
}
#else
// Deterministic CTA-wide dot product: strided per-thread partials, warp shuffle tree, then the 8
// warp sums added in warp order.  (The reference sums sequentially, lbfgs.hpp:467-475; the
// re-association is part of the stated fp64 tolerance, DESIGN.md "Numerics".)
__device__ __forceinline__ double block_dot(const double* __restrict__ a, const double* __restrict__ b, int n,
                                            double* red) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += UPD_THREADS) s += a[i] * b[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __syncthreads();  // protect red[] from the previous use
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < UPD_THREADS / 32; w++) t += red[w];
    return t;
}
// two dot products in one pass: (a.b, c.e)
__device__ __forceinline__ void block_dot2(const double* __restrict__ a, const double* __restrict__ b,
                                           const double* __restrict__ c, const double* __restrict__ e, int n, double* red,
                                           double& r0, double& r1) {
    double s0 = 0.0, s1 = 0.0;
    for (int i = threadIdx.x; i < n; i += UPD_THREADS) {
        s0 += a[i] * b[i];
        s1 += c[i] * e[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_down_sync(0xffffffffu, s0, o);
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5] = s0;
        red[8 + (threadIdx.x >> 5)] = s1;
    }
    __syncthreads();
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int w = 0; w < UPD_THREADS / 32; w++) {
        t0 += red[w];
        t1 += red[8 + w];
    }
    r0 = t0, r1 = t1;
}
#endif

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// Search direction d = -H g of the two-loop recursion (lbfgs.hpp:1359-1411), evaluated in the
// coefficient space of the basis {s_l, y_l, g} ("vector-free" L-BFGS): the sequential recursion
// (2*bound dependent dot products of length n, each a CTA-wide reduction) becomes
//   pass 1  one streaming pass over the history: the Gram entries that involve the NEW pair and the
//           new gradient (row . (y_new, s_new, g) for every row of S and Y: streamed through a TMA-fed
//           shared-memory ring and contracted on the fp64 tensor cores, see direction_vector_free;
//           plain loads, one warp per history slot, when the ring does not apply),
//   loops   the two recursions on bound x bound Gram matrices (blocked_recursion),
//   pass 2  one streaming pass: d = dg*g + sum_l a_l s_l + b_l y_l (same ring, whole rows).
// Same mathematics as the reference's recursion; the summation order differs (stated tolerance).
// Gram matrices per trajectory: SY[i][j] = s_i.y_j, SYT = SY^T, YY[i][j] = y_i.y_j (m x m each).
// Requires m <= 128 and UPD_THREADS == 256.
// ---------------------------------------------------------------------------------------------
// Phase clocks of the update kernel (debug builds: FRX_NVCC_EXTRA=-DFRX_PHASE_CLK): thread 0 of CTA 0 accumulates the
// clock64() differences between marks over all launches; scripts/phase_clocks_solver.py prints them.
__device__ long long g_upd_clk[16];
#ifdef FRX_PHASE_CLK
#define UPD_MARK(i)                                   \
    do {                                              \
        if (blockIdx.x == 0 && threadIdx.x == 0) {    \
            const long long now_ = clock64();         \
            g_upd_clk[i] += now_ - upd_t0_;           \
            upd_t0_ = now_;                           \
        }                                             \
    } while (0)
#define UPD_MARK_DECL long long upd_t0_ = clock64()
#define UPD_MARK_ARG , long long& upd_t0_
#define UPD_MARK_PASS , upd_t0_
#else
#define UPD_MARK(i) \
    do {            \
    } while (0)
#define UPD_MARK_DECL
#define UPD_MARK_ARG
#define UPD_MARK_PASS
#endif

// One of the two L-BFGS recursions, expressed on a Gram matrix Mx (row = the pair being solved, column = an already solved
// pair) as a unit-triangular solve over the pairs in age order, blocked by 16, RIGHT-LOOKING:
//   tacc[o]  running value of (vector . current direction) for the pair of age o (initialised by the caller),
//   step     val = tacc * (1 / ys);  coefficient = -val (LOOP 1: -alpha) or alpha - val (LOOP 2: alpha - beta),
//   update   every later pair:  tacc[o2] += coefficient * Mx[slot(o2)][slot(o)].
// Per block: the 16 x 16 diagonal part is sequential on 16 lanes (each lane owns one row in registers; one shuffle
// broadcast + one FMA per step), while the panel below it - whose loads do not depend on the new coefficients - is
// fetched BEFORE the sequential part and applied after it by the whole CTA; the next diagonal block is fetched in the
// same shadow.  Deterministic (fixed order).  Requires UPD_THREADS == 256, bound <= 128.
template <int LOOP>
__device__ __forceinline__ void blocked_recursion(const double* __restrict__ Mx, int m, int bound, double* tacc, const double* rys,
                                                  double* al, double* coef_out, double (*dblk)[16][17], double* cbk, const int* slot) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rr = tid >> 4, cc = tid & 15;  // 16 rows x 16 columns
    if (rr < min(16, bound) && cc < min(16, bound)) dblk[0][rr][cc] = Mx[(size_t)slot[rr] * m + slot[cc]];
    __syncthreads();
    int cur = 0;
    for (int ob = 0; ob < bound; ob += 16, cur ^= 1) {
        const int nb = min(16, bound - ob);
        double pv[7];
#pragma unroll
        for (int t = 0; t < 7; t++) {
            const int o2 = ob + 16 + rr + 16 * t;
            pv[t] = (o2 < bound && cc < nb) ? Mx[(size_t)slot[o2] * m + slot[ob + cc]] : 0.0;
        }
        double nd = 0.0;
        {
            const int r2 = ob + 16 + rr, c2 = ob + 16 + cc;
            if (r2 < bound && c2 < bound) nd = Mx[(size_t)slot[r2] * m + slot[c2]];
        }
        if (warp == 0) {
            const int r = lane & 15;  // both half-warps run the same recursion (width-16 shuffles); lanes < 16 publish
            const bool on = r < nb;
            double a = on ? tacc[ob + r] : 0.0;
            double drow[16];
#pragma unroll
            for (int c = 0; c < 16; c++) drow[c] = dblk[cur][r][c];
            const int js = on ? slot[ob + r] : 0;
            const double ry = on ? rys[js] : 0.0;
            const double alr = (LOOP == 2 && on) ? al[js] : 0.0;
            double mycoef = 0.0, myval = 0.0;
#pragma unroll
            for (int c = 0; c < 16; c++) {
                if (c < nb) {  // uniform
                    const double val = a * ry;
                    const double cf = (LOOP == 1) ? -val : (alr - val);
                    const double cfb = __shfl_sync(0xffffffffu, cf, c, 16);
                    if (r == c) mycoef = cfb, myval = val;
                    if (r > c) a += cfb * drow[c];
                }
            }
            if (lane < nb) {
                cbk[lane] = mycoef;
                coef_out[js] = mycoef;
                if (LOOP == 1) al[js] = myval;
            }
        }
        __syncthreads();
        {
            const double ck = cc < nb ? cbk[cc] : 0.0;
#pragma unroll
            for (int t = 0; t < 7; t++) {
                const int o2 = ob + 16 + rr + 16 * t;
                double v = pv[t] * ck;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 16);
                if (cc == 0 && o2 < bound) tacc[o2] += v;
            }
            dblk[cur ^ 1][rr][cc] = nd;
        }
        __syncthreads();
    }
}

// The history of one trajectory streamed through shared memory by the TMA unit, through a ring of UPD_NST stages of `ch`
// doubles, so that the next UPD_NST - 1 chunks are in flight while the CTA works on one.  (The plain-load passes below run at
// about one DRAM latency per unrolled trip: 1.8 TB/s over 256 CTAs, 3.1 TB/s over 1024; scripts/micro/stream_bench.cu is the
// stand-alone measurement behind the ring.)  Rows are padded to an even length ld, so any row segment that starts at an even
// column and has an even length is a legal bulk copy (16-byte aligned, multiple of 16 bytes).
//   pass 1  chunk = tile of 8 rows x W columns of S or Y (8 copies, one per row segment, row stride W + 4 doubles in the
//           stage: conflict-free fragment loads), order: row block, array, column tile;
//   pass 2  chunk = R2 whole rows of S or Y (one copy), order: the reverse of pass 1's rows (what was streamed last is
//           what the L2 still holds).
#define UPD_NST_MAX 8  // ring stages: sv.upd_nst <= UPD_NST_MAX
#define UPD_PART (2 * 8 * 8 * 3)  // doubles: the eight per-warp partial 8 x 3 tiles of two row blocks
#define UPD_COLS 8                // columns per thread in pass 2: n <= UPD_COLS * UPD_THREADS
// Geometry of one trajectory's chunks inside a ring whose stages hold `ch` doubles: tile width W (a multiple of 32, or the
// whole padded row) and column tiles per row for pass 1, rows per chunk for pass 2; ok = the streamed passes apply.  Shared by
// the kernel and the host (frx_debug_update_ring -> tests/test_host_cpu.py checks the invariants the bulk copies rely on).
struct RingGeom {
    int W, CT, R2;
    bool ok;
};
__host__ __device__ inline RingGeom ring_geometry(int n, int ld, int m, int ch) {
    RingGeom g;
    const int wmax = (ch / 8 - 4) & ~31;  // widest tile the stage holds, multiple of 32
    g.W = (ld <= wmax || wmax < 32) ? ld : wmax;
    g.CT = (ld + g.W - 1) / g.W;
    const int r = ch / ld;
    g.R2 = r < 1 ? 1 : (r > m ? m : r);
    g.ok = ld <= ch && 8 * (g.W + 4) <= ch && (m & 7) == 0 && n <= UPD_COLS * UPD_THREADS;
    return g;
}
struct HistStream {
    double* ring;              // nst * ch doubles
    unsigned long long* full;  // nst mbarriers
    const double* S;
    const double* Y;
    int n, ld, m, ch, nst;
    int W, CT;  // pass 1: tile width (multiple of 32, or ld when one tile covers a row), column tiles per row
    int R2;     // pass 2: rows per chunk
};
// pass 1, chunk q = (rb * 2 + arr) * CT + ct
__device__ __forceinline__ void hist_issue_tile(const HistStream& hs, int seq, int q) {
    const int ct = q % hs.CT, ga = q / hs.CT, arr = ga & 1, rb = ga >> 1;
    const int c0 = ct * hs.W, wt = min(hs.W, hs.ld - c0);
    const int stg = seq % hs.nst;
    double* dst = hs.ring + (size_t)stg * hs.ch;
    const double* src = (arr ? hs.Y : hs.S) + (size_t)(rb * 8) * hs.ld + c0;
    mbar_expect_tx(hs.full + stg, 8u * (unsigned)wt * 8u);
#pragma unroll
    for (int r = 0; r < 8; r++) bulk_g2s(dst + (size_t)r * (hs.W + 4), src + (size_t)r * hs.ld, (unsigned)wt * 8u, hs.full + stg);
}
// pass 2, chunk q: array q & 1, rows [R2 (q >> 1), ...)
__device__ __forceinline__ void hist_issue_rows(const HistStream& hs, int seq, int q) {
    const int r0 = (q >> 1) * hs.R2, rows = min(hs.R2, hs.m - r0);
    const unsigned bytes = (unsigned)(rows * hs.ld) * 8u;
    const int stg = seq % hs.nst;
    mbar_expect_tx(hs.full + stg, bytes);
    bulk_g2s(hs.ring + (size_t)stg * hs.ch, ((q & 1) ? hs.Y : hs.S) + (size_t)r0 * hs.ld, bytes, hs.full + stg);
}
// D (8 x 8) += A (8 x 4, row-major fragment: lane holds A[lane / 4][lane % 4]) * B (4 x 8: lane holds B[lane % 4][lane / 4]);
// lane holds D[lane / 4][2 (lane % 4) + {0, 1}].  fp64 tensor-core MMA (DMMA), fp64 accumulation.
__device__ __forceinline__ void dmma_884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ void direction_vector_free(const SolverView& sv, int b, int n, int ld, int m, int end, int bound,
                                      const double* __restrict__ g, double* __restrict__ d, const double* __restrict__ S,
                                      const double* __restrict__ Y, double* ysv, double* dyn UPD_MARK_ARG) {
    __shared__ double sg[128], yg[128], cb[128], ca[128], tv[128], al[128];
    __shared__ double sh_gamma;
    __shared__ unsigned long long full_bar[UPD_NST_MAX], empty_bar[UPD_NST_MAX];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double* SY = sv.SY + (size_t)b * m * m;
    double* SYT = sv.SYT + (size_t)b * m * m;
    double* YY = sv.YY + (size_t)b * m * m;
    const double* sn = S + (size_t)end * ld;
    const double* yn = Y + (size_t)end * ld;
    // streamed passes: a stage holds two rows (pass 2) and an 8 x 32 tile (pass 1) at least; history depth a multiple of 8;
    // UPD_COLS columns per thread in pass 2 at most
    HistStream hs;
    hs.ring = dyn, hs.full = full_bar, hs.S = S, hs.Y = Y, hs.n = n, hs.ld = ld, hs.m = m, hs.ch = sv.upd_ch, hs.nst = sv.upd_nst;
    const int NST = sv.upd_nst;
    const RingGeom geo = ring_geometry(n, ld, m, sv.upd_ch > 0 ? sv.upd_ch : 2);
    hs.W = geo.W, hs.CT = geo.CT, hs.R2 = geo.R2;
    const bool streamed = dyn != nullptr && geo.ok;
    const int Q1 = 2 * ((bound + 7) >> 3) * hs.CT;       // chunks of pass 1
    const int Q2 = 2 * ((bound + hs.R2 - 1) / hs.R2);    // chunks of pass 2
    const int nv = (n + 1) & ~1;
    double* part = dyn + (size_t)NST * sv.upd_ch;  // [2][8 warps][8 rows][3]
    double* vec = part + UPD_PART;                      // y_new, s_new, g: the B operand of pass 1
    if (streamed) {
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < UPD_NST_MAX; i++)
                if (i < NST) mbar_init(full_bar + i, 1), mbar_init(empty_bar + i, UPD_THREADS);
        }
        for (int i = tid; i < n; i += UPD_THREADS) vec[i] = yn[i], vec[nv + i] = sn[i], vec[2 * nv + i] = g[i];
        // rows `end` of S and Y were stored by this CTA a moment ago (generic proxy); the bulk copies below read them through
        // the async proxy
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncthreads();
        if (tid == 0)
            for (int q = 0; q < min(NST, Q1); q++) hist_issue_tile(hs, q, q);
        // ---- pass 1 (streamed): P[row][j] = row . (y_new, s_new, g)[j] for every row of S and Y, on the fp64 tensor cores:
        // A = 8 rows x 4 columns of the tile, B = 4 entries of the three vectors (columns 3..7 of B are zero), so one DMMA
        // retires 8 x 4 x 3 useful products for one shared-memory load per operand, and the sum over a row's columns happens
        // in the accumulators instead of in shuffles.  The eight warps split a tile's columns; their accumulators run over
        // the column tiles of a row block, and the eight partial 8 x 3 tiles are then added in warp order by the threads that
        // store the Gram entries, one chunk later (so a chunk costs one CTA barrier).
        const int gid = lane >> 2, TS = hs.W + 4;
        const double* bsrc = vec + (size_t)min(gid, 2) * nv;
        auto flush = [&](int ga) {  // partial tiles of (row block, array) group ga -> Gram entries
            const int r0 = (ga >> 1) * 8, rows = min(8, bound - r0);
            if (tid < 3 * rows) {
                const int row = tid / 3, col = tid - 3 * row, l = r0 + row;
                const double* pp = part + (size_t)(ga & 1) * (8 * 8 * 3) + row * 3 + col;
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < 8; w++) v += pp[w * 8 * 3];
                if (!(ga & 1)) {  // s_l
                    if (col == 0) SY[(size_t)l * m + end] = v, SYT[(size_t)end * m + l] = v;
                    if (col == 2) sg[l] = v;
                } else {  // y_l
                    if (col == 0) YY[(size_t)l * m + end] = v, YY[(size_t)end * m + l] = v;
                    if (col == 1 && l != end) SY[(size_t)end * m + l] = v, SYT[(size_t)l * m + end] = v;
                    if (col == 2) yg[l] = v;
                }
            }
        };
        // Ring protocol: a stage is released by its consumers through an "empty" mbarrier (one arrival per thread), and
        // lane 0 of warp 0 refills the stage of chunk q - 1 at the top of iteration q — the warps meet at a CTA barrier only
        // where their partial tiles are combined (once per row block and array), not once per chunk.
        double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;  // two accumulator pairs: the DMMAs of a tile alternate between them
        for (int q = 0; q < Q1; q++) {
            const int stg = q % NST;
            const int ct = q % hs.CT, ga = q / hs.CT;
            if (tid == 0 && q >= 1 && q - 1 + NST < Q1) {
                mbar_wait(empty_bar + (q - 1) % NST, ((q - 1) / NST) & 1);
                hist_issue_tile(hs, q - 1 + NST, q - 1 + NST);
            }
            mbar_wait(full_bar + stg, (q / NST) & 1);
            const int col0 = ct * hs.W, wt = min(hs.W, ld - col0);
            const int rows = min(8, bound - (ga >> 1) * 8);
            const int kper = ((((wt + 7) >> 3) + 7) >> 3) << 3, KS2 = kper >> 3;  // columns per warp: a multiple of 8
            const int kb = warp * kper + (lane & 3);
            const double* arow = dyn + (size_t)stg * hs.ch + (size_t)gid * TS;
            const double* brow = bsrc + col0;
            const bool rok = gid < rows, bok = gid < 3;
            const int klim = min(wt, n - col0);  // columns of this tile that exist
#pragma unroll 2
            for (int ks = 0; ks < KS2; ks++) {
                const int ka = kb + 8 * ks, kc = ka + 4;
                const double av = (rok && ka < klim) ? arow[ka] : 0.0;
                const double aw = (rok && kc < klim) ? arow[kc] : 0.0;
                const double bv = (bok && ka < klim) ? brow[ka] : 0.0;
                const double bw = (bok && kc < klim) ? brow[kc] : 0.0;
                dmma_884(c0, c1, av, bv);
                dmma_884(e0, e1, aw, bw);
            }
            mbar_arrive(empty_bar + stg);  // this thread is done with the stage (its own arrival orders its own reads)
            if (ct == hs.CT - 1) {
                double* pw = part + (size_t)(ga & 1) * (8 * 8 * 3) + warp * (8 * 3);
                if (rok) {
                    if ((lane & 3) == 0) pw[gid * 3] = c0 + e0, pw[gid * 3 + 1] = c1 + e1;
                    if ((lane & 3) == 1) pw[gid * 3 + 2] = c0 + e0;
                }
                c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
                __syncthreads();  // the eight partial tiles of the group are complete (the buffer of parity ga & 1 is next written two groups on)
                flush(ga);
            }
        }
        // the first chunks of pass 2 are requested now and arrive under the two recursions (every stage has been released: the
        // waits below return at once, the group barrier above was passed by every thread after its last tile)
        if (tid == 0) {
            for (int q = max(0, Q1 - NST); q < Q1; q++) mbar_wait(empty_bar + q % NST, (q / NST) & 1);
            for (int j = 0; j < min(NST, Q2); j++) hist_issue_rows(hs, Q1 + j, Q2 - 1 - j);
        }
    } else
    // ---- pass 1 ----
    for (int l = warp; l < bound; l += UPD_THREADS / 32) {
        const double* sl = S + (size_t)l * ld;
        const double* yl = Y + (size_t)l * ld;
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
        // 8 elements per lane per trip, all loads issued before use (HBM-latency bound otherwise)
        for (int i0 = lane; i0 < n; i0 += 256) {
            double s[8], y[8], yni[8], sni[8], gi[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int i = i0 + 32 * e;
                const bool in = i < n;
                s[e] = in ? sl[i] : 0.0;
                y[e] = in ? yl[i] : 0.0;
                yni[e] = in ? yn[i] : 0.0;
                sni[e] = in ? sn[i] : 0.0;
                gi[e] = in ? g[i] : 0.0;
            }
#pragma unroll
            for (int e = 0; e < 8; e++) {
                a0 += s[e] * yni[e];
                a1 += y[e] * yni[e];
                a2 += y[e] * sni[e];
                a3 += s[e] * gi[e];
                a4 += y[e] * gi[e];
            }
        }
        a0 = warp_sum(a0), a1 = warp_sum(a1), a2 = warp_sum(a2), a3 = warp_sum(a3), a4 = warp_sum(a4);
        if (lane == 0) {
            SY[(size_t)l * m + end] = a0;
            SYT[(size_t)end * m + l] = a0;
            YY[(size_t)l * m + end] = a1;
            YY[(size_t)end * m + l] = a1;
            if (l != end) {
                SY[(size_t)end * m + l] = a2;
                SYT[(size_t)l * m + end] = a2;
            }
            sg[l] = a3;
            yg[l] = a4;
        }
    }
    __syncthreads();
    UPD_MARK(4);
    // ---- the two recursions on the Gram matrices (see blocked_recursion) ----
    __shared__ double tacc[128], rys[128], cbk[16];
    __shared__ double dblk[2][16][17];
    __shared__ int slot[128];
    for (int l = tid; l < 128; l += UPD_THREADS) {
        cb[l] = 0.0, ca[l] = 0.0;
        if (l < bound) {
            rys[l] = 1.0 / ysv[l];
            slot[l] = (end - l + 2 * m) % m;  // loop 1: age o = 0 is the NEWEST pair
        }
    }
    __syncthreads();
    for (int o = tid; o < bound; o += UPD_THREADS) tacc[o] = -sg[slot[o]];  // d = -g: s_j . d starts at -(s_j . g)
    __syncthreads();
    // loop 1 (newest -> oldest): alpha_j = (s_j . q) / ys_j, q -= alpha_j y_j  ->  cb_j = -alpha_j
    blocked_recursion<1>(SY, m, bound, tacc, rys, al, cb, dblk, cbk, slot);
    UPD_MARK(5);
    if (tid == 0) sh_gamma = ysv[end] / YY[(size_t)end * m + end];  // ys / yy of the newest pair
    __syncthreads();
    const double gamma = sh_gamma;
    for (int l = tid; l < 128; l += UPD_THREADS) cb[l] *= gamma;
    __syncthreads();
    const double dg = -gamma;  // coefficient of g
    // t_j = dg * (y_j . g) + sum_l b_l (y_j . y_l): independent of the second recursion.  Two threads per j, four
    // independent partial sums each (the loads of a trip are in flight together).
    {
        const int j = tid & 127, half = tid >> 7;
        double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
        if (j < bound) {
            const int l0 = half * 64, l1 = min(bound, l0 + 64);
            int l = l0;
            for (; l + 4 <= l1; l += 4) {
                t0 += cb[l] * YY[(size_t)l * m + j];
                t1 += cb[l + 1] * YY[(size_t)(l + 1) * m + j];
                t2 += cb[l + 2] * YY[(size_t)(l + 2) * m + j];
                t3 += cb[l + 3] * YY[(size_t)(l + 3) * m + j];
            }
            for (; l < l1; l++) t0 += cb[l] * YY[(size_t)l * m + j];
        }
        if (half == 1 && j < bound) tv[j] = (t0 + t1) + (t2 + t3);
        __syncthreads();
        if (half == 0 && j < bound) tv[j] = dg * yg[j] + (((t0 + t1) + (t2 + t3)) + tv[j]);
    }
    __syncthreads();
    UPD_MARK(6);
    // loop 2 (oldest -> newest): beta = (y_j . r) / ys_j, r += (alpha_j - beta) s_j  ->  ca_j = alpha_j - beta_j
    for (int o = tid; o < bound; o += UPD_THREADS) slot[o] = (end - (bound - 1) + o + 2 * m) % m;  // age o = 0 is the OLDEST pair
    __syncthreads();
    for (int o = tid; o < bound; o += UPD_THREADS) tacc[o] = tv[slot[o]];
    __syncthreads();
    blocked_recursion<2>(SYT, m, bound, tacc, rys, al, ca, dblk, cbk, slot);
    UPD_MARK(7);
    // ---- pass 2: d = dg*g + sum_l ca_l s_l + cb_l y_l ----
    if (streamed) {
        double acc[UPD_COLS];
#pragma unroll
        for (int e = 0; e < UPD_COLS; e++) acc[e] = (tid + e * UPD_THREADS < n) ? dg * vec[2 * nv + tid + e * UPD_THREADS] : 0.0;
        // (pass 2 has no cross-warp exchange, so its ring could run on the empty mbarriers alone, and did: same speed, but
        // compute-sanitizer's racecheck does not follow mbarrier ordering between reader threads and the refilling copy and
        // reports the refill as a hazard; a CTA barrier per chunk costs nothing measurable here and keeps the tool quiet)
        for (int j = 0; j < Q2; j++) {
            const int seq = Q1 + j, q = Q2 - 1 - j, stg = seq % NST;
            mbar_wait(full_bar + stg, (seq / NST) & 1);
            const int r0 = (q >> 1) * hs.R2, rows = min(hs.R2, bound - r0);
            const double* chunk = dyn + (size_t)stg * hs.ch;
            const double* cf = ((q & 1) ? cb : ca) + r0;
#pragma unroll 2
            for (int r = 0; r < rows; r++) {
                const double co = cf[r];
#pragma unroll
                for (int e = 0; e < UPD_COLS; e++)
                    if (tid + e * UPD_THREADS < n) acc[e] += co * chunk[(size_t)r * ld + tid + e * UPD_THREADS];
            }
            __syncthreads();  // every thread is done with this stage: refill it
            if (tid == 0 && j + NST < Q2) hist_issue_rows(hs, seq + NST, q - NST);
        }
#pragma unroll
        for (int e = 0; e < UPD_COLS; e++)
            if (tid + e * UPD_THREADS < n) d[tid + e * UPD_THREADS] = acc[e];
    } else
    for (int base = 0; base < n; base += UPD_THREADS * 4) {
        double acc[4];
        int idx[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            idx[e] = base + e * UPD_THREADS + tid;
            acc[e] = idx[e] < n ? dg * g[idx[e]] : 0.0;
        }
        // newest-streamed slots first: pass 1 walked the slots upwards, so the tail of the history is what the L2 still holds
#pragma unroll 8
        for (int l = bound - 1; l >= 0; l--) {
            const double a = ca[l], bb = cb[l];
            const double* sl = S + (size_t)l * ld;
            const double* yl = Y + (size_t)l * ld;
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (idx[e] < n) acc[e] += a * sl[idx[e]] + bb * yl[idx[e]];
        }
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (idx[e] < n) d[idx[e]] = acc[e];
    }
    __syncthreads();
    UPD_MARK(8);
}

// The sequential two-loop recursion in the reference's order (lbfgs.hpp:1359-1411) on dv, which holds -g on entry.
__device__ void direction_two_loop(double* dv, int n, int ld, int m, int end1, int bound, const double* S, const double* Y, const double* ysv,
                                   double* alv, double ys, double yy, double* red) {
    const int tid = threadIdx.x;
    int j = end1;
    for (int it = 0; it < bound; it++) {
        j = (j + m - 1) % m;
        const double* sj = S + (size_t)j * ld;
        const double* yj = Y + (size_t)j * ld;
        const double a = block_dot(sj, dv, n, red) / ysv[j];
        if (tid == 0) alv[j] = a;
        const double c = -a;
        for (int i = tid; i < n; i += UPD_THREADS) dv[i] += c * yj[i];
    }
    {
        const double sc = ys / yy;
        for (int i = tid; i < n; i += UPD_THREADS) dv[i] *= sc;
    }
    __syncthreads();  // alv[] written by thread 0 above
    for (int it = 0; it < bound; it++) {
        const double* sj = S + (size_t)j * ld;
        const double* yj = Y + (size_t)j * ld;
        const double beta = block_dot(yj, dv, n, red) / ysv[j];
        const double c = alv[j] - beta;
        for (int i = tid; i < n; i += UPD_THREADS) dv[i] += c * sj[i];
        j = (j + 1) % m;
    }
}

// Outcome codes shared through smem
enum : int { ACT_TRIAL = 0, ACT_FINISH = 1, ACT_NEXT_ITER = 2 };

__global__ void __launch_bounds__(UPD_THREADS) lbfgs_update_kernel(SolverView sv) {
    const int b = sv.slot_list[blockIdx.x];
    if (b < 0) return;  // empty slot
    TrajSolve* st = sv.st + b;
    if (st->phase == PH_DONE) return;
    __shared__ double red[16];
    __shared__ int sh_act;
    __shared__ double sh_step;
    extern __shared__ __align__(128) double upd_dyn[];  // ring + vectors of the streamed direction update (0 bytes: plain loads)
    const int tid = threadIdx.x;
    const int n = sv.x_off[b + 1] - sv.x_off[b];
    const size_t xo = (size_t)sv.x_off[b];
    double* x = sv.x + xo;
    double* g = sv.g + xo;
    double* xp = sv.xp + xo;
    double* gp = sv.gp + xo;
    double* d = sv.d + xo;
    const int m = sv.cfg.mem_size;
    // history rows are padded to an even number of doubles (16-byte aligned rows: bulk copies of row segments)
    const int ld = (n + 1) & ~1;
    double* S = sv.S + (size_t)m * (xo + b);
    double* Y = sv.Y + (size_t)m * (xo + b);
    double* ysv = sv.ys + (size_t)m * b;
    double* alv = sv.alpha + (size_t)m * b;
    const LbfgsConfig& P = sv.cfg;
    const double f = sv.cost[b];
    const int phase = st->phase;
    UPD_MARK_DECL;

    bool start_iter = false;  // uniform across the CTA

    if (phase == PH_INIT) {
        // lbfgs.hpp:1211-1247
        for (int i = tid; i < n; i += UPD_THREADS) d[i] = -g[i];
        double xx, gg;
        block_dot2(x, x, g, g, n, red, xx, gg);
        if (tid == 0) {
            st->n_evals = 1;
            st->fx = f;
            if (P.past > 0) st->pf[0] = f;
            double xnorm = sqrt(xx);
            const double gnorm = sqrt(gg);
            if (xnorm < 1.0) xnorm = 1.0;
            if (gnorm / xnorm <= P.g_epsilon) {
                st->status = LBFGS_ALREADY_MINIMIZED;
                sh_act = ACT_FINISH;
            } else {
                st->step = 1.0 / sqrt(gg);  // vec2norminv(d): d = -g, same products
                st->k = 1;
                st->end = 0;
                sh_act = ACT_NEXT_ITER;
            }
        }
        __syncthreads();
        if (sh_act == ACT_NEXT_ITER) start_iter = true;
    } else {
        const double dg = block_dot(g, d, n, red);
        UPD_MARK(0);
        if (tid == 0) {
            st->n_evals++;
            int oc = (phase == PH_MT) ? st->ls.mt_feed(f, dg, P) : st->ls.bt_feed(f, dg, P);
            st->fx = f;
            if (oc == LS_FAILED && phase == PH_MT) {
                // lbfgs.hpp:1271-1282: restore and fall back to backtracking from xp
                st->step = st->stepp;
                st->fx = st->fp;
                const double dg0 = st->ls.dginit;  // = dot(gp, d), unchanged
                oc = st->ls.bt_begin(st->step, st->fx, dg0, P.min_step, P.max_step, P);
                st->phase = PH_BT;
            }
            if (oc == LS_CONTINUE) {
                sh_act = ACT_TRIAL;
                sh_step = st->ls.step;
            } else if (oc == LS_FAILED) {
                st->status = st->ls.code;  // lbfgs.hpp:1284-1292
                sh_act = ACT_FINISH;
                sh_step = -1.0;  // marks "revert to xp"
            } else {
                sh_act = ACT_NEXT_ITER;  // line search succeeded
                sh_step = 0.0;
            }
        }
        __syncthreads();
        if (sh_act == ACT_FINISH) {
            for (int i = tid; i < n; i += UPD_THREADS) {
                x[i] = xp[i];
                g[i] = gp[i];
            }
        } else if (sh_act == ACT_NEXT_ITER) {
            // ---- lbfgs.hpp:1294-1416 ----
            UPD_MARK(1);
            double xx, gg;
            block_dot2(x, x, g, g, n, red, xx, gg);
            UPD_MARK(2);
            if (tid == 0) {
                double xnorm = sqrt(xx);
                const double gnorm = sqrt(gg);
                if (xnorm < 1.0) xnorm = 1.0;
                const int k = st->k;
                int stop = 0;
                if (gnorm / xnorm <= P.g_epsilon) {
                    st->status = LBFGS_CONVERGENCE;
                    stop = 1;
                }
                if (!stop && P.past > 0) {
                    if (P.past <= k) {
                        const double rate = (st->pf[k % P.past] - st->fx) / st->fx;
                        if (fabs(rate) < P.delta) {
                            st->status = LBFGS_STOP;
                            stop = 1;
                        }
                    }
                    if (!stop) st->pf[k % P.past] = st->fx;
                }
                if (!stop && P.max_iterations != 0 && P.max_iterations < k + 1) {
                    st->status = LBFGSERR_MAXIMUMITERATION;
                    stop = 1;
                }
                sh_act = stop ? ACT_FINISH : ACT_NEXT_ITER;
                sh_step = 0.0;
            }
            __syncthreads();
            if (sh_act == ACT_NEXT_ITER) {
                const int end = st->end;
                const int k = st->k;
                double* se = S + (size_t)end * ld;
                double* ye = Y + (size_t)end * ld;
                for (int i = tid; i < n; i += UPD_THREADS) {
                    se[i] = x[i] - xp[i];
                    ye[i] = g[i] - gp[i];
                    d[i] = -g[i];
                }
                double ys, yy;
                block_dot2(ye, se, ye, ye, n, red, ys, yy);
                const int bound = (m <= k) ? m : k;
                const int end1 = (end + 1) % m;
                if (tid == 0) {
                    ysv[end] = ys;
                    st->k = k + 1;
                    st->end = end1;
                }
                __syncthreads();
                UPD_MARK(3);
                // measurement only: the direction update streams the history twice, (4 bound + 4) n doubles
                if (sv.count_bytes && tid == 0) atomicAdd(sv.hist_bytes, 8ull * (4ull * bound + 4ull) * (unsigned long long)n);
                if (sv.vector_free) {
                    direction_vector_free(sv, b, n, ld, m, end, bound, g, d, S, Y, ysv, sv.upd_smem ? upd_dyn : nullptr UPD_MARK_PASS);
                    if (sv.check_dir) {  // self-check (FRX_LBFGS_CHECK=1): the two-loop recursion on a scratch vector
                        double* dc = sv.dchk + xo;
                        for (int i = tid; i < n; i += UPD_THREADS) dc[i] = -g[i];
                        __syncthreads();
                        direction_two_loop(dc, n, ld, m, end1, bound, S, Y, ysv, alv, ys, yy, red);
                        double emax = 0.0, dmax = 0.0;
                        for (int i = tid; i < n; i += UPD_THREADS) emax = fmax(emax, fabs(d[i] - dc[i])), dmax = fmax(dmax, fabs(dc[i]));
                        __shared__ double chk[2][UPD_THREADS / 32];
                        for (int o = 16; o > 0; o >>= 1)
                            emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o)), dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
                        if ((tid & 31) == 0) chk[0][tid >> 5] = emax, chk[1][tid >> 5] = dmax;
                        __syncthreads();
                        if (tid == 0) {
                            for (int w = 1; w < UPD_THREADS / 32; w++) emax = fmax(emax, chk[0][w]), dmax = fmax(dmax, chk[1][w]);
                            const double rel = dmax > 0.0 ? emax / dmax : 0.0;
                            atomicMax(sv.dir_err_bits, (unsigned long long)__double_as_longlong(rel));  // rel >= 0: bit order = value order
                        }
                        __syncthreads();
                    }
                    if (tid == 0) st->step = 1.0;
                    start_iter = true;
                } else {
                direction_two_loop(d, n, ld, m, end1, bound, S, Y, ysv, alv, ys, yy, red);
                if (tid == 0) st->step = 1.0;
                start_iter = true;
                }
            }
        }
    }

    if (start_iter) {
        // lbfgs.hpp:1248-1268: save the iterate and enter the More-Thuente search
        for (int i = tid; i < n; i += UPD_THREADS) {
            xp[i] = x[i];
            gp[i] = g[i];
        }
        const double dg0 = block_dot(g, d, n, red);
        UPD_MARK(9);
        if (tid == 0) {
            st->stepp = st->step;
            st->fp = st->fx;
            int oc = st->ls.mt_begin(st->step, st->fx, dg0, P.min_step, P.max_step, P);
            st->phase = PH_MT;
            if (oc == LS_FAILED) {
                st->step = st->stepp;
                oc = st->ls.bt_begin(st->step, st->fx, dg0, P.min_step, P.max_step, P);
                st->phase = PH_BT;
            }
            if (oc == LS_FAILED) {
                st->status = st->ls.code;
                sh_act = ACT_FINISH;  // x == xp already
            } else {
                sh_act = ACT_TRIAL;
                sh_step = st->ls.step;
            }
        }
        __syncthreads();
    }

    if (sh_act == ACT_TRIAL) {
        const double stp = sh_step;
        for (int i = tid; i < n; i += UPD_THREADS) {  // x = xp + stp * d  (veccpy + vecadd, lbfgs.hpp:825-826)
            double v = xp[i];
            v += stp * d[i];
            x[i] = v;
        }
    } else if (sh_act == ACT_FINISH) {
        if (tid == 0) {
            st->phase = PH_DONE;
            sv.active[b] = 0;
            atomicSub(sv.n_active, 1);
            // hand the slot to the next pending trajectory (its state is PH_INIT, its x the start point): evaluated next tick
            const int nx = atomicAdd(sv.next_pending, 1);
            if (nx < sv.B) {
                sv.active[nx] = 1;
                sv.slot_list[blockIdx.x] = nx;
            } else
                sv.slot_list[blockIdx.x] = -1;
        }
    }
    UPD_MARK(10);
}
cudaError_t fetch_update_clocks(long long* out16, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out16, g_upd_clk, sizeof(long long) * 16);
    if (e == cudaSuccess && reset) {
        long long z[16] = {0};
        e = cudaMemcpyToSymbol(g_upd_clk, z, sizeof(z));
    }
    return e;
}

// ---------------------------------------------------------------------------------------------
// arg-min of cost[0..B) (ties -> lowest index), single CTA
__global__ void argmin_kernel(const double* __restrict__ cost, int B, int* out_idx, double* out_val) {
    __shared__ double sv[256];
    __shared__ int si[256];
    double best = INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const double c = cost[i];
        if (c < best || (c == best && i < bi)) best = c, bi = i;
    }
    sv[threadIdx.x] = best;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double c = sv[threadIdx.x + o];
            const int i = si[threadIdx.x + o];
            if (c < sv[threadIdx.x] || (c == sv[threadIdx.x] && i < si[threadIdx.x])) sv[threadIdx.x] = c, si[threadIdx.x] = i;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *out_idx = si[0];
        *out_val = sv[0];
    }
}

// Penalised objective AT the returned iterate: solver_run's last forward pass evaluates every trajectory at its final x
// into sv.cost.  (TrajSolve::fx is lbfgs.hpp's *ptr_fx: after a failed line search or a tick cut-off it still holds the
// rejected trial's value while x was reverted to xp — the reference discards it, CPU.hpp:1249; ranking candidates on it
// would compare costs that do not belong to the trajectories returned.)
__global__ void gather_fx_kernel(const double* __restrict__ cost, int B, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) out[i] = cost[i];
}

// ---------------------------------------------------------------------------------------------
struct SolverState {
    SolverView sv{};
    std::vector<void*> allocs;
    std::vector<int> x_off;
    int total_x = 0;
    int ticks = 0;
    double* fx_dev = nullptr;
    int* am_idx = nullptr;
    double* am_val = nullptr;
    int* pin_int = nullptr;
    double* pin_dbl = nullptr;
    std::vector<cudaEvent_t> ev;  // pairs around the sampled update launches
    int n_samples = 0;
    double upd_ms = 0.0, upd_bytes = 0.0;
};

template <class T>
static cudaError_t dalloc(SolverState* s, size_t n, T** out) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, sizeof(T) * (n ? n : 1));
    if (e != cudaSuccess) return e;
    s->allocs.push_back(p);
    *out = (T*)p;
    return cudaMemset(p, 0, sizeof(T) * (n ? n : 1));
}

// Ring of the streamed direction update for a batch whose largest trajectory has max_n variables: stages, doubles per stage,
// dynamic shared memory (0: plain loads).  Few trajectories (at most two per SM anyway): one CTA per SM with a deep ring streams
// fastest.  Many: two CTAs per SM, so that one streams while the other runs its recursions (profiles/r02_update_kernel.txt has
// the sweep).  budget_kb / nst > 0 override the first choice (experiments).
void update_ring_config(int max_n, int B, int budget_kb, int nst, int* out_nst, int* out_ch, int* out_smem) {
    *out_nst = 0, *out_ch = 0, *out_smem = 0;
    if (max_n < 1 || max_n > UPD_COLS * UPD_THREADS) return;
    const int ldm = (max_n + 1) & ~1, fixed = UPD_PART + 3 * ldm;
    struct Cfg {
        int budget_kb, nst;
    };
    const bool few = B <= 2 * 148;
    Cfg cfgs[3] = {few ? Cfg{200, 4} : Cfg{96, 3}, Cfg{200, 4}, Cfg{200, 2}};
    if (budget_kb > 0) cfgs[0].budget_kb = budget_kb;
    if (nst > 0) cfgs[0].nst = nst;
    for (const Cfg& c : cfgs) {
        const int ch = ((c.budget_kb * 1024 / 8 - fixed) / c.nst) & ~1;
        if (ch >= ldm && ch >= 8 * (32 + 4)) {
            *out_ch = std::min(ch, std::max(8 * (((ldm + 31) & ~31) + 4), 2 * ldm));  // no more than one row block x whole rows
            *out_nst = c.nst;
            *out_smem = (int)(sizeof(double) * ((size_t)c.nst * *out_ch + fixed));
            return;
        }
    }
}
// Host-side view of the ring for tests: out[8] = {stages, doubles per stage, dynamic smem bytes, W, CT, R2, streamed, fixed doubles}
// for a trajectory of n variables in a batch of `batch` whose largest has max_n.
void debug_update_ring(int max_n, int n, int batch, int* out) {
    int nst = 0, ch = 0, smem = 0;
    update_ring_config(max_n, batch, 0, 0, &nst, &ch, &smem);
    const int ld = (n + 1) & ~1;
    const RingGeom g = ring_geometry(n, ld, 128, ch > 0 ? ch : 2);
    out[0] = nst, out[1] = ch, out[2] = smem, out[3] = g.W, out[4] = g.CT, out[5] = g.R2, out[6] = (smem > 0 && g.ok) ? 1 : 0;
    out[7] = UPD_PART + 3 * ((max_n + 1) & ~1);
}

int solver_create(const BatchView& bv, const std::vector<int>& x_off, SolverState** out) {
    SolverState* s = new SolverState();
    s->x_off = x_off;
    s->total_x = bv.total_x;
    SolverView& sv = s->sv;
    sv.B = bv.B;
    sv.cfg = LbfgsConfig();
    sv.cfg.mem_size = 128;  // CPU.hpp:1243-1247
    sv.cfg.past = 3;
    sv.cfg.g_epsilon = 1.0e-16;
    sv.cfg.min_step = 1.0e-32;
    {
        const char* e = getenv("FRX_LBFGS_TWO_LOOP");  // A/B switch: sequential two-loop in the reference's order
        sv.vector_free = (e && e[0] == '1') ? 0 : 1;
#ifdef FRX_EXACT
        sv.vector_free = 0;  // reference order: the sequential two-loop recursion of lbfgs.hpp:1359-1411
#endif
    }
    const size_t nx = (size_t)bv.total_x;
    const size_t m = (size_t)sv.cfg.mem_size;
    sv.upd_smem = 0, sv.upd_ch = 0, sv.upd_nst = 0;
#ifndef FRX_EXACT
    {
        // Streamed direction update: ring of UPD_NST stages + partial tiles + (y_new, s_new, g).  Two CTAs per SM while the
        // largest trajectory allows it (one streams while the other runs its recursions), one CTA per SM with a deeper ring
        // beyond; plain loads when even that does not hold two rows per stage.  FRX_LBFGS_STREAM=0 forces plain loads.
        int max_n = 0;
        for (size_t i = 0; i + 1 < x_off.size(); i++) max_n = std::max(max_n, x_off[i + 1] - x_off[i]);
        const char* e = getenv("FRX_LBFGS_STREAM");
        if (sv.vector_free && !(e && e[0] == '0') && max_n <= UPD_COLS * UPD_THREADS) {
            // Few trajectories (at most two per SM anyway): one CTA per SM with a deep ring streams fastest.  Many: two CTAs per
            // SM, so that one streams while the other runs its recursions.  (profiles/r02_update_kernel.txt has the sweep.)
            int budget_kb = 0, nst = 0;
            if (const char* eb = getenv("FRX_UPD_BUDGET_KB")) budget_kb = atoi(eb);  // experiments
            if (const char* en = getenv("FRX_UPD_NST")) nst = std::max(2, std::min(UPD_NST_MAX, atoi(en)));
            update_ring_config(max_n, bv.B, budget_kb, nst, &sv.upd_nst, &sv.upd_ch, &sv.upd_smem);
            if (sv.upd_smem) {
                // the attribute is per function and per device, shared by every solver of the process: only ever raise it
                // (a second, smaller batch must not lower the limit under a larger one that is still alive)
                static std::mutex mu;
                static int cur[64];
                int dev = 0;
                cudaGetDevice(&dev);
                std::lock_guard<std::mutex> lock(mu);
                if (dev >= 0 && dev < 64 && sv.upd_smem > cur[dev]) {
                    if (cudaFuncSetAttribute(lbfgs_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sv.upd_smem) != cudaSuccess) {
                        delete s;
                        return FRX_ERR_CUDA;
                    }
                    cur[dev] = sv.upd_smem;
                }
            }
        }
    }
#endif
    int* xo = nullptr;
    bool ok = dalloc(s, x_off.size(), &xo) == cudaSuccess &&
              cudaMemcpy(xo, x_off.data(), sizeof(int) * x_off.size(), cudaMemcpyHostToDevice) == cudaSuccess;
    sv.x_off = xo;
    ok = ok && dalloc(s, (size_t)bv.B, &sv.st) == cudaSuccess && dalloc(s, (size_t)bv.B, &sv.active) == cudaSuccess &&
         dalloc(s, 1, &sv.n_active) == cudaSuccess && dalloc(s, (size_t)bv.B, &sv.slot_list) == cudaSuccess &&
         dalloc(s, 1, &sv.next_pending) == cudaSuccess && dalloc(s, nx, &sv.x) == cudaSuccess &&
         dalloc(s, nx, &sv.g) == cudaSuccess && dalloc(s, (size_t)bv.B, &sv.cost) == cudaSuccess &&
         dalloc(s, nx, &sv.xp) == cudaSuccess && dalloc(s, nx, &sv.gp) == cudaSuccess && dalloc(s, nx, &sv.d) == cudaSuccess &&
         dalloc(s, m * (nx + bv.B), &sv.S) == cudaSuccess && dalloc(s, m * (nx + bv.B), &sv.Y) == cudaSuccess &&
         dalloc(s, m * bv.B, &sv.ys) == cudaSuccess && dalloc(s, m * bv.B, &sv.alpha) == cudaSuccess &&
         dalloc(s, m * m * bv.B, &sv.SY) == cudaSuccess && dalloc(s, m * m * bv.B, &sv.SYT) == cudaSuccess &&
         dalloc(s, m * m * bv.B, &sv.YY) == cudaSuccess &&
         dalloc(s, (size_t)bv.B, &s->fx_dev) == cudaSuccess && dalloc(s, 1, &s->am_idx) == cudaSuccess &&
         dalloc(s, 1, &s->am_val) == cudaSuccess && dalloc(s, 1, &sv.hist_bytes) == cudaSuccess &&
         dalloc(s, 1, &sv.dir_err_bits) == cudaSuccess;
    {
        const char* e = getenv("FRX_LBFGS_CHECK");  // self-check of the vector-free direction against the two-loop recursion
        sv.check_dir = (e && e[0] == '1' && sv.vector_free) ? 1 : 0;
        sv.dchk = nullptr;
        if (sv.check_dir) ok = ok && dalloc(s, nx, &sv.dchk) == cudaSuccess;
    }
    ok = ok && cudaMallocHost(&s->pin_int, sizeof(int) * 4) == cudaSuccess &&
         cudaMallocHost(&s->pin_dbl, sizeof(double) * 4) == cudaSuccess;
    if (!ok) {
        std::fprintf(stderr, "[frx] solver_create: %s\n", cudaGetErrorString(cudaGetLastError()));
        solver_destroy(s);
        return FRX_ERR_CUDA;
    }
    *out = s;
    return FRX_OK;
}

void solver_destroy(SolverState* s) {
    if (!s) return;
    for (void* p : s->allocs) cudaFree(p);
    for (cudaEvent_t e : s->ev) cudaEventDestroy(e);
    if (s->pin_int) cudaFreeHost(s->pin_int);
    if (s->pin_dbl) cudaFreeHost(s->pin_dbl);
    delete s;
}

int solver_reset(SolverState* s, const SolverIO& io);
cudaError_t solver_launch_update(SolverState* s, cudaStream_t st);
int solver_poll_active(SolverState* s, cudaStream_t st, int* n_active);
void solver_bind_active(BatchView& bv, SolverState* s, bool on);

int solver_run(SolverState* s, BatchView& bv, const SolverIO& io, const EvalFn& eval) {
    int r = solver_reset(s, io);
    if (r != FRX_OK) return r;
    SolverView& sv = s->sv;
    solver_bind_active(bv, s, true);
    struct Unbind {  // bv.active must be nullptr again on EVERY exit: a stale mask would make later evaluations skip trajectories
        BatchView& bv;
        ~Unbind() { bv.active = nullptr, bv.slot_list = nullptr, bv.n_slots = 0; }
    } unbind{bv};
    int ticks = 0, n_active = bv.B;
    // Measurement: the first update launch of every poll group is bracketed by events and counts the history bytes it
    // streams (-> achieved HBM bandwidth of the direction update, bench.py "solves.update_kernel").
    s->n_samples = 0;
    if (cudaMemsetAsync(sv.hist_bytes, 0, sizeof(unsigned long long), io.stream) != cudaSuccess) return FRX_ERR_CUDA;
    if (cudaMemsetAsync(sv.dir_err_bits, 0, sizeof(unsigned long long), io.stream) != cudaSuccess) return FRX_ERR_CUDA;
    // A tick is four launches of a few microseconds to a few hundred each; a poll group (poll_every ticks = 4 * poll_every
    // launches with identical arguments) is captured ONCE into a CUDA graph and replayed, so the host pays one launch per group
    // and the device sees no inter-kernel launch gaps — what matters when few trajectories are still alive (or B = 1, the
    // reference's own call).  Every GRAPH_SAMPLE_EVERY-th group runs un-captured with the first update launch bracketed by
    // events and counting its history bytes (-> bench.py "solves.update_kernel").  FRX_SOLVER_GRAPH=0 disables the capture.
    constexpr int GRAPH_SAMPLE_EVERY = 8;
    bool use_graph = true;
    if (const char* e = getenv("FRX_SOLVER_GRAPH")) use_graph = !(e[0] == '0');
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    struct GraphGuard {
        cudaGraph_t& g;
        cudaGraphExec_t& e;
        ~GraphGuard() {
            if (e) cudaGraphExecDestroy(e);
            if (g) cudaGraphDestroy(g);
        }
    } gguard{graph, gexec};
    int group = 0;
    while (n_active > 0 && ticks < io.max_ticks) {
        const bool sampled_group = !use_graph || (group % GRAPH_SAMPLE_EVERY == 0);
        if (sampled_group) {
            for (int t = 0; t < io.poll_every; t++) {
                if (eval(sv.x, sv.cost, sv.g, io.stream) != cudaSuccess) return FRX_ERR_CUDA;
                const bool sample = (t == 0);
                if (sample) {
                    if ((int)s->ev.size() < 2 * (s->n_samples + 1)) {
                        cudaEvent_t a, b2;
                        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b2) != cudaSuccess) return FRX_ERR_CUDA;
                        s->ev.push_back(a), s->ev.push_back(b2);
                    }
                    cudaEventRecord(s->ev[2 * s->n_samples], io.stream);
                }
                sv.count_bytes = sample ? 1 : 0;
                if (solver_launch_update(s, io.stream) != cudaSuccess) return FRX_ERR_CUDA;
                if (sample) cudaEventRecord(s->ev[2 * s->n_samples + 1], io.stream), s->n_samples++;
                ticks++;
            }
        } else {
            if (!gexec) {
                sv.count_bytes = 0;
                if (cudaStreamBeginCapture(io.stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return FRX_ERR_CUDA;
                bool ok = true;
                for (int t = 0; t < io.poll_every && ok; t++)
                    ok = eval(sv.x, sv.cost, sv.g, io.stream) == cudaSuccess && solver_launch_update(s, io.stream) == cudaSuccess;
                const cudaError_t ce = cudaStreamEndCapture(io.stream, &graph);
                if (!ok || ce != cudaSuccess || cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) return FRX_ERR_CUDA;
            }
            if (cudaGraphLaunch(gexec, io.stream) != cudaSuccess) return FRX_ERR_CUDA;
            ticks += io.poll_every;
        }
        group++;
        r = solver_poll_active(s, io.stream, &n_active);
        if (r != FRX_OK) return r;
    }
    s->ticks = ticks;
    {
        unsigned long long hb = 0;
        if (cudaMemcpyAsync(&hb, sv.hist_bytes, sizeof(hb), cudaMemcpyDeviceToHost, io.stream) != cudaSuccess) return FRX_ERR_CUDA;
        if (cudaStreamSynchronize(io.stream) != cudaSuccess) return FRX_ERR_CUDA;
        s->upd_bytes = (double)hb, s->upd_ms = 0.0;
        for (int i = 0; i < s->n_samples; i++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, s->ev[2 * i], s->ev[2 * i + 1]);
            s->upd_ms += ms;
        }
    }
    // final forward pass at the final iterates of ALL trajectories (CPU.hpp:1258-1262)
    solver_bind_active(bv, s, false);
    if (eval(sv.x, sv.cost, sv.g, io.stream) != cudaSuccess) return FRX_ERR_CUDA;
    return FRX_OK;
}
int solver_ticks(SolverState* s) { return s->ticks; }
int solver_direction_check(SolverState* s, double* max_rel) {
    unsigned long long bits = 0;
    if (cudaMemcpy(&bits, s->sv.dir_err_bits, sizeof(bits), cudaMemcpyDeviceToHost) != cudaSuccess) return FRX_ERR_CUDA;
    double v;
    memcpy(&v, &bits, sizeof(v));
    *max_rel = s->sv.check_dir ? v : -1.0;
    return FRX_OK;
}
int solver_update_stats(SolverState* s, double* ms, double* bytes, int* samples) {
    *ms = s->upd_ms, *bytes = s->upd_bytes, *samples = s->n_samples;
    return FRX_OK;
}
const double* solver_final_cost_dev(SolverState* s) { return s->fx_dev; }
void solver_bind_active(BatchView& bv, SolverState* s, bool on) {
    bv.active = on ? s->sv.active : nullptr;
    bv.slot_list = on ? s->sv.slot_list : nullptr;
    bv.n_slots = on ? s->sv.n_slots : 0;
}

__global__ void init_active_kernel(int* active, int* n_active, int* slot_list, int* next_pending, int B, int n_slots) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) active[i] = i < n_slots;  // the first n_slots trajectories start, the others wait for a free slot
    if (i < n_slots) slot_list[i] = i;
    if (i == 0) *n_active = B, *next_pending = n_slots;
}

int solver_reset(SolverState* s, const SolverIO& io) {
    SolverView& sv = s->sv;
    sv.cfg.delta = io.rel_tol;
    if (cudaMemsetAsync(sv.st, 0, sizeof(TrajSolve) * sv.B, io.stream) != cudaSuccess) return FRX_ERR_CUDA;
    sv.n_slots = (io.slots > 0 && io.slots < sv.B) ? io.slots : sv.B;
    init_active_kernel<<<(sv.B + 255) / 256, 256, 0, io.stream>>>(sv.active, sv.n_active, sv.slot_list, sv.next_pending, sv.B, sv.n_slots);
    if (cudaMemcpyAsync(sv.x, io.x0_host, sizeof(double) * s->total_x, cudaMemcpyHostToDevice, io.stream) != cudaSuccess)
        return FRX_ERR_CUDA;
    if (cudaStreamSynchronize(io.stream) != cudaSuccess) return FRX_ERR_CUDA;  // x0_host may be pageable
    return FRX_OK;
}

cudaError_t solver_launch_update(SolverState* s, cudaStream_t st) {
    lbfgs_update_kernel<<<s->sv.n_slots, UPD_THREADS, (size_t)s->sv.upd_smem, st>>>(s->sv);
    return cudaGetLastError();
}

int solver_poll_active(SolverState* s, cudaStream_t st, int* n_active) {
    if (cudaMemcpyAsync(s->pin_int, s->sv.n_active, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) return FRX_ERR_CUDA;
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        std::fprintf(stderr, "[frx] solver: %s\n", cudaGetErrorString(e));
        return FRX_ERR_CUDA;
    }
    *n_active = s->pin_int[0];
    return FRX_OK;
}

int solver_fetch(SolverState* s, double* final_cost, int* status, int* iters, int* evals, double* x_final) {
    SolverView& sv = s->sv;
    gather_fx_kernel<<<(sv.B + 255) / 256, 256>>>(sv.cost, sv.B, s->fx_dev);
    if (cudaDeviceSynchronize() != cudaSuccess) return FRX_ERR_CUDA;
    std::vector<TrajSolve> h(sv.B);
    if (cudaMemcpy(h.data(), sv.st, sizeof(TrajSolve) * sv.B, cudaMemcpyDeviceToHost) != cudaSuccess) return FRX_ERR_CUDA;
    if (final_cost && cudaMemcpy(final_cost, s->fx_dev, sizeof(double) * sv.B, cudaMemcpyDeviceToHost) != cudaSuccess) return FRX_ERR_CUDA;
    for (int i = 0; i < sv.B; i++) {
        if (status) status[i] = (h[i].phase == PH_DONE) ? h[i].status : LBFGSERR_MAXIMUMITERATION;
        if (iters) iters[i] = h[i].k;
        if (evals) evals[i] = h[i].n_evals;
    }
    if (x_final && cudaMemcpy(x_final, sv.x, sizeof(double) * s->total_x, cudaMemcpyDeviceToHost) != cudaSuccess)
        return FRX_ERR_CUDA;
    return FRX_OK;
}

int solver_argmin(SolverState* s, const double* cost_dev, int B, cudaStream_t st, int* idx, double* val) {
    argmin_kernel<<<1, 256, 0, st>>>(cost_dev, B, s->am_idx, s->am_val);
    if (cudaMemcpyAsync(s->pin_int, s->am_idx, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) return FRX_ERR_CUDA;
    if (cudaMemcpyAsync(s->pin_dbl, s->am_val, sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) return FRX_ERR_CUDA;
    if (cudaStreamSynchronize(st) != cudaSuccess) return FRX_ERR_CUDA;
    *idx = s->pin_int[0];
    *val = s->pin_dbl[0];
    return FRX_OK;
}

}  // namespace FRX_IMPL
