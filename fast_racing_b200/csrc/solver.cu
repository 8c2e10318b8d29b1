// Batched device-side L-BFGS — see solver.cuh for the design.
#include "solver.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

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
//           new gradient (5 independent dot products per history slot, one warp per slot),
//   loops   the two recursions on bound x bound Gram matrices, done by one warp in registers,
//   pass 2  one streaming pass: d = dg*g + sum_l a_l s_l + b_l y_l.
// Same mathematics as the reference's recursion; the summation order differs (stated tolerance).
// Gram matrices per trajectory: SY[i][j] = s_i.y_j, SYT = SY^T, YY[i][j] = y_i.y_j (m x m each).
// Requires m <= 128 and UPD_THREADS == 256.
// ---------------------------------------------------------------------------------------------
__device__ void direction_vector_free(const SolverView& sv, int b, int n, int m, int end, int bound,
                                      const double* __restrict__ g, double* __restrict__ d, const double* __restrict__ S,
                                      const double* __restrict__ Y, double* ysv) {
    __shared__ double sg[128], yg[128], cb[128], ca[128], tv[128], al[128];
    __shared__ double sh_gamma;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double* SY = sv.SY + (size_t)b * m * m;
    double* SYT = sv.SYT + (size_t)b * m * m;
    double* YY = sv.YY + (size_t)b * m * m;
    const double* sn = S + (size_t)end * n;
    const double* yn = Y + (size_t)end * n;
    // ---- pass 1 ----
    for (int l = warp; l < bound; l += UPD_THREADS / 32) {
        const double* sl = S + (size_t)l * n;
        const double* yl = Y + (size_t)l * n;
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
    // ---- loop 1 (newest -> oldest): alpha_j = (s_j . q) / ys_j,  q -= alpha_j y_j ----
    // Blocked by 16 history pairs: the contribution of all EARLIER blocks to the 16 rows of a block is a dense
    // 16 x (#earlier) mat-vec done by the whole CTA (16 threads per row, coalesced row reads, one global round
    // trip per block); only the 16 x 16 diagonal block is sequential, from shared memory, on 16 lanes.
    // Age order: o = 0 is the newest pair, slot(o) = (end - o) mod m.
    __shared__ double tpart[16];
    __shared__ double dblk[16][17];
    const int rr = tid >> 4, cc = tid & 15;  // 16 rows x 16 threads
    for (int l = tid; l < 128; l += UPD_THREADS) cb[l] = 0.0, ca[l] = 0.0;
    __syncthreads();
    for (int ob = 0; ob < bound; ob += 16) {
        const int nb = min(16, bound - ob);
        {
            double t = 0.0;
            if (rr < nb) {
                const int j = (end - (ob + rr) + 2 * m) % m;
                const double* row = SY + (size_t)j * m;
                for (int o2 = cc; o2 < ob; o2 += 16) {
                    const int l = (end - o2 + 2 * m) % m;
                    t += cb[l] * row[l];
                }
                if (cc < nb) dblk[rr][cc] = row[(end - (ob + cc) + 2 * m) % m];
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o, 16);
            if (cc == 0) tpart[rr] = t;
        }
        __syncthreads();
        if (warp == 0) {
            double bl = 0.0;  // lane c < 16 holds the coefficient of pair ob + c once computed
            for (int r2 = 0; r2 < nb; r2++) {
                const int j = (end - (ob + r2) + 2 * m) % m;
                double v = (lane < r2) ? bl * dblk[r2][lane] : 0.0;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 16);
                const double alpha = (v + tpart[r2] - sg[j]) / ysv[j];
                if (lane == r2) {
                    bl = -alpha;
                    al[j] = alpha;
                    cb[j] = -alpha;
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) sh_gamma = ysv[end] / YY[(size_t)end * m + end];  // ys / yy of the newest pair
    __syncthreads();
    const double gamma = sh_gamma;
    for (int l = tid; l < 128; l += UPD_THREADS) cb[l] *= gamma;
    __syncthreads();
    const double dg = -gamma;  // coefficient of g
    // t_j = dg * (y_j . g) + sum_l b_l (y_j . y_l): independent of the second recursion
    if (tid < bound) {
        double t = 0.0;
#pragma unroll 4
        for (int l = 0; l < bound; l++) t += cb[l] * YY[(size_t)l * m + tid];
        tv[tid] = dg * yg[tid] + t;
    }
    __syncthreads();
    // ---- loop 2 (oldest -> newest): beta = (y_j . r) / ys_j,  r += (alpha_j - beta) s_j ----
    // Same blocking; age order now o = 0 is the OLDEST pair: slot(o) = (end - (bound-1) + o) mod m; the row of
    // pair j is SYT[j][l] = s_l . y_j.
    for (int ob = 0; ob < bound; ob += 16) {
        const int nb = min(16, bound - ob);
        {
            double t = 0.0;
            if (rr < nb) {
                const int j = (end - (bound - 1) + (ob + rr) + 2 * m) % m;
                const double* row = SYT + (size_t)j * m;
                for (int o2 = cc; o2 < ob; o2 += 16) {
                    const int l = (end - (bound - 1) + o2 + 2 * m) % m;
                    t += ca[l] * row[l];
                }
                if (cc < nb) dblk[rr][cc] = row[(end - (bound - 1) + (ob + cc) + 2 * m) % m];
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o, 16);
            if (cc == 0) tpart[rr] = t;
        }
        __syncthreads();
        if (warp == 0) {
            double al2 = 0.0;
            for (int r2 = 0; r2 < nb; r2++) {
                const int j = (end - (bound - 1) + (ob + r2) + 2 * m) % m;
                double v = (lane < r2) ? al2 * dblk[r2][lane] : 0.0;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 16);
                const double beta = (tv[j] + tpart[r2] + v) / ysv[j];
                if (lane == r2) {
                    al2 = al[j] - beta;
                    ca[j] = al2;
                }
            }
        }
        __syncthreads();
    }
    // ---- pass 2: d = dg*g + sum_l ca_l s_l + cb_l y_l ----
    for (int base = 0; base < n; base += UPD_THREADS * 4) {
        double acc[4];
        int idx[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            idx[e] = base + e * UPD_THREADS + tid;
            acc[e] = idx[e] < n ? dg * g[idx[e]] : 0.0;
        }
#pragma unroll 8
        for (int l = 0; l < bound; l++) {
            const double a = ca[l], bb = cb[l];
            const double* sl = S + (size_t)l * n;
            const double* yl = Y + (size_t)l * n;
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (idx[e] < n) acc[e] += a * sl[idx[e]] + bb * yl[idx[e]];
        }
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (idx[e] < n) d[idx[e]] = acc[e];
    }
    __syncthreads();
}

// Outcome codes shared through smem
enum : int { ACT_TRIAL = 0, ACT_FINISH = 1, ACT_NEXT_ITER = 2 };

__global__ void __launch_bounds__(UPD_THREADS) lbfgs_update_kernel(SolverView sv) {
    const int b = blockIdx.x;
    TrajSolve* st = sv.st + b;
    if (st->phase == PH_DONE) return;
    __shared__ double red[16];
    __shared__ int sh_act;
    __shared__ double sh_step;
    const int tid = threadIdx.x;
    const int n = sv.x_off[b + 1] - sv.x_off[b];
    const size_t xo = (size_t)sv.x_off[b];
    double* x = sv.x + xo;
    double* g = sv.g + xo;
    double* xp = sv.xp + xo;
    double* gp = sv.gp + xo;
    double* d = sv.d + xo;
    const int m = sv.cfg.mem_size;
    double* S = sv.S + (size_t)m * xo;
    double* Y = sv.Y + (size_t)m * xo;
    double* ysv = sv.ys + (size_t)m * b;
    double* alv = sv.alpha + (size_t)m * b;
    const LbfgsConfig& P = sv.cfg;
    const double f = sv.cost[b];
    const int phase = st->phase;

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
            double xx, gg;
            block_dot2(x, x, g, g, n, red, xx, gg);
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
                double* se = S + (size_t)end * n;
                double* ye = Y + (size_t)end * n;
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
                if (sv.vector_free) {
                    direction_vector_free(sv, b, n, m, end, bound, g, d, S, Y, ysv);
                    if (tid == 0) st->step = 1.0;
                    start_iter = true;
                } else {
                int j = end1;
                for (int it = 0; it < bound; it++) {
                    j = (j + m - 1) % m;
                    const double* sj = S + (size_t)j * n;
                    const double* yj = Y + (size_t)j * n;
                    const double a = block_dot(sj, d, n, red) / ysv[j];
                    if (tid == 0) alv[j] = a;
                    const double c = -a;
                    for (int i = tid; i < n; i += UPD_THREADS) d[i] += c * yj[i];
                }
                {
                    const double sc = ys / yy;
                    for (int i = tid; i < n; i += UPD_THREADS) d[i] *= sc;
                }
                __syncthreads();  // alv[] written by thread 0 above
                for (int it = 0; it < bound; it++) {
                    const double* sj = S + (size_t)j * n;
                    const double* yj = Y + (size_t)j * n;
                    const double beta = block_dot(yj, d, n, red) / ysv[j];
                    const double c = alv[j] - beta;
                    for (int i = tid; i < n; i += UPD_THREADS) d[i] += c * sj[i];
                    j = (j + 1) % m;
                }
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
        }
    }
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

__global__ void gather_fx_kernel(const TrajSolve* st, int B, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) out[i] = st[i].fx;
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
    int* xo = nullptr;
    bool ok = dalloc(s, x_off.size(), &xo) == cudaSuccess &&
              cudaMemcpy(xo, x_off.data(), sizeof(int) * x_off.size(), cudaMemcpyHostToDevice) == cudaSuccess;
    sv.x_off = xo;
    ok = ok && dalloc(s, (size_t)bv.B, &sv.st) == cudaSuccess && dalloc(s, (size_t)bv.B, &sv.active) == cudaSuccess &&
         dalloc(s, 1, &sv.n_active) == cudaSuccess && dalloc(s, nx, &sv.x) == cudaSuccess &&
         dalloc(s, nx, &sv.g) == cudaSuccess && dalloc(s, (size_t)bv.B, &sv.cost) == cudaSuccess &&
         dalloc(s, nx, &sv.xp) == cudaSuccess && dalloc(s, nx, &sv.gp) == cudaSuccess && dalloc(s, nx, &sv.d) == cudaSuccess &&
         dalloc(s, m * nx, &sv.S) == cudaSuccess && dalloc(s, m * nx, &sv.Y) == cudaSuccess &&
         dalloc(s, m * bv.B, &sv.ys) == cudaSuccess && dalloc(s, m * bv.B, &sv.alpha) == cudaSuccess &&
         dalloc(s, m * m * bv.B, &sv.SY) == cudaSuccess && dalloc(s, m * m * bv.B, &sv.SYT) == cudaSuccess &&
         dalloc(s, m * m * bv.B, &sv.YY) == cudaSuccess &&
         dalloc(s, (size_t)bv.B, &s->fx_dev) == cudaSuccess && dalloc(s, 1, &s->am_idx) == cudaSuccess &&
         dalloc(s, 1, &s->am_val) == cudaSuccess;
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
    int ticks = 0, n_active = bv.B;
    while (n_active > 0 && ticks < io.max_ticks) {
        for (int t = 0; t < io.poll_every; t++) {
            if (eval(sv.x, sv.cost, sv.g, io.stream) != cudaSuccess) return FRX_ERR_CUDA;
            if (solver_launch_update(s, io.stream) != cudaSuccess) return FRX_ERR_CUDA;
            ticks++;
        }
        r = solver_poll_active(s, io.stream, &n_active);
        if (r != FRX_OK) return r;
    }
    s->ticks = ticks;
    // final forward pass at the final iterates of ALL trajectories (CPU.hpp:1258-1262)
    solver_bind_active(bv, s, false);
    if (eval(sv.x, sv.cost, sv.g, io.stream) != cudaSuccess) return FRX_ERR_CUDA;
    return FRX_OK;
}
int solver_ticks(SolverState* s) { return s->ticks; }
const double* solver_final_cost_dev(SolverState* s) { return s->fx_dev; }
void solver_bind_active(BatchView& bv, SolverState* s, bool on) { bv.active = on ? s->sv.active : nullptr; }

__global__ void init_active_kernel(int* active, int* n_active, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) active[i] = 1;
    if (i == 0) *n_active = B;
}

int solver_reset(SolverState* s, const SolverIO& io) {
    SolverView& sv = s->sv;
    sv.cfg.delta = io.rel_tol;
    if (cudaMemsetAsync(sv.st, 0, sizeof(TrajSolve) * sv.B, io.stream) != cudaSuccess) return FRX_ERR_CUDA;
    init_active_kernel<<<(sv.B + 255) / 256, 256, 0, io.stream>>>(sv.active, sv.n_active, sv.B);
    if (cudaMemcpyAsync(sv.x, io.x0_host, sizeof(double) * s->total_x, cudaMemcpyHostToDevice, io.stream) != cudaSuccess)
        return FRX_ERR_CUDA;
    if (cudaStreamSynchronize(io.stream) != cudaSuccess) return FRX_ERR_CUDA;  // x0_host may be pageable
    return FRX_OK;
}

cudaError_t solver_launch_update(SolverState* s, cudaStream_t st) {
    lbfgs_update_kernel<<<s->sv.B, UPD_THREADS, 0, st>>>(s->sv);
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
    gather_fx_kernel<<<(sv.B + 255) / 256, 256>>>(sv.st, sv.B, s->fx_dev);
    if (cudaDeviceSynchronize() != cudaSuccess) return FRX_ERR_CUDA;
    std::vector<TrajSolve> h(sv.B);
    if (cudaMemcpy(h.data(), sv.st, sizeof(TrajSolve) * sv.B, cudaMemcpyDeviceToHost) != cudaSuccess) return FRX_ERR_CUDA;
    for (int i = 0; i < sv.B; i++) {
        if (final_cost) final_cost[i] = h[i].fx;
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
