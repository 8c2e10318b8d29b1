// backwardP for ONE waypoint (se3gcopter_cpu.hpp:777-813 with objectiveNLS :749-774): the xi that reproduces a target point
// inside a V-polytope, min_xi | V r(xi)^2 + v0 - target |^2 with r = 2 xi / (1 + |xi|^2), by L-BFGS with the library
// defaults except g_epsilon = FLT_EPSILON, max_iterations = 128 (:790-793).  Shared by the host path (host_model.cpp) and the
// batched device kernel (geometry_kernels.cu, one thread per waypoint): plain sequential loops in the reference's order
// (lbfgs.hpp:422-487) on a caller-provided workspace, scalar line-search logic from lbfgs_core.cuh.  Compiled without FMA
// contraction on both sides, the two paths are bit-identical (and bit-identical to the reference's start point).
#pragma once
#include <cfloat>
#include <cmath>

#include "lbfgs_core.cuh"

namespace frx {

// doubles of workspace for a waypoint with k free variables (mem_size = 8: the library default, lbfgs.hpp:128)
FRX_HD int nls_work_doubles(int k) { return (6 + 2 * 8) * k + 2 * 8; }

// vp: 3 (k+1) doubles, "origin + edges" columns (se3gcopter_cpu.hpp:1049); r, gdr: k doubles of scratch
FRX_HD double nls_objective(const double* vp, int k, const double* target, const double* pp, double* grad, double* r, double* gdr) {
    double qn = 0.0;
    for (int a = 0; a < k; a++) qn += pp[a] * pp[a];
    const double q1 = qn + 1.0, q1s = q1 * q1;
    for (int a = 0; a < k; a++) r[a] = 2.0 / q1 * pp[a];
    double delta[3];
    for (int d = 0; d < 3; d++) {
        double acc = 0.0;
        for (int a = 0; a < k; a++) acc += vp[3 * (a + 1) + d] * (r[a] * r[a]);
        delta[d] = acc + vp[d] - target[d];
    }
    const double cost = delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2];
    const double g3[3] = {2 * delta[0], 2 * delta[1], 2 * delta[2]};
    double gp = 0.0;
    for (int a = 0; a < k; a++) {
        const double* e = vp + 3 * (a + 1);
        gdr[a] = (e[0] * g3[0] + e[1] * g3[1] + e[2] * g3[2]) * r[a] * 2.0;
    }
    for (int a = 0; a < k; a++) gp += gdr[a] * pp[a];
    for (int a = 0; a < k; a++) grad[a] = gdr[a] * 2.0 / q1 - pp[a] * 4.0 * gp / q1s;
    return cost;
}

FRX_HD double nls_dot(const double* a, const double* b, int n) {
    double s = 0.;
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

// x (in: start, out: solution), work: nls_work_doubles(k).  Returns the L-BFGS status (ignored by the reference).
FRX_HD int waypoint_backward(const double* vp, int k, const double* target, double* x, double* work) {
    LbfgsConfig P;
    P.g_epsilon = FLT_EPSILON;
    P.max_iterations = 128;
    const int n = k, m = P.mem_size;
    double *g = work, *xp = g + n, *gp = xp + n, *d = gp + n, *r = d + n, *gdr = r + n;
    double *S = gdr + n, *Y = S + (size_t)m * n, *al = Y + (size_t)m * n, *ysv = al + m;
    for (int a = 0; a < k; a++) x[a] = 1.0 / (sqrt(k + 1.0) + 1.0);
    double fx = nls_objective(vp, k, target, x, g, r, gdr);
    for (int i = 0; i < n; i++) d[i] = -g[i];
    double xnorm = sqrt(nls_dot(x, x, n)), gnorm = sqrt(nls_dot(g, g, n));
    if (xnorm < 1.0) xnorm = 1.0;
    if (gnorm / xnorm <= P.g_epsilon) return LBFGS_ALREADY_MINIMIZED;
    double step = 1.0 / sqrt(nls_dot(d, d, n));
    int kk = 1, end = 0;
    for (;;) {
        for (int i = 0; i < n; i++) xp[i] = x[i], gp[i] = g[i];
        const double stepp = step, fp = fx;
        LineSearch ls;
        int lsr = 0;
        for (int pass = 0; pass < 2; pass++) {  // More-Thuente, then backtracking from xp on failure (lbfgs.hpp:1268-1282)
            const bool mt = pass == 0;
            if (!mt) step = stepp, fx = fp;
            const double dg0 = nls_dot(gp, d, n);
            int st = mt ? ls.mt_begin(step, fx, dg0, P.min_step, P.max_step, P) : ls.bt_begin(step, fx, dg0, P.min_step, P.max_step, P);
            while (st == LS_CONTINUE) {
                for (int i = 0; i < n; i++) x[i] = xp[i];
                for (int i = 0; i < n; i++) x[i] += ls.step * d[i];
                fx = nls_objective(vp, k, target, x, g, r, gdr);
                const double dg = nls_dot(g, d, n);
                st = mt ? ls.mt_feed(fx, dg, P) : ls.bt_feed(fx, dg, P);
            }
            lsr = ls.code;
            if (lsr >= 0) break;
        }
        if (lsr < 0) {
            for (int i = 0; i < n; i++) x[i] = xp[i], g[i] = gp[i];
            return lsr;
        }
        step = ls.step;
        xnorm = sqrt(nls_dot(x, x, n)), gnorm = sqrt(nls_dot(g, g, n));
        if (xnorm < 1.0) xnorm = 1.0;
        if (gnorm / xnorm <= P.g_epsilon) return LBFGS_CONVERGENCE;
        if (P.max_iterations != 0 && P.max_iterations < kk + 1) return LBFGSERR_MAXIMUMITERATION;
        double *se = S + (size_t)end * n, *ye = Y + (size_t)end * n;
        for (int i = 0; i < n; i++) se[i] = x[i] - xp[i];
        for (int i = 0; i < n; i++) ye[i] = g[i] - gp[i];
        const double ys = nls_dot(ye, se, n), yy = nls_dot(ye, ye, n);
        ysv[end] = ys;
        const int bound = (m <= kk) ? m : kk;
        ++kk;
        end = (end + 1) % m;
        for (int i = 0; i < n; i++) d[i] = -g[i];
        int j = end;
        for (int i = 0; i < bound; i++) {
            j = (j + m - 1) % m;
            al[j] = nls_dot(S + (size_t)j * n, d, n) / ysv[j];
            const double c = -al[j];
            const double* yj = Y + (size_t)j * n;
            for (int q = 0; q < n; q++) d[q] += c * yj[q];
        }
        const double sc = ys / yy;
        for (int q = 0; q < n; q++) d[q] *= sc;
        for (int i = 0; i < bound; i++) {
            const double beta = nls_dot(Y + (size_t)j * n, d, n) / ysv[j];
            const double c = al[j] - beta;
            const double* sj = S + (size_t)j * n;
            for (int q = 0; q < n; q++) d[q] += c * sj[q];
            j = (j + 1) % m;
        }
        step = 1.0;
    }
}

}  // namespace frx
