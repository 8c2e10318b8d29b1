// ORACLE — TEST INFRASTRUCTURE ONLY (see lbfgs_rc.hpp header).
// ctypes-facing C API of the CPU oracle.  Takes the same frx_problem description as the product
// ABI (include/frx.h is included only for the argument structs), but REQUIRES the V-polytopes to
// be supplied (the tests pass the product's, after checking them against brute_force_vertices).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <thread>
#include <vector>

#include "../include/frx.h"
#include "gcopter_oracle.hpp"
#include "geometry_oracle.hpp"

using namespace frx_oracle;

struct OracleHandle {
    Se3Gcopter opt;
    bool ok = false;
};

extern "C" {

void* fro_create(const frx_problem* p) {
    if (!p || !p->vert_off || !p->verts || p->n_poly < 1) return nullptr;
    auto* h = new OracleHandle();
    std::vector<HPoly> hp(p->n_poly);
    for (int i = 0; i < p->n_poly; i++)
        hp[i].f.assign(p->faces + 6 * (size_t)p->face_off[i], p->faces + 6 * (size_t)p->face_off[i + 1]);
    const int nv = 2 * p->n_poly - 1;
    std::vector<VPoly> vp(nv);
    for (int i = 0; i < nv; i++)
        vp[i].v.assign(p->verts + 3 * (size_t)p->vert_off[i], p->verts + 3 * (size_t)p->vert_off[i + 1]);
    const frx_params& q = p->params;
    h->ok = h->opt.setup(q.rho, q.total_t, p->ini, p->fin, hp, vp, q.grid_res, q.itg_spaces, q.horiz_half_len,
                         q.vert_half_len, q.safe_margin, q.v_max, q.thr_acc_min, q.thr_acc_max, q.body_rate_max,
                         q.grav_acc, q.chi, q.c2_diffeo != 0);
    if (!h->ok) {
        delete h;
        return nullptr;
    }
    return h;
}
void fro_destroy(void* h) { delete (OracleHandle*)h; }

int fro_dims(void* hv, int* n, int* pieces, int* dim_t, int* dim_p) {
    auto* h = (OracleHandle*)hv;
    if (n) *n = h->opt.dim();
    if (pieces) *pieces = h->opt.fineN;
    if (dim_t) *dim_t = h->opt.dimFreeT;
    if (dim_p) *dim_p = h->opt.dimFreeP;
    return 0;
}
int fro_tables(void* hv, int* intervals, int* idx_hs, int* idx_vs, int* cons) {
    auto& o = ((OracleHandle*)hv)->opt;
    if (intervals) std::copy(o.intervals.begin(), o.intervals.end(), intervals);
    if (idx_hs) std::copy(o.idxHs.begin(), o.idxHs.end(), idx_hs);
    if (idx_vs) std::copy(o.idxVs.begin(), o.idxVs.end(), idx_vs);
    if (cons) std::copy(o.cons.begin(), o.cons.end(), cons);
    return 0;
}
void fro_initial_guess(void* hv, double* x) { ((OracleHandle*)hv)->opt.initialGuess(x); }

// lbfgs_evaluate_t-compatible (lbfgs.hpp:223-226)
double fro_eval(void* hv, const double* x, double* g, const int n) {
    return Se3Gcopter::objectiveFunc(&((OracleHandle*)hv)->opt, x, g, n);
}

// Intermediate quantities of one evaluation, for component-level parity:
// T[pieces], q[3*(pieces-1)] col-major, c[18*pieces] (6x3 row-major blocks = rows 6i..6i+5 of b).
int fro_forward(void* hv, const double* x, double* T, double* q, double* c) {
    auto& o = ((OracleHandle*)hv)->opt;
    Se3Gcopter::forwardT(x, o.dimFreeT, o.coarseT, o.softT, o.sumT, o.c2dfm);
    Se3Gcopter::splitToFineT(o.coarseT, o.intervals, o.fineT);
    Se3Gcopter::forwardP(x + o.dimFreeT, o.idxVs, o.cfgVs, o.innerP);
    o.jerkOpt.generate(o.innerP.data(), o.fineT.data());
    if (T) std::copy(o.fineT.begin(), o.fineT.end(), T);
    if (q) std::copy(o.innerP.begin(), o.innerP.end(), q);
    if (c) std::copy(o.jerkOpt.b.begin(), o.jerkOpt.b.end(), c);
    return 0;
}
// Penalty integral alone on the coefficients of the last fro_forward: cost, gdT[pieces], gdC[18*pieces]
// (CPU.hpp:188-408 only; energy terms excluded).
int fro_penalty_only(void* hv, double* cost, double* gdT, double* gdC) {
    auto& o = ((OracleHandle*)hv)->opt;
    std::vector<double> gC((size_t)18 * o.fineN, 0.0);
    std::vector<double> gT(o.fineN, 0.0);
    double c = 0.0;
    o.jerkOpt.addTimeIntPenalty(o.cons.data(), o.idxHs.data(), o.cfgHs, o.pen, c, gT.data(), gC);
    if (cost) *cost = c;
    if (gdT) std::copy(gT.begin(), gT.end(), gdT);
    if (gdC) std::copy(gC.begin(), gC.end(), gdC);
    return 0;
}

double fro_optimize(void* hv, double rel_tol, double* durations, double* coeffs, double* x_final, double* final_cost,
                    int* status, int* iters, int* evals) {
    auto& o = ((OracleHandle*)hv)->opt;
    o.n_evals = 0;
    std::vector<double> dur(o.fineN), co((size_t)18 * o.fineN);
    const double j = o.optimize(rel_tol, dur.data(), co.data(), x_final, final_cost);
    if (durations) std::copy(dur.begin(), dur.end(), durations);
    if (coeffs) std::copy(co.begin(), co.end(), coeffs);
    if (status) *status = o.last_status;
    if (iters) *iters = o.last_iters;
    if (evals) *evals = o.n_evals;
    return j;
}

// Safety net for timing runs only (see Se3Gcopter::safety_max_iterations): 0 restores the reference's "no limit".
void fro_set_max_iterations(void* hv, int max_iterations) { ((OracleHandle*)hv)->opt.safety_max_iterations = max_iterations; }

// L-BFGS with the reference's optimize() parameters (CPU.hpp:1243-1247) from an arbitrary start x (in/out).
double fro_optimize_from(void* hv, double* x, double rel_tol) {
    auto& o = ((OracleHandle*)hv)->opt;
    LbfgsParams lp;
    lp.mem_size = 128, lp.past = 3, lp.g_epsilon = 1.0e-16, lp.min_step = 1.0e-32, lp.delta = rel_tol;
    double fx = 0.0;
    o.last_status = lbfgs_minimize(o.dim(), x, &fx, &Se3Gcopter::objectiveFunc, &o, lp, &o.last_iters, nullptr);
    return fx;
}

// ---- timing helpers for bench.py's cpu_baseline -------------------------------------------
// Seconds for `reps` evaluations at fixed x on one core.
double fro_time_evals(void* hv, const double* x, int reps) {
    auto& o = ((OracleHandle*)hv)->opt;
    std::vector<double> g(o.dim());
    volatile double sink = 0.0;
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; r++) sink = sink + Se3Gcopter::objectiveFunc(&o, x, g.data(), o.dim());
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}
// All-cores variant: handles[i] evaluated reps times each, spread over `threads` std::threads
// (one independent optimiser per handle, as the reference is single-threaded per solve).
double fro_time_evals_mt(void** handles, const double** xs, int count, int reps, int threads) {
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([=]() {
            for (int i = t; i < count; i += threads) {
                auto& o = ((OracleHandle*)handles[i])->opt;
                std::vector<double> g(o.dim());
                for (int r = 0; r < reps; r++) Se3Gcopter::objectiveFunc(&o, xs[i], g.data(), o.dim());
            }
        });
    for (auto& th : pool) th.join();
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}
// Full solves on `threads` threads; returns seconds; per-handle outputs as in fro_optimize.
double fro_time_solves_mt(void** handles, int count, double rel_tol, int threads, double* final_cost, int* status,
                          int* iters, int* evals) {
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([=]() {
            for (int i = t; i < count; i += threads)
                fro_optimize(handles[i], rel_tol, nullptr, nullptr, nullptr, final_cost ? final_cost + i : nullptr,
                             status ? status + i : nullptr, iters ? iters + i : nullptr, evals ? evals + i : nullptr);
        });
    for (auto& th : pool) th.join();
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- geometry check ------------------------------------------------------------------------
// Brute-force vertices of one H-polytope (unit normals).  Returns V; writes up to cap vertices.
int fro_brute_force_vertices(const double* faces, int F, double* out, int cap) {
    auto v = brute_force_vertices(faces, F);
    const int V = (int)v.size() / 3;
    for (int i = 0; i < V && i < cap; i++) {
        out[3 * i] = v[3 * i], out[3 * i + 1] = v[3 * i + 1], out[3 * i + 2] = v[3 * i + 2];
    }
    return V;
}

// V-polytopes of a corridor by the oracle's own brute-force enumeration, in the layout SE3GCOPTER::extractVs builds
// (CPU.hpp:1031-1074: polytope i at slot 2i, the intersection of i and i+1 at slot 2i+1, "origin + edges" columns): lets
// bench.py's reference arm run without the product library.  Vertex ORDER is the enumeration's, not quickhull's.
// Two calls: verts == nullptr fills off[2*n_poly]; then verts[3*off[2*n_poly-1]].  Returns 0, or -2 on an empty polytope.
int fro_vpolys(const frx_problem* p, int* off, double* verts) {
    const int nv = 2 * p->n_poly - 1;
    off[0] = 0;
    for (int k = 0; k < nv; k++) {
        const int i = k / 2;
        std::vector<double> f;
        auto add = [&](int poly) {
            for (int a = p->face_off[poly]; a < p->face_off[poly + 1]; a++) {
                const double* src = p->faces + 6 * (size_t)a;
                const double nn = std::sqrt(src[0] * src[0] + src[1] * src[1] + src[2] * src[2]);
                for (int q = 0; q < 3; q++) f.push_back(src[q] / nn);  // normals normalised as CPU.hpp:1116
                for (int q = 3; q < 6; q++) f.push_back(src[q]);
            }
        };
        add(i);
        if (k & 1) add(i + 1);
        auto v = brute_force_vertices(f.data(), (int)f.size() / 6);
        const int V = (int)v.size() / 3;
        if (V < 4) return -2;
        if (verts) {
            double* dst = verts + 3 * (size_t)off[k];
            for (int c = 0; c < V; c++)
                for (int q = 0; q < 3; q++) dst[3 * c + q] = c == 0 ? v[q] : v[3 * c + q] - v[q];
        }
        off[k + 1] = off[k] + V;
    }
    return 0;
}

// ---- stand-alone pieces for unit tests -------------------------------------------------------
// Banded solve of A x = b with A given dense (N x N row-major, bandwidth 6/6) — checks Banded
// against dense Gaussian elimination done in the test.
int fro_banded_solve(const double* dense, int N, double* b /* N x 3 row-major, in/out */, int adjoint) {
    Banded A;
    A.create(N, 6, 6);
    for (int i = 0; i < N; i++)
        for (int j = std::max(0, i - 6); j <= std::min(N - 1, i + 6); j++) A.at(i, j) = dense[(size_t)i * N + j];
    A.factorizeLU();
    if (adjoint)
        A.solveAdj(b);
    else
        A.solve(b);
    return 0;
}

// Restated L-BFGS on built-in analytic test functions (to compare with the verbatim reference
// lbfgs.hpp in oracle/_ref): kind 0 = extended Rosenbrock, 1 = a badly scaled quartic.
static double test_fn(void* inst, const double* x, double* g, const int n) {
    const int kind = *(int*)inst;
    double f = 0.0;
    if (kind == 0) {
        for (int i = 0; i < n; i++) g[i] = 0.0;
        for (int i = 0; i + 1 < n; i += 2) {
            const double t1 = 1.0 - x[i];
            const double t2 = 10.0 * (x[i + 1] - x[i] * x[i]);
            g[i + 1] = 20.0 * t2;
            g[i] = -2.0 * (x[i] * g[i + 1] + t1);
            f += t1 * t1 + t2 * t2;
        }
    } else {
        for (int i = 0; i < n; i++) {
            const double s = 1.0 + 10.0 * i;
            const double d = x[i] - 0.1 * i;
            f += s * d * d * d * d + 0.5 * d * d;
            g[i] = 4.0 * s * d * d * d + d;
        }
    }
    return f;
}
int fro_lbfgs_testfn(int kind, int n, double* x, double* fx, int mem_size, int past, double delta, double g_eps,
                     int max_iter, int* iters, int* evals) {
    LbfgsParams p;
    p.mem_size = mem_size, p.past = past, p.delta = delta, p.g_epsilon = g_eps, p.max_iterations = max_iter;
    return lbfgs_minimize(n, x, fx, test_fn, &kind, p, iters, evals);
}
double fro_testfn_eval(int kind, int n, const double* x, double* g) { return test_fn(&kind, x, g, n); }

}  // extern "C"
