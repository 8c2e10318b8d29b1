// ORACLE — TEST INFRASTRUCTURE ONLY.
// Compiles the reference's hot-path headers VERBATIM, from where they lie under /root/reference:
//     se3gcopter_cpu.hpp (MINCO_S3 + SE3GCOPTER), trajectory.hpp (BandedSystem, Piece, Trajectory), lbfgs.hpp
// over the Eigen-API shim in oracle/eigen_shim (the image has no Eigen).  Two of the reference's headers are
// kept OUT by pre-defining their include guards, because they are off the hot path and need far more of Eigen:
//   * root_finder.hpp — only used by Piece::getMaxVelRate/AccRate for the printed report (MinCoPlan_CPU.cpp:131-132);
//   * geoutils.hpp (+ sdlp.hpp, quickhull.hpp) — the H->V polytope conversion of SE3GCOPTER::extractVs; the V-polytopes
//     are SUPPLIED instead (the same ones the oracle / product use), through the geoutils::enumerateVs stub below.
// Built by oracle/Makefile into oracle/_ref/libfrx_ref.so; no reference source is copied into this repository.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <deque>
#include <set>
#include <vector>

#include <Eigen/Eigen>

#include "../include/frx.h"

#define ROOT_FINDER_HPP
#define GEOUTILS_HPP
namespace RootFinder {  // reporting-only helpers: never called through this wrapper
template <class A>
inline Eigen::VectorXd polySqr(const A&) { return Eigen::VectorXd(1); }
template <class A>
inline double polyVal(const A&, double) { return 1.0; }
template <class A>
inline std::set<double> solvePolynomial(const A&, double, double, double, bool = true) { return std::set<double>(); }
template <class A>
inline int countRoots(const A&, double, double) { return 0; }
}  // namespace RootFinder
namespace geoutils {
// V-polytopes (absolute vertex coordinates, 3 x V) handed to the reference's extractVs in call order.
static std::deque<Eigen::MatrixXd> g_supplied;
inline bool enumerateVs(const Eigen::MatrixXd&, Eigen::MatrixXd& vPoly, const double = 1.0e-6) {
    if (g_supplied.empty()) return false;
    vPoly = g_supplied.front();
    g_supplied.pop_front();
    return true;
}
}  // namespace geoutils

#define private public
#include "se3gcopter_cpu.hpp"  // -I /root/reference/src/plan_manage/include/se3gcopter
#undef private

struct RefHandle {
    SE3GCOPTER opt;
    int n = 0;
};

extern "C" {

// Same problem description as the product ABI; V-polytopes are required (origin+edge form, as frx_batch_vpolys returns).
void* ref_create(const frx_problem* p) {
    if (!p || !p->vert_off || !p->verts) return nullptr;
    geoutils::g_supplied.clear();
    const int nv = 2 * p->n_poly - 1;
    for (int i = 0; i < nv; i++) {
        const int V = p->vert_off[i + 1] - p->vert_off[i];
        Eigen::MatrixXd m(3, V);
        const double* v = p->verts + 3 * (size_t)p->vert_off[i];
        for (int k = 0; k < V; k++)
            for (int d = 0; d < 3; d++) m(d, k) = (k == 0) ? v[d] : v[3 * k + d] + v[d];  // absolute vertex = v0 + edge
        geoutils::g_supplied.push_back(m);
    }
    std::vector<Eigen::MatrixXd> hPolys;
    for (int i = 0; i < p->n_poly; i++) {
        const int F = p->face_off[i + 1] - p->face_off[i];
        Eigen::MatrixXd h(6, F);
        for (int k = 0; k < F; k++)
            for (int d = 0; d < 6; d++) h(d, k) = p->faces[6 * (size_t)(p->face_off[i] + k) + d];
        hPolys.push_back(h);
    }
    Eigen::MatrixXd ini(3, 3), fin(3, 3);
    for (int k = 0; k < 9; k++) ini.d[k] = p->ini[k], fin.d[k] = p->fin[k];
    Eigen::Vector4d w;
    for (int k = 0; k < 4; k++) w(k) = p->params.chi[k];
    RefHandle* h = new RefHandle();
    const frx_params& q = p->params;
    const bool ok = h->opt.setup(q.rho, q.total_t, ini, fin, hPolys, q.grid_res, q.itg_spaces, q.horiz_half_len, q.vert_half_len,
                                 q.safe_margin, q.v_max, q.thr_acc_min, q.thr_acc_max, q.body_rate_max, q.grav_acc, w,
                                 q.c2_diffeo != 0);
    if (!ok) {
        delete h;
        return nullptr;
    }
    h->n = h->opt.dimFreeT + h->opt.dimFreeP;
    return h;
}
void ref_destroy(void* hv) { delete (RefHandle*)hv; }
int ref_dims(void* hv, int* n, int* pieces, int* dim_t, int* dim_p) {
    RefHandle* h = (RefHandle*)hv;
    if (n) *n = h->n;
    if (pieces) *pieces = h->opt.fineN;
    if (dim_t) *dim_t = h->opt.dimFreeT;
    if (dim_p) *dim_p = h->opt.dimFreeP;
    return 0;
}
int ref_tables(void* hv, int* intervals, int* idx_hs, int* idx_vs, int* cons) {
    SE3GCOPTER& o = ((RefHandle*)hv)->opt;
    for (int i = 0; i < o.intervals.size(); i++) intervals[i] = o.intervals(i);
    for (int i = 0; i < o.idxHs.size(); i++) idx_hs[i] = o.idxHs(i);
    for (int i = 0; i < o.idxVs.size(); i++) idx_vs[i] = o.idxVs(i);
    for (int i = 0; i < o.cons.size(); i++) cons[i] = o.cons(i);
    return 0;
}
// The reference's own evaluator (SE3GCOPTER::objectiveFunc, se3gcopter_cpu.hpp:961) — lbfgs_evaluate_t signature.
double ref_eval(void* hv, const double* x, double* g, const int n) { return SE3GCOPTER::objectiveFunc(&((RefHandle*)hv)->opt, x, g, n); }
// Starting point exactly as SE3GCOPTER::optimize builds it (se3gcopter_cpu.hpp:1237-1240).
void ref_initial_guess(void* hv, double* x) {
    SE3GCOPTER& o = ((RefHandle*)hv)->opt;
    Eigen::Map<Eigen::VectorXd> t(x, o.dimFreeT);
    Eigen::Map<Eigen::VectorXd> p(x + o.dimFreeT, o.dimFreeP);
    o.setInitial(o.cfgVs, o.intervals, o.coarseT, o.innerP);
    SE3GCOPTER::backwardT(o.coarseT, t, o.softT, o.c2dfm);
    SE3GCOPTER::backwardP(o.innerP, o.idxVs, o.cfgVs, p);
}
// Coefficients after the reference's forward pass at x: T[pieces], c[18*pieces] (rows 6i..6i+5, row-major 6x3 blocks).
void ref_forward(void* hv, const double* x, double* T, double* c) {
    SE3GCOPTER& o = ((RefHandle*)hv)->opt;
    Eigen::Map<const Eigen::VectorXd> t(x, o.dimFreeT);
    Eigen::Map<const Eigen::VectorXd> p(x + o.dimFreeT, o.dimFreeP);
    SE3GCOPTER::forwardT(t, o.coarseT, o.softT, o.sumT, o.c2dfm);
    SE3GCOPTER::splitToFineT(o.coarseT, o.intervals, o.fineT);
    SE3GCOPTER::forwardP(p, o.idxVs, o.cfgVs, o.innerP);
    o.jerkOpt.generate(o.innerP, o.fineT);
    for (int i = 0; i < o.fineN; i++) T[i] = o.fineT(i);
    for (int r = 0; r < 6 * o.fineN; r++)
        for (int d = 0; d < 3; d++) c[3 * r + d] = o.jerkOpt.b(r, d);
}
// The reference's own optimize() (se3gcopter_cpu.hpp:1230): returns the jerk cost; durations / coefficients (3x6
// column-major, highest order first) of the resulting Trajectory.
double ref_optimize(void* hv, double rel_tol, double* durations, double* coeffs) {
    SE3GCOPTER& o = ((RefHandle*)hv)->opt;
    Trajectory traj;
    const double j = o.optimize(traj, rel_tol);
    for (int i = 0; i < traj.getPieceNum(); i++) {
        durations[i] = traj[i].getDuration();
        const CoefficientMat& cm = traj[i].getCoeffMat();
        for (int col = 0; col < 6; col++)
            for (int d = 0; d < 3; d++) coeffs[18 * i + 3 * col + d] = cm(d, col);
    }
    return j;
}

}  // extern "C"
