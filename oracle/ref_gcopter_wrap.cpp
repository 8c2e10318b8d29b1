// ORACLE — TEST INFRASTRUCTURE ONLY.
// Compiles the reference's hot-path headers VERBATIM, from where they lie under /root/reference:
//     se3gcopter_cpu.hpp (MINCO_S3 + SE3GCOPTER), trajectory.hpp (BandedSystem, Piece, Trajectory), lbfgs.hpp
// over the Eigen-API shim in oracle/eigen_shim (the image has no Eigen).  root_finder.hpp (Piece::getMaxVelRate/AccRate,
// the printed report of MinCoPlan_CPU.cpp:131-132) is compiled verbatim too — its one use of a dense eigen-solver
// (eigenSolveRealRoots, root_finder.hpp:409) is never reached because the trajectory code always isolates roots.
// This file is compiled TWICE (oracle/Makefile):
//   * libfrx_ref_full.so — everything verbatim, including geoutils.hpp + sdlp.hpp + quickhull.hpp: SE3GCOPTER::setup runs the
//     reference's own LP + dual quickhull, so the V-polytopes (and their vertex ORDER, which defines the xi variables,
//     SURVEY.md App. B.5) are the reference's;
//   * libfrx_ref.so (-DREF_SUPPLIED_VPOLYS) — geoutils.hpp kept out by pre-defining its include guard; the V-polytopes are
//     SUPPLIED (the same ones the oracle / product use) through the geoutils::enumerateVs stub below, so that evaluation
//     parity of arbitrary workloads does not depend on vertex order.
// Built by oracle/Makefile into oracle/_ref/libfrx_ref.so; no reference source is copied into this repository.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <deque>
#include <set>
#include <vector>

#include <Eigen/Eigen>

#include "../include/frx.h"

#ifdef REF_SUPPLIED_VPOLYS
#define GEOUTILS_HPP
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
#endif  // REF_SUPPLIED_VPOLYS (otherwise: the verbatim geoutils.hpp + sdlp.hpp + quickhull.hpp)

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
#ifdef REF_SUPPLIED_VPOLYS
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
#else
    if (!p) return nullptr;  // full build: SE3GCOPTER::extractVs runs the reference's own LP + quickhull; supplied V-polytopes are ignored
#endif
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
// The V-polytopes the reference's setup holds (cfgVs, "origin + edges" columns, se3gcopter_cpu.hpp:1049), in ITS order.
// Two calls: verts == nullptr fills off[n_vpolys + 1]; then verts[3 * off[n_vpolys]].  Returns the number of V-polytopes.
int ref_vpolys(void* hv, int* off, double* verts) {
    SE3GCOPTER& o = ((RefHandle*)hv)->opt;
    const int nv = (int)o.cfgVs.size();
    off[0] = 0;
    for (int i = 0; i < nv; i++) {
        const int V = (int)o.cfgVs[i].cols();
        if (verts)
            for (int k = 0; k < V; k++)
                for (int d = 0; d < 3; d++) verts[3 * (size_t)(off[i] + k) + d] = o.cfgVs[i](d, k);
        off[i + 1] = off[i] + V;
    }
    return nv;
}
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

// ---- the reference's own Trajectory / Piece containers (trajectory.hpp:43-593), for the output side of the path ----
static Trajectory make_traj(const double* durations, const double* coeffs, int pieces) {
    Trajectory traj;
    for (int i = 0; i < pieces; i++) {
        CoefficientMat cm;
        for (int col = 0; col < 6; col++)
            for (int d = 0; d < 3; d++) cm(d, col) = coeffs[18 * i + 3 * col + d];  // 3x6 column-major, highest order first
        traj.emplace_back(durations[i], cm);
    }
    return traj;
}
// summary[0..2] = total duration, max velocity rate, max acceleration rate (MinCoPlan_CPU.cpp:131-133);
// per query time t: 24 doubles = pos, vel, acc, jer, rotation (row-major x|y|z columns -> xB, yB, zB), tilt rate;
// per piece: 18 doubles of normalizePosCoeffMat (3x6 column-major) = the payload of traj2msg (se3_planner.cpp:39-53).
void ref_traj_query(const double* durations, const double* coeffs, int pieces, const double* ts, int nt, double yaw, double g_acc,
                    double* summary, double* per_t, double* normalized) {
    const Trajectory traj = make_traj(durations, coeffs, pieces);
    summary[0] = traj.getTotalDuration();
    summary[1] = traj.getMaxVelRate();
    summary[2] = traj.getMaxAccRate();
    for (int k = 0; k < nt; k++) {
        double* o = per_t + 24 * k;
        const Eigen::Vector3d p = traj.getPos(ts[k]), v = traj.getVel(ts[k]), a = traj.getAcc(ts[k]), j = traj.getJer(ts[k]);
        for (int d = 0; d < 3; d++) o[d] = p(d), o[3 + d] = v(d), o[6 + d] = a(d), o[9 + d] = j(d);
        Eigen::Matrix3d R;
        traj.getRotation(ts[k], yaw, g_acc, R);
        for (int c = 0; c < 3; c++)
            for (int d = 0; d < 3; d++) o[12 + 3 * c + d] = R(d, c);
        const Eigen::Vector3d tr = traj.getTiltRate(ts[k], g_acc);
        for (int d = 0; d < 3; d++) o[21 + d] = tr(d);
    }
    for (int i = 0; i < pieces; i++) {
        const CoefficientMat n = traj[i].normalizePosCoeffMat();
        for (int col = 0; col < 6; col++)
            for (int d = 0; d < 3; d++) normalized[18 * i + 3 * col + d] = n(d, col);
    }
}
// checkMaxVelRate / checkMaxAccRate (trajectory.hpp:543-563) and the junction states (:481-515) at junction `junc`.
void ref_traj_checks(const double* durations, const double* coeffs, int pieces, double vel_limit, double acc_limit, int junc,
                     int* vel_ok, int* acc_ok, double* junc_pva) {
    const Trajectory traj = make_traj(durations, coeffs, pieces);
    *vel_ok = traj.checkMaxVelRate(vel_limit) ? 1 : 0;
    *acc_ok = traj.checkMaxAccRate(acc_limit) ? 1 : 0;
    const Eigen::Vector3d p = traj.getJuncPos(junc), v = traj.getJuncVel(junc), a = traj.getJuncAcc(junc);
    for (int d = 0; d < 3; d++) junc_pva[d] = p(d), junc_pva[3 + d] = v(d), junc_pva[6 + d] = a(d);
}

}  // extern "C"
