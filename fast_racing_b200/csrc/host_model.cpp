// Host-side problem model: replaces the body of SE3GCOPTER::setup (CPU.hpp:1076-1186) and the
// initialisation half of SE3GCOPTER::optimize (CPU.hpp:1237-1240).  See host_model.hpp.
#include "host_model.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

#include "lbfgs_core.cuh"
#include "waypoint_nls.hpp"

namespace frx {
namespace {

inline double norm3(const double* a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// centroid used by gridMesh / setInitial (CPU.hpp:1018-1019, 1206-1207):
//   rightCols(k).rowwise().sum() / (1 + k) + col(0)
void vpoly_centroid(const VPolytope& vp, double* out) {
    const int k = vp.cols() - 1;
    for (int d = 0; d < 3; d++) {
        double s = 0.0;
        for (int a = 0; a < k; a++) s += vp.v[3 * (a + 1) + d];
        out[d] = s / (1.0 + k) + vp.v[d];
    }
}

}  // namespace

int build_model(const frx_problem& p, TrajectoryModel& m) {
    if (p.n_poly < 1 || !p.face_off || !p.faces) return FRX_ERR_INVALID_ARG;
    if (p.params.itg_spaces < 1) return FRX_ERR_INVALID_ARG;
    m.par = p.params;
    m.c2 = p.params.c2_diffeo != 0;
    m.softT = p.params.rho > 0;  // CPU.hpp:1097-1107
    if (m.softT) {
        m.rho = p.params.rho;
        m.sumT = 1.0;
    } else {
        m.rho = 0.0;
        m.sumT = p.params.total_t;
    }
    std::memcpy(m.ini, p.ini, sizeof(m.ini));
    std::memcpy(m.fin, p.fin, sizeof(m.fin));
    for (int d = 0; d < 3; d++) m.ini_clamped_p[d] = p.ini[d], m.fin_clamped_p[d] = p.fin[d];
    m.coarseN = p.n_poly;
    m.face_off.assign(p.face_off, p.face_off + p.n_poly + 1);
    if (m.face_off[0] != 0) return FRX_ERR_INVALID_ARG;
    for (int i = 0; i < p.n_poly; i++)
        if (m.face_off[i + 1] - m.face_off[i] < 4) return FRX_ERR_INVALID_ARG;
    m.faces.assign(p.faces, p.faces + 6 * (size_t)m.face_off[p.n_poly]);
    // CPU.hpp:1116: cfgHs[i].topRows<3>().colwise().normalize()
    for (int k = 0; k < m.face_off[p.n_poly]; k++) {
        double* f = &m.faces[6 * (size_t)k];
        const double nn = norm3(f);
        f[0] /= nn, f[1] /= nn, f[2] /= nn;
    }
    // V-polytopes: supplied, or extractVs (CPU.hpp:1031-1074): P_0, P_0^P_1, P_1, ..., P_{N-1}
    const int nv = 2 * p.n_poly - 1;
    m.vpolys.assign(nv, VPolytope());
    if (p.vert_off && p.verts) {
        for (int i = 0; i < nv; i++) {
            if (p.vert_off[i + 1] - p.vert_off[i] < 2) return FRX_ERR_INVALID_ARG;
            m.vpolys[i].v.assign(p.verts + 3 * (size_t)p.vert_off[i], p.verts + 3 * (size_t)p.vert_off[i + 1]);
        }
    } else {
        std::vector<double> both, verts;
        for (int i = 0; i < nv; i++) {
            const int a = i / 2;
            const double* fa = &m.faces[6 * (size_t)m.face_off[a]];
            const int Fa = m.face_off[a + 1] - m.face_off[a];
            bool ok;
            if (i % 2 == 0) {
                ok = enumerate_vertices(fa, Fa, verts);
            } else {
                const int Fb = m.face_off[a + 2] - m.face_off[a + 1];
                both.assign(fa, fa + 6 * (size_t)(Fa + Fb));  // contiguous: [hPs[i], hPs[i+1]] (CPU.hpp:1052-1053)
                ok = enumerate_vertices(both.data(), Fa + Fb, verts);
            }
            if (!ok) return FRX_ERR_INFEASIBLE_GEOMETRY;
            // origin + edge form, CPU.hpp:1049
            VPolytope& vp = m.vpolys[i];
            vp.v = verts;
            for (size_t c = 1; c < verts.size() / 3; c++)
                for (int d = 0; d < 3; d++) vp.v[3 * c + d] = verts[3 * c + d] - verts[d];
        }
    }
    // gridMesh, CPU.hpp:1003-1029
    m.intervals.assign(m.coarseN, 1);
    {
        double lastP[3], curP[3] = {m.ini_clamped_p[0], m.ini_clamped_p[1], m.ini_clamped_p[2]};
        for (int i = 0; i < m.coarseN - 1; i++) {
            std::memcpy(lastP, curP, sizeof(curP));
            vpoly_centroid(m.vpolys[2 * i + 1], curP);
            const double dd[3] = {curP[0] - lastP[0], curP[1] - lastP[1], curP[2] - lastP[2]};
            const int cur = (int)std::ceil(norm3(dd) / p.params.grid_res);
            m.intervals[i] = cur > 0 ? cur : 1;
        }
        std::memcpy(lastP, curP, sizeof(curP));
        const double dd[3] = {m.fin_clamped_p[0] - lastP[0], m.fin_clamped_p[1] - lastP[1], m.fin_clamped_p[2] - lastP[2]};
        const int cur = (int)std::ceil(norm3(dd) / p.params.grid_res);
        m.intervals[m.coarseN - 1] = cur > 0 ? cur : 1;
    }
    // index tables, CPU.hpp:1125-1152
    m.fineN = 0;
    for (int v : m.intervals) m.fineN += v;
    m.cons.assign(m.fineN, p.params.itg_spaces);
    m.idxVs.assign(std::max(m.fineN - 1, 0), 0);
    m.idxHs.assign(m.fineN, 0);
    m.xi_off.assign(m.fineN, 0);
    m.dimT = m.softT ? m.coarseN : m.coarseN - 1;
    m.dimP = 0;
    int offset = 0;
    for (int i = 0; i < m.coarseN; i++) {
        const int interval = m.intervals[i];
        for (int j = 0; j < interval; j++) {
            if (j < interval - 1) {
                m.idxVs[offset] = 2 * i;
                m.xi_off[offset] = m.dimP;
                m.dimP += m.vpolys[2 * i].cols() - 1;
            } else if (i < m.coarseN - 1) {
                m.idxVs[offset] = 2 * i + 1;
                m.xi_off[offset] = m.dimP;
                m.dimP += m.vpolys[2 * i + 1].cols() - 1;
            }
            m.idxHs[offset] = i;
            offset++;
        }
    }
    m.xi_off[m.fineN - 1] = m.dimP;
    if (m.n() <= 0) return FRX_ERR_INVALID_ARG;
    return FRX_OK;
}

void initial_guess_prepare(const TrajectoryModel& m, double* x, std::vector<double>& innerP) {
    // setInitial, CPU.hpp:1188-1228
    std::vector<double> coarseT(m.coarseN);
    innerP.assign(3 * (size_t)std::max(m.fineN - 1, 0), 0.0);
    const double spd = std::min(m.par.v_max, 10.0);
    {
        double lastP[3], curP[3] = {m.ini_clamped_p[0], m.ini_clamped_p[1], m.ini_clamped_p[2]}, delta[3];
        int offset = 0;
        for (int i = 0; i < m.coarseN - 1; i++) {
            std::memcpy(lastP, curP, sizeof(curP));
            const int interv = m.intervals[i];
            vpoly_centroid(m.vpolys[2 * i + 1], curP);
            for (int d = 0; d < 3; d++) delta[d] = curP[d] - lastP[d];
            coarseT[i] = norm3(delta) / spd;
            for (int d = 0; d < 3; d++) delta[d] /= interv;
            for (int j = 0; j < interv; j++) {
                for (int d = 0; d < 3; d++) innerP[3 * (size_t)offset + d] = (j + 1) * delta[d] + lastP[d];
                offset++;
            }
        }
        const int interv = m.intervals[m.coarseN - 1];
        std::memcpy(lastP, curP, sizeof(curP));
        for (int d = 0; d < 3; d++) delta[d] = m.fin_clamped_p[d] - lastP[d];
        coarseT[m.coarseN - 1] = norm3(delta) / spd;
        for (int d = 0; d < 3; d++) delta[d] /= interv;
        for (int j = 0; j < interv - 1; j++) {
            for (int d = 0; d < 3; d++) innerP[3 * (size_t)offset + d] = (j + 1) * delta[d] + lastP[d];
            offset++;
        }
    }
    // backwardT, CPU.hpp:679-726
    if (m.softT) {
        for (int i = 0; i < m.coarseN; i++)
            x[i] = m.c2 ? (coarseT[i] > 1.0 ? (std::sqrt(2.0 * coarseT[i] - 1.0) - 1.0) : (1.0 - std::sqrt(2.0 / coarseT[i] - 1.0)))
                        : std::log(coarseT[i]);
    } else {
        const int Ms1 = m.dimT;
        for (int i = 0; i < Ms1; i++) {
            const double r = coarseT[i] / coarseT[Ms1];
            x[i] = m.c2 ? (r > 1.0 ? (std::sqrt(2.0 * r - 1.0) - 1.0) : (1.0 - std::sqrt(2.0 / r - 1.0))) : std::log(r);
        }
    }
}

void initial_guess(const TrajectoryModel& m, double* x) {
    std::vector<double> innerP;
    initial_guess_prepare(m, x, innerP);
    // backwardP, CPU.hpp:777-813: per waypoint, min || V r(xi)^2 + v0 - q ||^2 by L-BFGS (waypoint_nls.hpp)
    double* P = x + m.dimT;
    std::vector<double> work;
    for (int i = 0; i < m.fineN - 1; i++) {
        const VPolytope& vp = m.vpolys[m.idxVs[i]];
        const int k = vp.cols() - 1;
        work.resize((size_t)nls_work_doubles(k));
        waypoint_backward(vp.v.data(), k, &innerP[3 * (size_t)i], P + m.xi_off[i], work.data());
    }
}

}  // namespace frx
