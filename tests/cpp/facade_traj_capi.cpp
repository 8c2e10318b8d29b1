// C wrapper around the header-only frx::Trajectory / frx::Piece (include/se3gcopter_b200.hpp) with the SAME entry points as
// the reference-side wrapper in oracle/ref_gcopter_wrap.cpp (ref_traj_query / ref_traj_checks), so that
// tests/test_facade_cpp.py can compare the two containers number by number.  CPU only, no CUDA, libfrx not linked.
#include "se3gcopter_b200.hpp"

static frx::Trajectory make_traj(const double* durations, const double* coeffs, int pieces) {
    frx::Trajectory traj;
    for (int i = 0; i < pieces; i++) traj.emplace_back(durations[i], coeffs + 18 * i);
    return traj;
}

extern "C" {
void facade_traj_query(const double* durations, const double* coeffs, int pieces, const double* ts, int nt, double yaw, double g_acc,
                       double* summary, double* per_t, double* normalized, double* wire) {
    const frx::Trajectory traj = make_traj(durations, coeffs, pieces);
    summary[0] = traj.getTotalDuration();
    summary[1] = traj.getMaxVelRate();
    summary[2] = traj.getMaxAccRate();
    for (int k = 0; k < nt; k++) {
        double* o = per_t + 24 * k;
        const auto p = traj.getPos(ts[k]), v = traj.getVel(ts[k]), a = traj.getAcc(ts[k]), j = traj.getJer(ts[k]);
        for (int d = 0; d < 3; d++) o[d] = p[d], o[3 + d] = v[d], o[6 + d] = a[d], o[9 + d] = j[d];
        traj.getRotation(ts[k], yaw, g_acc, o + 12);
        const auto tr = traj.getTiltRate(ts[k], g_acc);
        for (int d = 0; d < 3; d++) o[21 + d] = tr[d];
    }
    for (int i = 0; i < pieces; i++) {
        const auto n = traj[i].normalizePosCoeffMat();
        for (int q = 0; q < 18; q++) normalized[18 * i + q] = n[q];
    }
    // wire[pieces][19]: per segment coef_x[6], coef_y[6], coef_z[6], time  — the traj2msg payload (se3_planner.cpp:39-53)
    const frx::PolynomialTrajectoryFields m = frx::traj2fields(traj);
    for (unsigned i = 0; i < m.num_segment; i++) {
        for (int j = 0; j < 6; j++) {
            wire[19 * i + j] = m.coef_x[6 * i + j];
            wire[19 * i + 6 + j] = m.coef_y[6 * i + j];
            wire[19 * i + 12 + j] = m.coef_z[6 * i + j];
        }
        wire[19 * i + 18] = m.time[i];
    }
}
void facade_traj_checks(const double* durations, const double* coeffs, int pieces, double vel_limit, double acc_limit, int junc,
                        int* vel_ok, int* acc_ok, double* junc_pva) {
    const frx::Trajectory traj = make_traj(durations, coeffs, pieces);
    *vel_ok = traj.checkMaxVelRate(vel_limit) ? 1 : 0;
    *acc_ok = traj.checkMaxAccRate(acc_limit) ? 1 : 0;
    const auto p = traj.getJuncPos(junc), v = traj.getJuncVel(junc), a = traj.getJuncAcc(junc);
    for (int d = 0; d < 3; d++) junc_pva[d] = p[d], junc_pva[3 + d] = v[d], junc_pva[6 + d] = a[d];
}
}
