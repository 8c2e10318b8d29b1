// Shape of the reference's caller (MavGlobalPlanner::plan, MinCoPlan_CPU.cpp:93-135) against the B200 façade:
// build hPolys, SE3GCOPTER::setup, SE3GCOPTER::optimize, print the report.  Exit code 0 = plausible result.
#include <cmath>
#include <cstdio>
#include <vector>

#include "se3gcopter_b200.hpp"

static std::vector<double> box(double cx, double cy, double cz, double hx, double hy, double hz) {
    const double n[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    const double h[6] = {hx, hx, hy, hy, hz, hz};
    std::vector<double> f;
    for (int k = 0; k < 6; k++) {
        for (int d = 0; d < 3; d++) f.push_back(n[k][d] * 2.0);  // un-normalised on purpose: setup normalises (CPU.hpp:1116)
        f.push_back(cx + n[k][0] * h[k]);
        f.push_back(cy + n[k][1] * h[k]);
        f.push_back(cz + n[k][2] * h[k]);
    }
    return f;
}

int main() {
    std::vector<std::vector<double>> hPolys;
    for (int i = 0; i < 5; i++) hPolys.push_back(box(3.0 * i + 1.5, 0.4 * i, 1.5, 3.0, 2.5, 1.4));
    double ini[9] = {0, 0, 1.5, 0, 0, 0, 0, 0, 0}, fin[9] = {15.0, 1.6, 1.5, 0, 0, 0, 0, 0, 0};
    const double w[4] = {1.0e7, 1.0e4, 1.0e4, 1.0e4};  // PenaltyPVTB of zhangjiajie_params.yaml
    frx::SE3GCOPTER nonlinOpt;
    frx::Trajectory traj;
    if (!nonlinOpt.setup(1000.0, 0.0, ini, fin, hPolys, INFINITY, 48, 0.5, 0.15, 0.08, 14.0, 5.0, 12.0, 3.8, 9.81, w, true)) {
        std::printf("setup: infeasible corridor\n");
        return 2;
    }
    const double finalObj = nonlinOpt.optimize(traj, 1.0e-6);
    nonlinOpt.kill_kernel();
    std::printf("Final cost: %.6f\nMaximum Vel: %.4f\nMaximum Acc: %.4f\nTotal Duation: %.4f\npieces %d lbfgs_status %d\n", finalObj,
                traj.getMaxVelRate(), traj.getMaxAccRate(), traj.getTotalDuration(), traj.getPieceNum(), nonlinOpt.lbfgs_status());
    const auto p0 = traj.getPos(0.0), p1 = traj.getPos(traj.getTotalDuration());
    const bool ok = std::isfinite(finalObj) && finalObj > 0 && traj.getPieceNum() == 5 && std::fabs(p0[0] - ini[0]) < 1e-9 &&
                    std::fabs(p1[0] - fin[0]) < 1e-6 && std::fabs(p1[1] - fin[1]) < 1e-6 && traj.getMaxVelRate() < 14.0 * 1.05 &&
                    nonlinOpt.lbfgs_status() >= 0;
    // an empty intersection must make setup() return false, as the reference does
    std::vector<std::vector<double>> bad = {box(0, 0, 1.5, 1, 1, 1), box(100, 0, 1.5, 1, 1, 1)};
    frx::SE3GCOPTER o2;
    const bool rejected = !o2.setup(1000.0, 0.0, ini, fin, bad, INFINITY, 48, 0.5, 0.15, 0.08, 14.0, 5.0, 12.0, 3.8, 9.81, w, true);
    std::printf("infeasible corridor rejected: %d\n", (int)rejected);
    return (ok && rejected) ? 0 : 1;
}
