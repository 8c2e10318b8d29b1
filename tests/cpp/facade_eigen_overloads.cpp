// The caller of the reference, VERBATIM in shape (MavGlobalPlanner::plan, MinCoPlan_CPU.cpp:93-124): Eigen::MatrixXd hPolys,
// Eigen::Matrix3d states, Eigen::Vector4d weights, SE3GCOPTER::setup(...) / optimize(traj, tol) with the reference's exact
// argument list — compiled against the façade's Eigen overloads (enabled by __has_include(<Eigen/Eigen>)).  The image has no
// Eigen: the test builds this with -I oracle/eigen_shim standing in for it (data(), size(), operator(), Matrix3d from
// MatrixXd are all the overloads need).  Runs on the GPU box; exit code 0 = same trajectory as the plain-array call.
#include <cmath>
#include <cstdio>
#include <vector>

#include "se3gcopter_b200.hpp"

#ifndef FRX_HAVE_EIGEN
#error "the Eigen overloads are not enabled: <Eigen/Eigen> was not found on the include path"
#endif

int main() {
    const double n[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    const double h[6] = {3.0, 3.0, 2.5, 2.5, 1.4, 1.4};
    std::vector<Eigen::MatrixXd> hPolys;
    std::vector<std::vector<double>> flat;
    for (int i = 0; i < 4; i++) {
        Eigen::MatrixXd hPoly(6, 6);  // one face per column: (outer normal; point), MinCoPlan_CPU.cpp:101
        std::vector<double> f;
        for (int k = 0; k < 6; k++) {
            const double c[3] = {3.0 * i + 1.5, 0.4 * i, 1.5};
            for (int d = 0; d < 3; d++) hPoly(d, k) = n[k][d], hPoly(3 + d, k) = c[d] + n[k][d] * h[k];
            for (int d = 0; d < 6; d++) f.push_back(hPoly(d, k));
        }
        hPolys.push_back(hPoly);
        flat.push_back(f);
    }
    Eigen::Matrix3d iniState, finState;
    iniState.setZero(), finState.setZero();
    iniState(2, 0) = 1.5;
    finState(0, 0) = 12.0, finState(1, 0) = 1.2, finState(2, 0) = 1.5;
    Eigen::Vector4d w;
    w(0) = 1.0e7, w(1) = 1.0e4, w(2) = 1.0e4, w(3) = 1.0e4;
    frx::SE3GCOPTER nonlinOpt;
    frx::Trajectory traj;
    if (!nonlinOpt.setup(1000.0, 0.0, iniState, finState, hPolys, INFINITY, 48, 0.5, 0.15, 0.08, 14.0, 5.0, 12.0, 3.8, 9.81, w, true)) return 2;
    const double obj = nonlinOpt.optimize(traj, 1.0e-6);
    // the same problem through the plain-array overload
    const double ini[9] = {0, 0, 1.5, 0, 0, 0, 0, 0, 0}, fin[9] = {12.0, 1.2, 1.5, 0, 0, 0, 0, 0, 0}, wa[4] = {1.0e7, 1.0e4, 1.0e4, 1.0e4};
    frx::SE3GCOPTER plain;
    frx::Trajectory traj2;
    if (!plain.setup(1000.0, 0.0, ini, fin, flat, INFINITY, 48, 0.5, 0.15, 0.08, 14.0, 5.0, 12.0, 3.8, 9.81, wa, true)) return 3;
    const double obj2 = plain.optimize(traj2, 1.0e-6);
    std::printf("Eigen overload: cost %.9f pieces %d | plain arrays: cost %.9f\n", obj, traj.getPieceNum(), obj2);
    return (obj == obj2 && traj.getPieceNum() == 4 && traj.getTotalDuration() == traj2.getTotalDuration()) ? 0 : 1;
}
