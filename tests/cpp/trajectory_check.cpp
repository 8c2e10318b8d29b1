// CPU-only check of the header-only trajectory containers (include/se3gcopter_b200.hpp): exact max-rate queries against dense
// sampling, checkMax*Rate consistency, getRotation orthonormality / yaw convention, normalizePosCoeffMat, junction states.
// Prints "OK" or the first failing check.  Built by tests/test_facade_cpp.py with g++ (no CUDA, libfrx not linked).
#include <cstdio>
#include <cstdlib>
#include <random>

#include "se3gcopter_b200.hpp"

static int fail(const char* what, double a, double b) {
    std::printf("FAIL %s: %.17g vs %.17g\n", what, a, b);
    return 1;
}

int main() {
    std::mt19937_64 rng(12345);
    std::normal_distribution<double> N01(0.0, 1.0);
    frx::Trajectory traj;
    for (int piece = 0; piece < 40; piece++) {
        double c[18];
        for (int i = 0; i < 18; i++) c[i] = N01(rng) * (i < 6 ? 0.05 : 1.0);  // small high-order terms, like real pieces
        const double T = 0.3 + 2.0 * (rng() % 1000) / 1000.0;
        traj.emplace_back(T, c);
        const frx::Piece p(T, c);
        for (int deriv = 1; deriv <= 2; deriv++) {
            const double exact = deriv == 1 ? p.getMaxVelRate() : p.getMaxAccRate();
            double sampled = 0.0;
            const int S = 200000;
            for (int k = 0; k <= S; k++) {
                const auto v = deriv == 1 ? p.getVel(T * k / S) : p.getAcc(T * k / S);
                sampled = std::max(sampled, std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
            }
            if (exact < sampled * (1.0 - 1e-12)) return fail("exact max below a sample", exact, sampled);
            if (exact > sampled * (1.0 + 1e-6)) return fail("exact max not attained", exact, sampled);
            const bool below = deriv == 1 ? p.checkMaxVelRate(exact * 1.0001) : p.checkMaxAccRate(exact * 1.0001);
            const bool above = deriv == 1 ? p.checkMaxVelRate(exact * 0.9999) : p.checkMaxAccRate(exact * 0.9999);
            if (!below || above) return fail("checkMax*Rate", below, above);
        }
        // duration-normalised coefficients reproduce the position at the same fraction of the piece
        const auto n = p.normalizePosCoeffMat();
        const frx::Piece unit(1.0, n.data());
        for (double s : {0.0, 0.37, 1.0}) {
            const auto a = p.getPos(s * T), b = unit.getPos(s);
            for (int d = 0; d < 3; d++)
                if (std::fabs(a[d] - b[d]) > 1e-9 * (1.0 + std::fabs(a[d]))) return fail("normalizePosCoeffMat", a[d], b[d]);
        }
    }
    // constant-velocity piece: derivative of |v|^2 vanishes identically
    {
        double c[18] = {0};
        c[3 * 4 + 0] = 2.0, c[3 * 4 + 1] = -1.0;  // linear term (column 4 = power 1)
        const frx::Piece p(1.5, c);
        if (std::fabs(p.getMaxVelRate() - std::sqrt(5.0)) > 1e-12) return fail("constant velocity", p.getMaxVelRate(), std::sqrt(5.0));
        if (p.getMaxAccRate() != 0.0) return fail("zero acceleration", p.getMaxAccRate(), 0.0);
    }
    // trajectory-level maxima and junctions
    {
        double m = 0.0;
        for (int i = 0; i < traj.getPieceNum(); i++) m = std::max(m, traj[i].getMaxVelRate());
        if (m != traj.getMaxVelRate()) return fail("Trajectory::getMaxVelRate", m, traj.getMaxVelRate());
        if (!traj.checkMaxVelRate(m * 1.001) || traj.checkMaxVelRate(m * 0.999)) return fail("Trajectory::checkMaxVelRate", m, 0);
        const auto j0 = traj.getJuncPos(0), p0 = traj.getPos(0.0);
        const auto jN = traj.getJuncVel(traj.getPieceNum()), pN = traj[traj.getPieceNum() - 1].getVel(traj[traj.getPieceNum() - 1].getDuration());
        for (int d = 0; d < 3; d++)
            if (j0[d] != p0[d] || jN[d] != pN[d]) return fail("junction states", j0[d], p0[d]);
    }
    // getRotation: orthonormal, right-handed, zB along thrust, xB heading = yaw when level
    for (double t : {0.1, 3.3, 17.0}) {
        for (double yaw : {0.0, 0.7, 1.5707963267948966}) {
            double R[9];
            traj.getRotation(t, yaw, 9.81, R);
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) {
                    double dot = 0.0;
                    for (int d = 0; d < 3; d++) dot += R[3 * a + d] * R[3 * b + d];
                    if (std::fabs(dot - (a == b ? 1.0 : 0.0)) > 1e-12) return fail("getRotation orthonormal", dot, a == b);
                }
            auto acc = traj.getAcc(t);
            acc[2] += 9.81;
            const double n = std::sqrt(acc[0] * acc[0] + acc[1] * acc[1] + acc[2] * acc[2]);
            for (int d = 0; d < 3; d++)
                if (std::fabs(R[6 + d] - acc[d] / n) > 1e-12) return fail("getRotation zB", R[6 + d], acc[d] / n);
            const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[3] * (R[1] * R[8] - R[2] * R[7]) + R[6] * (R[1] * R[5] - R[2] * R[4]);
            if (std::fabs(det - 1.0) > 1e-12) return fail("getRotation right-handed", det, 1.0);
            // yB is orthogonal to the heading direction (cos yaw, sin yaw, 0)
            const double yh = R[3] * std::cos(yaw) + R[4] * std::sin(yaw);
            if (std::fabs(yh) > 1e-12) return fail("getRotation yaw", yh, 0.0);
        }
    }
    // wire format: normalised coefficients, highest order first, evaluate to the same positions (traj_server's convention:
    // p(t) = sum_j coef[j] * (t / T)^(order - j))
    {
        const frx::PolynomialTrajectoryFields m = frx::traj2fields(traj);
        if ((int)m.num_segment != traj.getPieceNum() || m.coef_x.size() != 6u * m.num_segment || m.num_order != 5) return fail("traj2fields sizes", m.num_segment, m.coef_x.size());
        for (unsigned i = 0; i < m.num_segment; i += 7) {
            const double T = m.time[i], s = 0.61;
            double px = 0.0;
            for (int j = 0; j <= 5; j++) px += m.coef_x[6 * i + j] * std::pow(s, 5 - j);
            const auto p = traj[(int)i].getPos(s * T);
            if (std::fabs(px - p[0]) > 1e-9 * (1.0 + std::fabs(px))) return fail("traj2fields coefficients", px, p[0]);
        }
    }
    std::printf("OK\n");
    return 0;
}
