// The corridor part of MavGlobalPlanner::plan (MinCoPlan_CPU.cpp:37-105) written against the façade exactly as the reference
// writes it against DecompROS: EllipsoidDecomp3D::set_obs / set_local_bbox / dilate / get_polyhedrons, Polyhedron3D::inside /
// add / hyperplanes, then hPolys (6 x F, one (normal; point) column per face).  CPU only (the segment step runs on the host here).
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "se3gcopter_b200.hpp"

int main() {
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::vector<frx::Vec3f> path, obs;
    for (int i = 0; i <= 60; i++) path.push_back({0.25 * i, 0.6 * std::sin(0.2 * i), 1.5});
    while (obs.size() < 20000) {  // random cloud with a free tube of radius 0.7 m around the path
        frx::Vec3f p = {-3.0 + 21.0 * U(rng), -5.0 + 10.0 * U(rng), 3.0 * U(rng)};
        double dmin = 1e9;
        for (const auto& q : path) dmin = std::min(dmin, std::sqrt((p[0] - q[0]) * (p[0] - q[0]) + (p[1] - q[1]) * (p[1] - q[1]) + (p[2] - q[2]) * (p[2] - q[2])));
        if (dmin > 0.7) obs.push_back(p);
    }
    frx::EllipsoidDecomp3D decomp_util;
    decomp_util.set_obs(obs);
    decomp_util.set_local_bbox({4.0, 4.0, 2.5});
    std::vector<frx::Polyhedron3D> decompPolys;
    for (int i = 0; i < (int)path.size() - 1;) {
        int k;
        for (k = i + 1; k < (int)path.size(); k++) {
            const double d = std::sqrt(std::pow(path[i][0] - path[k][0], 2) + std::pow(path[i][1] - path[k][1], 2) + std::pow(path[i][2] - path[k][2], 2));
            if (d >= 4.0) {
                k--;
                break;
            }
        }
        if (k < i + 1) k = i + 1;
        if (k >= (int)path.size()) k = (int)path.size() - 1;
        decomp_util.dilate({path[i], path[k]});
        frx::Polyhedron3D poly = decomp_util.get_polyhedrons()[0];
        decompPolys.push_back(poly);
        int j;
        for (j = k; j < (int)path.size(); j++)
            if (!poly.inside(path[j])) break;
        j--;
        if (j >= (int)path.size() - 1) break;
        i = (int)std::round((1 * i + 4 * j) / 5);
    }
    bool ok = decompPolys.size() >= 3;
    std::vector<std::vector<double>> hPolys;
    for (auto& p : decompPolys) {
        p.add(frx::Hyperplane3D({0.0, 0.0, 3.0}, {0.0, 0.0, 1.0}));
        p.add(frx::Hyperplane3D({0.0, 0.0, 0.0}, {0.0, 0.0, -1.0}));
        std::vector<double> h;
        for (const auto& v : p.hyperplanes()) h.insert(h.end(), {v.n_[0], v.n_[1], v.n_[2], v.p_[0], v.p_[1], v.p_[2]});
        hPolys.push_back(h);
        ok = ok && p.hyperplanes().size() > 8;
    }
    ok = ok && decompPolys.front().inside(path.front()) && decompPolys.back().inside(path.back());
    for (const auto& o : obs)  // no obstacle point is strictly inside any polytope
        for (const auto& p : decompPolys) {
            bool in = true;
            for (const auto& v : p.hyperplanes()) in = in && v.signed_dist(o) < -1e-9;
            ok = ok && !in;
        }
    std::printf("polytopes %zu, faces of the first %zu: %s\n", decompPolys.size(), hPolys[0].size() / 6, ok ? "OK" : "FAIL");
    return ok ? 0 : 1;
}
