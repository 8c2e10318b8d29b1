// ORACLE — TEST INFRASTRUCTURE ONLY.
// Compiles the reference's H-polytope -> V-polytope conversion VERBATIM, from where it lies under /root/reference:
//     geoutils.hpp (findInterior, enumerateVs, filterVs), sdlp.hpp (Seidel's LP), quickhull.hpp
// over the Eigen-API shim in oracle/eigen_shim.  This is the code SE3GCOPTER::extractVs runs (se3gcopter_cpu.hpp:1031-1074);
// its vertex ORDER (quickhull's triangle order + first-seen dedup) defines the order of the xi
// variables.  Built by oracle/Makefile into oracle/_ref/libref_geo.so; no reference source is copied into this repository.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include <Eigen/Eigen>

#include "geoutils.hpp"  // -I /root/reference/src/plan_manage/include/se3gcopter

extern "C" {

// One H-polytope: faces = F rows (nx,ny,nz,px,py,pz), normals are normalised here as SE3GCOPTER::setup does
// (se3gcopter_cpu.hpp:1116).  Returns the number of vertices V (or -1 when the polytope has no interior) and writes up to
// `cap` vertices (absolute coordinates, in the reference's order) to out[3*V].
int ref_enumerate_vertices(const double* faces, int F, double* out, int cap) {
    Eigen::MatrixXd h(6, F);
    for (int k = 0; k < F; k++) {
        const double* f = faces + 6 * (size_t)k;
        const double nn = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
        for (int d = 0; d < 3; d++) h(d, k) = f[d] / nn, h(3 + d, k) = f[3 + d];
    }
    Eigen::MatrixXd v;
    if (!geoutils::enumerateVs(h, v)) return -1;
    const int V = (int)v.cols();
    for (int c = 0; c < V && c < cap; c++)
        for (int d = 0; d < 3; d++) out[3 * c + d] = v(d, c);
    return V;
}

// The interior point of the LP alone (geoutils::findInterior): returns 1 and the point, or 0.
int ref_find_interior(const double* faces, int F, double* interior) {
    Eigen::MatrixXd h(6, F);
    for (int k = 0; k < F; k++) {
        const double* f = faces + 6 * (size_t)k;
        const double nn = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
        for (int d = 0; d < 3; d++) h(d, k) = f[d] / nn, h(3 + d, k) = f[3 + d];
    }
    Eigen::Vector3d p;
    const bool ok = geoutils::findInterior(h, p);
    for (int d = 0; d < 3; d++) interior[d] = p(d);
    return ok ? 1 : 0;
}
}
