// Vertex enumeration of the small H-polytopes of a flight corridor (replaces geoutils::findInterior + enumerateVs,
// geoutils.hpp:43-149: an interior-point LP, the polar dual and a quickhull), shared by the host path and the batched
// device kernel (geometry_kernels.cu) so that both produce the SAME vertices in the SAME order:
//   * a vertex is the intersection of three face planes that satisfies every other half-space (corridor polytopes have
//     6..100 faces: at most 161 700 triples, one warp-trip each on the device);
//   * candidates are taken in lexicographic order of the face triple (a < b < c) and de-duplicated first-seen
//     (several triples meet in one point wherever more than three faces do — always at the duplicated floor / ceiling
//     planes of an intersection polytope, MinCoPlan_CPU.cpp:85-91);
//   * the polytope is rejected when fewer than four vertices remain, when their centroid is not strictly interior
//     (empty interior, geoutils.hpp:59), or when a recession direction exists: an extreme ray of the cone {r : n_k.r <= 0}
//     is +-(n_a x n_b) for some pair of faces (unbounded polytope: the reference's LP then returns infinity).
// The vertex ORDER differs from the reference's (quickhull's triangle order); callers that need that order pass V-polytopes
// in through frx_problem::verts (tests/golden/*_refgeo.npz carry the reference's own).
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define GEO_HD __host__ __device__ __forceinline__
#else
#define GEO_HD inline
#endif

namespace frx {
namespace geo {

constexpr double DET_MIN = 1.0e-10;   // |n_a . (n_b x n_c)| below this: (nearly) dependent planes, no vertex
constexpr double FEAS_TOL = 1.0e-9;   // a point satisfies half-space k when n_k.x - d_k <= FEAS_TOL
constexpr double RAY_TOL = 1.0e-12;   // a direction is a recession direction when n_k.r <= RAY_TOL for every k
constexpr double MERGE_TOL = 1.0e-6;  // vertices closer than this (max norm) are one vertex (geoutils.hpp:135 epsilon)
constexpr int MAX_FACES = 127;        // key packing below: 7 bits per face index

// plane k: nd[4k .. 4k+3] = (nx, ny, nz, d), |n| = 1, half-space n.x <= d
GEO_HD bool triple_vertex(const double* nd, int F, int a, int b, int c, double* x) {
    const double *A = nd + 4 * a, *B = nd + 4 * b, *C = nd + 4 * c;
    const double bc0 = B[1] * C[2] - B[2] * C[1], bc1 = B[2] * C[0] - B[0] * C[2], bc2 = B[0] * C[1] - B[1] * C[0];
    const double det = A[0] * bc0 + A[1] * bc1 + A[2] * bc2;
    if (!(fabs(det) >= DET_MIN)) return false;
    const double ca0 = C[1] * A[2] - C[2] * A[1], ca1 = C[2] * A[0] - C[0] * A[2], ca2 = C[0] * A[1] - C[1] * A[0];
    const double ab0 = A[1] * B[2] - A[2] * B[1], ab1 = A[2] * B[0] - A[0] * B[2], ab2 = A[0] * B[1] - A[1] * B[0];
    const double x0 = (A[3] * bc0 + B[3] * ca0 + C[3] * ab0) / det;
    const double x1 = (A[3] * bc1 + B[3] * ca1 + C[3] * ab1) / det;
    const double x2 = (A[3] * bc2 + B[3] * ca2 + C[3] * ab2) / det;
    for (int k = 0; k < F; k++) {
        if (k == a || k == b || k == c) continue;
        const double* N = nd + 4 * k;
        if (N[0] * x0 + N[1] * x1 + N[2] * x2 - N[3] > FEAS_TOL) return false;
    }
    x[0] = x0, x[1] = x1, x[2] = x2;
    return true;
}

// Is sign * (n_a x n_b) a recession direction of the polytope?
GEO_HD bool pair_ray(const double* nd, int F, int a, int b, double sign) {
    const double *A = nd + 4 * a, *B = nd + 4 * b;
    double r0 = A[1] * B[2] - A[2] * B[1], r1 = A[2] * B[0] - A[0] * B[2], r2 = A[0] * B[1] - A[1] * B[0];
    const double len = sqrt(r0 * r0 + r1 * r1 + r2 * r2);
    if (!(len >= 1.0e-8)) return false;
    r0 *= sign / len, r1 *= sign / len, r2 *= sign / len;
    for (int k = 0; k < F; k++) {
        const double* N = nd + 4 * k;
        if (N[0] * r0 + N[1] * r1 + N[2] * r2 > RAY_TOL) return false;
    }
    return true;
}

GEO_HD int triple_key(int a, int b, int c) { return (a << 14) | (b << 7) | c; }  // lexicographic order of (a, b, c)

}  // namespace geo
}  // namespace frx
