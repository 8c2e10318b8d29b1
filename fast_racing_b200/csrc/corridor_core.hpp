// Safe-flight-corridor generation around one path segment — the reference's LineSegment<3>::dilate
// (sfc/DecompROS/decomp_ros_utils/include/decomp_util/line_segment.h:33-37 -> find_ellipsoid :128-208, DecompBase::find_polyhedron
// decomp_base.h:60-79, add_local_bbox line_segment.h:49-85) with Ellipsoid / Hyperplane / Polyhedron of decomp_geometry/
// {ellipsoid,polyhedron}.h and vec3_to_rotation of geometric_utils.h:27-35, restated on plain arrays.  Shared by the host path
// (corridor.cpp) and the batched device kernel (corridor_kernels.cu): the scalar algebra lives here, the scans over the obstacle
// cloud (points inside, closest point, filters) are plain loops on the host and CTA-wide reductions on the device.
// Arithmetic follows the reference's expression order (the 3x3 inverse by cofactors as Eigen does for fixed size 3).
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define COR_HD __host__ __device__ __forceinline__
#else
#define COR_HD inline
#endif

namespace frx {
namespace cor {

constexpr double EPS = 1e-10;  // decomp_basis/data_type.h:127 epsilon_

struct Ellipsoid {
    double C[9];   // row-major 3x3
    double Ci[9];  // C^-1 (recomputed whenever C changes; the reference inverts on every dist() call: same value)
    double d[3];
};

COR_HD void mat3_mul(const double* A, const double* B, double* O) {  // O = A B, left-to-right sums
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) O[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}
COR_HD void mat3_mul_t(const double* A, const double* B, double* O) {  // O = A B^T
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) O[3 * r + c] = A[3 * r] * B[3 * c] + A[3 * r + 1] * B[3 * c + 1] + A[3 * r + 2] * B[3 * c + 2];
}
COR_HD void mat3_vec(const double* A, const double* v, double* o) {
    for (int r = 0; r < 3; r++) o[r] = A[3 * r] * v[0] + A[3 * r + 1] * v[1] + A[3 * r + 2] * v[2];
}
COR_HD void mat3_tvec(const double* A, const double* v, double* o) {  // A^T v
    for (int r = 0; r < 3; r++) o[r] = A[r] * v[0] + A[3 + r] * v[1] + A[6 + r] * v[2];
}
COR_HD void mat3_inverse(const double* m, double* r) {  // cofactors of the first column, determinant from them, then the rest
#define M_(i, j) m[3 * (i) + (j)]
#define COF_(i, j) (M_(((i) + 1) % 3, ((j) + 1) % 3) * M_(((i) + 2) % 3, ((j) + 2) % 3) - M_(((i) + 1) % 3, ((j) + 2) % 3) * M_(((i) + 2) % 3, ((j) + 1) % 3))
    const double c00 = COF_(0, 0), c10 = COF_(1, 0), c20 = COF_(2, 0);
    const double det = c00 * M_(0, 0) + c10 * M_(1, 0) + c20 * M_(2, 0);
    const double invdet = 1.0 / det;
    r[0] = c00 * invdet, r[1] = c10 * invdet, r[2] = c20 * invdet;
    r[3] = COF_(0, 1) * invdet, r[4] = COF_(1, 1) * invdet, r[5] = COF_(2, 1) * invdet;
    r[6] = COF_(0, 2) * invdet, r[7] = COF_(1, 2) * invdet, r[8] = COF_(2, 2) * invdet;
#undef COF_
#undef M_
}
// rotation matrix of the unit quaternion (w, x, y, z), Eigen's QuaternionBase::toRotationMatrix
COR_HD void quat_to_mat(double w, double x, double y, double z, double* R) {
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0 - (tyy + tzz), R[1] = txy - twz, R[2] = txz + twy;
    R[3] = txy + twz, R[4] = 1.0 - (txx + tzz), R[5] = tyz - twx;
    R[6] = txz - twy, R[7] = tyz + twx, R[8] = 1.0 - (txx + tyy);
}
COR_HD void quat_mul(const double* a, const double* b, double* o) {  // (w, x, y, z)
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
    o[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
}
// geometric_utils.h:27-35: zero roll, pitch = atan2(-v_z, |v_xy|), yaw = atan2(v_y, v_x); R = Mat3f(qz * qy * qx)
COR_HD void vec3_to_rotation(const double* v, double* R) {
    const double rpy0 = 0.0, rpy1 = atan2(-v[2], sqrt(v[0] * v[0] + v[1] * v[1])), rpy2 = atan2(v[1], v[0]);
    const double qx[4] = {cos(rpy0 / 2), sin(rpy0 / 2), 0, 0}, qy[4] = {cos(rpy1 / 2), 0, sin(rpy1 / 2), 0}, qz[4] = {cos(rpy2 / 2), 0, 0, sin(rpy2 / 2)};
    double qzy[4], q[4];
    quat_mul(qz, qy, qzy);
    quat_mul(qzy, qx, q);
    quat_to_mat(q[0], q[1], q[2], q[3], R);
}
COR_HD void set_C(Ellipsoid& E, const double* C) {
    for (int k = 0; k < 9; k++) E.C[k] = C[k];
    mat3_inverse(E.C, E.Ci);
}
// E.C_ = R diag(a) R^T  ((R * new_C) * R^T as the reference writes it)
COR_HD void set_C_axes(Ellipsoid& E, const double* R, double a0, double a1, double a2) {
    const double D[9] = {a0, 0, 0, 0, a1, 0, 0, 0, a2};
    double RD[9], C[9];
    mat3_mul(R, D, RD);
    mat3_mul_t(RD, R, C);
    set_C(E, C);
}
// Ellipsoid::dist (ellipsoid.h:19-21): |C^-1 (pt - d)|
COR_HD double ell_dist(const Ellipsoid& E, const double* pt) {
    const double v[3] = {pt[0] - E.d[0], pt[1] - E.d[1], pt[2] - E.d[2]};
    double u[3];
    mat3_vec(E.Ci, v, u);
    return sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
}
// Ellipsoid::closest_hyperplane (ellipsoid.h:53-58): n = (C^-1 C^-T)(pt - d), normalised; the plane passes through pt
COR_HD void ell_hyperplane(const Ellipsoid& E, const double* pt, double* n) {
    double CC[9];
    mat3_mul_t(E.Ci, E.Ci, CC);
    const double v[3] = {pt[0] - E.d[0], pt[1] - E.d[1], pt[2] - E.d[2]};
    double u[3];
    mat3_vec(CC, v, u);
    const double nn = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    n[0] = u[0] / nn, n[1] = u[1] / nn, n[2] = u[2] / nn;
}
// line_segment.h:49-85: the six planes of the local bounding box, (n, p) each, in the reference's order
COR_HD void local_bbox(const double* p1, const double* p2, const double* bbox, double* faces /* 36 */) {
    double dir[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    const double dn = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    for (int k = 0; k < 3; k++) dir[k] /= dn;
    double dh[3] = {dir[1], -dir[0], 0.0};
    if (sqrt(dh[0] * dh[0] + dh[1] * dh[1] + dh[2] * dh[2]) == 0) dh[0] = -1, dh[1] = 0, dh[2] = 0;
    const double hn = sqrt(dh[0] * dh[0] + dh[1] * dh[1] + dh[2] * dh[2]);
    for (int k = 0; k < 3; k++) dh[k] /= hn;
    const double dv[3] = {dir[1] * dh[2] - dir[2] * dh[1], dir[2] * dh[0] - dir[0] * dh[2], dir[0] * dh[1] - dir[1] * dh[0]};
    for (int k = 0; k < 3; k++) {
        faces[0 + k] = dh[k], faces[3 + k] = p1[k] + dh[k] * bbox[1];
        faces[6 + k] = -dh[k], faces[9 + k] = p1[k] - dh[k] * bbox[1];
        faces[12 + k] = dir[k], faces[15 + k] = p2[k] + dir[k] * bbox[0];
        faces[18 + k] = -dir[k], faces[21 + k] = p1[k] - dir[k] * bbox[0];
        faces[24 + k] = dv[k], faces[27 + k] = p1[k] + dv[k] * bbox[2];
        faces[30 + k] = -dv[k], faces[33 + k] = p1[k] - dv[k] * bbox[2];
    }
}
// Hyperplane::signed_dist (polyhedron.h): n . (pt - p)
COR_HD double signed_dist(const double* face, const double* pt) {
    return face[0] * (pt[0] - face[3]) + face[1] * (pt[1] - face[4]) + face[2] * (pt[2] - face[5]);
}
// Polyhedron::inside (polyhedron.h:54-62)
COR_HD bool inside_faces(const double* faces, int F, const double* pt) {
    for (int k = 0; k < F; k++)
        if (signed_dist(faces + 6 * k, pt) > EPS) return false;
    return true;
}

// The scalar steps of find_ellipsoid between two scans (line_segment.h:128-208), as a small state machine so that host and
// device run the same code around their own scans.
struct DilateState {
    Ellipsoid E;
    double Ri[9], Rf[9];
    double axes[3];
    double f;
};
COR_HD void dilate_begin(DilateState& S, const double* p1, const double* p2, double offset_x) {
    const double dx[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    S.f = sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]) / 2;
    double C[9] = {S.f, 0, 0, 0, S.f, 0, 0, 0, S.f};
    S.axes[0] = S.axes[1] = S.axes[2] = S.f;
    C[0] += offset_x;
    S.axes[0] += offset_x;
    if (S.axes[0] > 0) {
        const double ratio = S.axes[1] / S.axes[0];
        for (int k = 0; k < 3; k++) S.axes[k] *= ratio;
        for (int k = 0; k < 9; k++) C[k] *= ratio;
    }
    const double v[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    vec3_to_rotation(v, S.Ri);
    double RC[9], CC[9];
    mat3_mul(S.Ri, C, RC);
    mat3_mul_t(RC, S.Ri, CC);
    set_C(S.E, CC);
    for (int k = 0; k < 3; k++) S.E.d[k] = (p1[k] + p2[k]) / 2;
    for (int k = 0; k < 9; k++) S.Rf[k] = S.Ri[k];
}
// first loop body, given the closest point pw (line_segment.h:150-165)
COR_HD void dilate_shrink_yz(DilateState& S, const double* pw) {
    const double v[3] = {pw[0] - S.E.d[0], pw[1] - S.E.d[1], pw[2] - S.E.d[2]};
    double p[3];
    mat3_tvec(S.Ri, v, p);
    const double roll = atan2(p[2], p[1]);
    double Q[9];
    quat_to_mat(cos(roll / 2), sin(roll / 2), 0, 0, Q);
    mat3_mul(S.Ri, Q, S.Rf);
    mat3_tvec(S.Rf, v, p);
    if (p[0] < S.axes[0]) S.axes[1] = fabs(p[1]) / sqrt(1 - pow(p[0] / S.axes[0], 2));
    set_C_axes(S.E, S.Rf, S.axes[0], S.axes[1], S.axes[1]);
}
// between the loops (line_segment.h:175-181): reset the ellipsoid with the old axes(2)
COR_HD void dilate_reset_z(DilateState& S) { set_C_axes(S.E, S.Rf, S.axes[0], S.axes[1], S.axes[2]); }
// second loop body (line_segment.h:183-204)
COR_HD void dilate_shrink_z(DilateState& S, const double* pw) {
    const double v[3] = {pw[0] - S.E.d[0], pw[1] - S.E.d[1], pw[2] - S.E.d[2]};
    double p[3];
    mat3_tvec(S.Rf, v, p);
    const double dd = 1 - pow(p[0] / S.axes[0], 2) - pow(p[1] / S.axes[1], 2);
    if (dd > EPS) S.axes[2] = fabs(p[2]) / sqrt(dd);
    set_C_axes(S.E, S.Rf, S.axes[0], S.axes[1], S.axes[2]);
}

}  // namespace cor
}  // namespace frx
