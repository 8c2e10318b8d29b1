// ORACLE — TEST INFRASTRUCTURE ONLY.
// Compiles the reference's corridor decomposition VERBATIM, from where it lies under /root/reference:
//     sfc/DecompROS/decomp_ros_utils/include/decomp_util/{ellipsoid_decomp,line_segment,decomp_base}.h,
//     decomp_geometry/{ellipsoid,polyhedron,geometric_utils}.h, decomp_basis/{data_type,data_utils}.h
// over the Eigen-API shim in oracle/eigen_shim.  This is what MavGlobalPlanner::plan calls for every segment of the JPS path
// (MinCoPlan_CPU.cpp:37-63: set_obs, set_local_bbox, dilate, get_polyhedrons / get_ellipsoids).
// Built by oracle/Makefile into oracle/_ref/libref_decomp.so; no reference source is copied into this repository.
#include <cmath>
#include <vector>

#include <decomp_util/ellipsoid_decomp.h>

extern "C" {

// One segment p1 -> p2 against the obstacle cloud obs[3 n]: faces_out[6 cap] receives (n, p) per hyperplane in the order the
// reference builds them (obstacle-tangent planes, then the six local bounding-box planes); ellipsoid_out[12] = C (row-major
// 3x3) and d.  Returns the number of hyperplanes.
int ref_decomp_dilate(const double* obs, int n, const double* p1, const double* p2, const double* bbox, double* faces_out, int cap,
                      double* ellipsoid_out) {
    vec_Vec3f cloud(n);
    for (int i = 0; i < n; i++) cloud[i] = Vec3f(obs[3 * i], obs[3 * i + 1], obs[3 * i + 2]);
    EllipsoidDecomp3D decomp_util;
    decomp_util.set_obs(cloud);
    decomp_util.set_local_bbox(Vec3f(bbox[0], bbox[1], bbox[2]));
    vec_Vec3f line;
    line.push_back(Vec3f(p1[0], p1[1], p1[2]));
    line.push_back(Vec3f(p2[0], p2[1], p2[2]));
    decomp_util.dilate(line);
    const Polyhedron3D poly = decomp_util.get_polyhedrons()[0];
    const Ellipsoid3D ell = decomp_util.get_ellipsoids()[0];
    const auto hs = poly.hyperplanes();
    for (size_t k = 0; k < hs.size() && (int)k < cap; k++)
        for (int d = 0; d < 3; d++) faces_out[6 * k + d] = hs[k].n_(d), faces_out[6 * k + 3 + d] = hs[k].p_(d);
    if (ellipsoid_out) {
        const Mat3f C = ell.C();
        const Vec3f d = ell.d();
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) ellipsoid_out[3 * r + c] = C(r, c);
        for (int k = 0; k < 3; k++) ellipsoid_out[9 + k] = d(k);
    }
    return (int)hs.size();
}

// Polyhedron3D::inside (polyhedron.h: signed distance <= epsilon_ for every hyperplane)
int ref_polyhedron_inside(const double* faces, int F, const double* pt) {
    Polyhedron3D poly;
    for (int k = 0; k < F; k++)
        poly.add(Hyperplane3D(Vec3f(faces[6 * k + 3], faces[6 * k + 4], faces[6 * k + 5]), Vec3f(faces[6 * k], faces[6 * k + 1], faces[6 * k + 2])));
    return poly.inside(Vec3f(pt[0], pt[1], pt[2])) ? 1 : 0;
}
}
