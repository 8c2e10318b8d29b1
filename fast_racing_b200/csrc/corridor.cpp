// Host side of the corridor generation: LineSegment3D::dilate for one segment (sequential scans over the
// obstacle cloud around the shared algebra of corridor_core.hpp) and the walk of MavGlobalPlanner::plan along the front-end
// path (MinCoPlan_CPU.cpp:37-91).  The batched device form of the segment step is corridor_kernels.cu.
#include "corridor.hpp"

#include <cmath>
#include <cstring>
#include <limits>

namespace frx {

using namespace cor;

// Ellipsoid::closest_point (ellipsoid.h:38-50): the FIRST point of minimal dist in list order
static int closest(const Ellipsoid& E, const double* obs, const std::vector<int>& idx) {
    int best = -1;
    double min_dist = std::numeric_limits<double>::max();
    for (int i : idx) {
        const double d = ell_dist(E, obs + 3 * (size_t)i);
        if (d < min_dist) min_dist = d, best = i;
    }
    return best;
}

int dilate_segment(const double* obs, int n_obs, const double* p1, const double* p2, const double* bbox, std::vector<double>& faces,
                   double* ellipsoid12) {
    faces.clear();
    double box[36];
    const bool have_box = std::sqrt(bbox[0] * bbox[0] + bbox[1] * bbox[1] + bbox[2] * bbox[2]) != 0;
    if (have_box) local_bbox(p1, p2, bbox, box);
    // DecompBase::set_obs (decomp_base.h:33-38): only the points inside the local bounding box, in cloud order
    std::vector<int> obs_;
    for (int i = 0; i < n_obs; i++)
        if (inside_faces(box, have_box ? 6 : 0, obs + 3 * (size_t)i)) obs_.push_back(i);
    // ---- find_ellipsoid (line_segment.h:128-208) ----
    DilateState S;
    dilate_begin(S, p1, p2, 0.0);
    std::vector<int> inside0, cur, nxt;
    for (int i : obs_)
        if (ell_dist(S.E, obs + 3 * (size_t)i) <= 1) inside0.push_back(i);
    cur = inside0;
    while (!cur.empty()) {
        dilate_shrink_yz(S, obs + 3 * (size_t)closest(S.E, obs, cur));
        nxt.clear();
        for (int i : cur)
            if (1 - ell_dist(S.E, obs + 3 * (size_t)i) > EPS) nxt.push_back(i);
        cur.swap(nxt);
    }
    dilate_reset_z(S);
    cur.clear();
    for (int i : inside0)
        if (ell_dist(S.E, obs + 3 * (size_t)i) <= 1) cur.push_back(i);
    while (!cur.empty()) {
        dilate_shrink_z(S, obs + 3 * (size_t)closest(S.E, obs, cur));
        nxt.clear();
        for (int i : cur)
            if (1 - ell_dist(S.E, obs + 3 * (size_t)i) > EPS) nxt.push_back(i);
        cur.swap(nxt);
    }
    // ---- find_polyhedron (decomp_base.h:60-79) ----
    cur = obs_;
    while (!cur.empty()) {
        const double* pt = obs + 3 * (size_t)closest(S.E, obs, cur);
        double f[6];
        ell_hyperplane(S.E, pt, f);
        f[3] = pt[0], f[4] = pt[1], f[5] = pt[2];
        faces.insert(faces.end(), f, f + 6);
        nxt.clear();
        for (int i : cur)
            if (signed_dist(f, obs + 3 * (size_t)i) < 0) nxt.push_back(i);
        cur.swap(nxt);
    }
    if (have_box) faces.insert(faces.end(), box, box + 36);  // add_local_bbox(polyhedron_)
    if (ellipsoid12) {
        std::memcpy(ellipsoid12, S.E.C, sizeof(double) * 9);
        std::memcpy(ellipsoid12 + 9, S.E.d, sizeof(double) * 3);
    }
    return (int)faces.size() / 6;
}

// One step of the walk (MinCoPlan_CPU.cpp:44-54): the farthest path point k > i that is neither blocked from path[i] nor 4 m away
int corridor_walk_target(const double* path, int n_path, int i, frx_is_blocked_t blocked, void* ctx, double max_len) {
    int k;
    for (k = i + 1; k < n_path; k++) {
        const double* a = path + 3 * (size_t)i;
        const double* b = path + 3 * (size_t)k;
        const double d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
        if ((blocked && blocked(ctx, a, b)) || std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) >= max_len) {
            k--;
            break;
        }
    }
    if (k < i + 1) k = i + 1;
    if (k >= n_path) k = n_path - 1;
    return k;
}
// MinCoPlan_CPU.cpp:63-79: the last path point (from k on) still inside the new polytope, then the 1:4 advance rule.
// Returns the next i, or -1 when the corridor reached the end of the path.
int corridor_walk_advance(const double* path, int n_path, int i, int k, const double* faces, int F) {
    int j;
    for (j = k; j < n_path; j++)
        if (!inside_faces(faces, F, path + 3 * (size_t)j)) break;
    j--;
    if (j >= n_path - 1) return -1;
    return (int)std::round((1 * i + 4 * j) / 5);  // integer division first, as the reference writes it
}

int generate_corridor(const double* obs, int n_obs, const double* path, int n_path, const double* bbox, double map_height,
                      frx_is_blocked_t blocked, void* ctx, std::vector<int>& face_off, std::vector<double>& faces) {
    face_off.assign(1, 0);
    faces.clear();
    if (n_path < 2) return 0;
    std::vector<double> f;
    for (int i = 0; i < n_path - 1;) {
        const int k = corridor_walk_target(path, n_path, i, blocked, ctx, 4.0);
        const int F = dilate_segment(obs, n_obs, path + 3 * (size_t)i, path + 3 * (size_t)k, bbox, f, nullptr);
        // ceiling and floor (MinCoPlan_CPU.cpp:85-91) are appended after the walk in the reference: the advance rule sees F faces
        const int next = corridor_walk_advance(path, n_path, i, k, f.data(), F);
        const double ceil_floor[12] = {0, 0, 1, 0, 0, map_height, 0, 0, -1, 0, 0, 0};
        f.insert(f.end(), ceil_floor, ceil_floor + 12);
        faces.insert(faces.end(), f.begin(), f.end());
        face_off.push_back(face_off.back() + F + 2);
        if (next < 0) break;
        i = next > i ? next : i + 1;  // the reference's advance rule can stall (wp == i, an endless loop there): move on
    }
    return (int)face_off.size() - 1;
}

}  // namespace frx
