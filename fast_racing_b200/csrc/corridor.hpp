// Corridor generation, host entry points (corridor.cpp) — see corridor_core.hpp for the reference citations.
#pragma once
#include <vector>

#include "../../include/frx.h"
#include "corridor_core.hpp"

namespace frx {

// LineSegment3D::dilate(0): faces = (n, p) rows, obstacle-tangent planes in the order they are found, then the six local
// bounding-box planes; ellipsoid12 (optional) = C row-major, d.  Returns the number of faces.
int dilate_segment(const double* obs, int n_obs, const double* p1, const double* p2, const double* bbox, std::vector<double>& faces,
                   double* ellipsoid12);
int corridor_walk_target(const double* path, int n_path, int i, frx_is_blocked_t blocked, void* ctx, double max_len);
int corridor_walk_advance(const double* path, int n_path, int i, int k, const double* faces, int F);
// The whole walk of MavGlobalPlanner::plan (MinCoPlan_CPU.cpp:37-91); returns the number of polytopes.
int generate_corridor(const double* obs, int n_obs, const double* path, int n_path, const double* bbox, double map_height,
                      frx_is_blocked_t blocked, void* ctx, std::vector<int>& face_off, std::vector<double>& faces);

}  // namespace frx
