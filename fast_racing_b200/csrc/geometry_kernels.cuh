// Launcher of the batched V-polytope extraction (geometry_kernels.cu).
#pragma once
#include <cuda_runtime.h>

namespace frx {

#define GEO_CAP 192          // candidate intersections staged per polytope
#define GEO_FALLBACK (-1)    // vcount: the kernel declines this polytope (too many faces / candidates): host path
#define GEO_INFEASIBLE (-2)  // vcount: empty interior or unbounded (SE3GCOPTER::setup returns false, se3gcopter_cpu.hpp:1118-1121)

// planes: 4 doubles per face (unit normal, offset n.p); polys[k] = (first face, face count); vstart[k] = first output vertex
// slot of polytope k (2 * F_k slots each); verts: 3 doubles per slot; vcount[k] = vertices written, or GEO_*.
cudaError_t launch_enumerate_vertices(const double* planes, const int2* polys, const int* vstart, int npoly, double* verts, int* vcount,
                                      cudaStream_t st);

// Per-waypoint start-point solves (backwardP): slot gp = global piece index of the waypoint's LEFT piece; targets: 3 per slot;
// work_off[gp]: first double of the slot's workspace (nls_work_doubles(k) each); x: packed variables of the batch (device).
cudaError_t launch_waypoint_backward(int total_slots, const int* piece_traj, const int* x_off, const int* dimT, const int* wp_xi_off,
                                     const int* wp_vert_off, const int* wp_vert_cnt, const double* verts, const double* targets,
                                     const long long* work_off, double* x, double* work, cudaStream_t st);

}  // namespace frx
