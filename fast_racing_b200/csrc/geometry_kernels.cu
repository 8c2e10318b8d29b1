// Batched V-polytope extraction on the device: SE3GCOPTER::extractVs (se3gcopter_cpu.hpp:1031-1074 -> geoutils::enumerateVs,
// geoutils.hpp:135) for every polytope and every consecutive intersection of every corridor of a batch in ONE launch.
// One warp per polytope: the face planes sit in shared memory, the lanes share the face triples (a < b < c) of
// geometry_core.hpp, feasible intersections are collected in shared memory, ordered by their triple key, merged first-seen
// and checked (interior, boundedness) exactly as the host path does (host_geometry.cpp: finish_vertices) — this translation
// unit is compiled WITHOUT FMA contraction so that both paths produce bit-identical vertices in the same order.
// Work per polytope is O(F^3) plane triples x an early-exit feasibility scan: ~2 k triples for the 24-face intersections of
// the benchmark corridors, 162 k for two 50-face polytopes; HBM traffic is negligible (the planes in, <= 2F vertices out).
#include <cuda_runtime.h>

#include "geometry_core.hpp"
#include "geometry_kernels.cuh"
#include "waypoint_nls.hpp"

namespace frx {

#define GEO_WARPS 4

__global__ void __launch_bounds__(32 * GEO_WARPS)
    enumerate_vertices_kernel(const double* __restrict__ planes, const int2* __restrict__ polys, const int* __restrict__ vstart, int npoly,
                              double* __restrict__ verts, int* __restrict__ vcount) {
    __shared__ double snd[GEO_WARPS][4 * (geo::MAX_FACES + 1)];
    __shared__ double scand[GEO_WARPS][GEO_CAP][3];
    __shared__ int skey[GEO_WARPS][GEO_CAP];
    __shared__ int sord[GEO_WARPS][GEO_CAP];
    __shared__ int scnt[GEO_WARPS], sbad[GEO_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int poly = blockIdx.x * GEO_WARPS + warp;
    if (poly >= npoly) return;  // warp-uniform; only warp-level synchronisation below
    const int off = polys[poly].x, F = polys[poly].y;
    if (F < 4 || F > geo::MAX_FACES) {
        if (lane == 0) vcount[poly] = GEO_FALLBACK;
        return;
    }
    double* nd = snd[warp];
    for (int e = lane; e < 4 * F; e += 32) nd[e] = planes[4 * (size_t)off + e];
    if (lane == 0) scnt[warp] = 0, sbad[warp] = 0;
    __syncwarp();
    // ---- recession directions: +-(n_a x n_b) over all pairs ----
    const int npair = F * (F - 1) / 2;
    for (int p = lane; p < npair; p += 32) {
        int a = (int)((2.0 * F - 1.0 - sqrt((2.0 * F - 1.0) * (2.0 * F - 1.0) - 8.0 * p)) * 0.5);
        while (a > 0 && a * (2 * F - a - 1) / 2 > p) a--;
        while ((a + 1) * (2 * F - a - 2) / 2 <= p) a++;
        const int b = p - a * (2 * F - a - 1) / 2 + a + 1;
        if (geo::pair_ray(nd, F, a, b, 1.0) || geo::pair_ray(nd, F, a, b, -1.0)) sbad[warp] = 1;
    }
    // ---- vertices: for every first face a, the pairs (b, c) behind it are shared by the lanes ----
    for (int a = 0; a < F - 2; a++) {
        const int m = F - a - 1, np = m * (m - 1) / 2;
        for (int p = lane; p < np; p += 32) {
            int i = (int)((2.0 * m - 1.0 - sqrt((2.0 * m - 1.0) * (2.0 * m - 1.0) - 8.0 * p)) * 0.5);
            while (i > 0 && i * (2 * m - i - 1) / 2 > p) i--;
            while ((i + 1) * (2 * m - i - 2) / 2 <= p) i++;
            const int j = p - i * (2 * m - i - 1) / 2 + i + 1;
            const int b = a + 1 + i, c = a + 1 + j;
            double x[3];
            if (geo::triple_vertex(nd, F, a, b, c, x)) {
                const int idx = atomicAdd(&scnt[warp], 1);
                if (idx < GEO_CAP) {
                    scand[warp][idx][0] = x[0], scand[warp][idx][1] = x[1], scand[warp][idx][2] = x[2];
                    skey[warp][idx] = geo::triple_key(a, b, c);
                }
            }
        }
    }
    __syncwarp();
    const int n = scnt[warp];
    if (n > GEO_CAP) {  // more candidates than the staging area holds (heavily degenerate polytope): the host path takes over
        if (lane == 0) vcount[poly] = GEO_FALLBACK;
        return;
    }
    if (sbad[warp] || n < 4) {
        if (lane == 0) vcount[poly] = GEO_INFEASIBLE;
        return;
    }
    // ---- order by key (rank = number of smaller keys; keys are distinct), then the first-seen merge by one lane ----
    for (int i = lane; i < n; i += 32) {
        const int k = skey[warp][i];
        int r = 0;
        for (int j = 0; j < n; j++) r += skey[warp][j] < k;
        sord[warp][r] = i;
    }
    __syncwarp();
    if (lane == 0) {
        double* out = verts + 3 * (size_t)vstart[poly];
        const int cap = 2 * F;  // a simple 3-polytope with F faces has at most 2F - 4 vertices
        int V = 0;
        bool overflow = false;
        for (int r = 0; r < n && !overflow; r++) {
            const double* c = scand[warp][sord[warp][r]];
            bool dup = false;
            for (int v = 0; v < V && !dup; v++)
                dup = fabs(out[3 * v] - c[0]) < geo::MERGE_TOL && fabs(out[3 * v + 1] - c[1]) < geo::MERGE_TOL && fabs(out[3 * v + 2] - c[2]) < geo::MERGE_TOL;
            if (!dup) {
                if (V >= cap) {
                    overflow = true;
                    break;
                }
                out[3 * V] = c[0], out[3 * V + 1] = c[1], out[3 * V + 2] = c[2];
                V++;
            }
        }
        int res = V;
        if (overflow)
            res = GEO_FALLBACK;
        else if (V < 4)
            res = GEO_INFEASIBLE;
        else {
            double cen[3] = {0, 0, 0};
            for (int v = 0; v < V; v++)
                for (int d = 0; d < 3; d++) cen[d] += out[3 * v + d] / V;
            for (int k = 0; k < F; k++)
                if (!(nd[4 * k] * cen[0] + nd[4 * k + 1] * cen[1] + nd[4 * k + 2] * cen[2] - nd[4 * k + 3] < -geo::FEAS_TOL)) res = GEO_INFEASIBLE;
        }
        vcount[poly] = res;
    }
}

// backwardP of SE3GCOPTER::optimize's start point (se3gcopter_cpu.hpp:1240 -> :777-813), batched: one THREAD per inner
// waypoint runs the tiny L-BFGS of waypoint_nls.hpp on its own slice of a global workspace (M - 1 independent problems per
// trajectory, 20-odd variables each).  Same code as the host path, no FMA contraction: bit-identical start points.
__global__ void __launch_bounds__(128)
    waypoint_backward_kernel(int total_slots, const int* __restrict__ piece_traj, const int* __restrict__ x_off, const int* __restrict__ dimT,
                             const int* __restrict__ wp_xi_off, const int* __restrict__ wp_vert_off, const int* __restrict__ wp_vert_cnt,
                             const double* __restrict__ verts, const double* __restrict__ targets, const long long* __restrict__ work_off,
                             double* __restrict__ x, double* __restrict__ work) {
    const int gp = blockIdx.x * blockDim.x + threadIdx.x;
    if (gp >= total_slots) return;
    const int cnt = wp_vert_cnt[gp];
    if (cnt < 2) return;  // the last slot of a trajectory carries no waypoint
    const int b = piece_traj[gp];
    waypoint_backward(verts + 3 * (size_t)wp_vert_off[gp], cnt - 1, targets + 3 * (size_t)gp, x + x_off[b] + dimT[b] + wp_xi_off[gp],
                      work + work_off[gp]);
}

cudaError_t launch_waypoint_backward(int total_slots, const int* piece_traj, const int* x_off, const int* dimT, const int* wp_xi_off,
                                     const int* wp_vert_off, const int* wp_vert_cnt, const double* verts, const double* targets,
                                     const long long* work_off, double* x, double* work, cudaStream_t st) {
    if (total_slots <= 0) return cudaSuccess;
    waypoint_backward_kernel<<<(total_slots + 127) / 128, 128, 0, st>>>(total_slots, piece_traj, x_off, dimT, wp_xi_off, wp_vert_off, wp_vert_cnt,
                                                                      verts, targets, work_off, x, work);
    return cudaGetLastError();
}

cudaError_t launch_enumerate_vertices(const double* planes, const int2* polys, const int* vstart, int npoly, double* verts, int* vcount,
                                      cudaStream_t st) {
    if (npoly <= 0) return cudaSuccess;
    enumerate_vertices_kernel<<<(npoly + GEO_WARPS - 1) / GEO_WARPS, 32 * GEO_WARPS, 0, st>>>(planes, polys, vstart, npoly, verts, vcount);
    return cudaGetLastError();
}

}  // namespace frx
