// Batched corridor step on the device: LineSegment3D::dilate (decomp_util/line_segment.h:33-37) for MANY
// segments against one obstacle cloud resident in HBM, one CTA per segment.
//   phase 0  local bounding box filter over the whole cloud (DecompBase::set_obs, decomp_base.h:33-38): coalesced scan,
//            survivors appended to this segment's slice of a global index buffer;
//   phase 1  find_ellipsoid: the two shrink loops of line_segment.h:128-208 — per iteration a CTA-wide arg-min of the
//            ellipsoid distance over the live points (ties -> lowest cloud index = the reference's first-in-order rule),
//            the scalar update by thread 0 (corridor_core.hpp), a parallel filter;
//   phase 2  find_polyhedron (decomp_base.h:60-79): arg-min, tangent plane, filter by signed distance, until no point is left;
//   phase 3  the six box planes.
// The scans are HBM/L2-bound integer-and-fp64 work (24 B per point and pass); the loops are sequential in the iterations
// (10-40 per segment), parallel over the points and over the segments.  No FMA contraction (compiled --fmad=false) so that
// the point SELECTION matches the host path except for last-bit differences of the device's sin/cos/atan2.
#include <cuda_runtime.h>

#include "../../include/frx.h"
#include "corridor_core.hpp"
#include "corridor_kernels.cuh"

namespace frx {

using namespace cor;
#define COR_THREADS 256

struct ArgMin {
    double d;
    int i;
};
__device__ __forceinline__ ArgMin better(ArgMin a, ArgMin b) { return (b.d < a.d || (b.d == a.d && b.i < a.i)) ? b : a; }

// CTA-wide arg-min of the ellipsoid distance over the live entries of list[0..n)
__device__ int block_closest(const Ellipsoid& E, const double* __restrict__ obs, const int* list, const unsigned char* alive, int n, ArgMin* red) {
    ArgMin m{1.7976931348623157e308, 0x7fffffff};
    for (int e = threadIdx.x; e < n; e += COR_THREADS)
        if (alive[e]) {
            const int i = list[e];
            m = better(m, ArgMin{ell_dist(E, obs + 3 * (size_t)i), i});
        }
    for (int o = 16; o > 0; o >>= 1) {
        ArgMin t{__shfl_xor_sync(0xffffffffu, m.d, o), __shfl_xor_sync(0xffffffffu, m.i, o)};
        m = better(m, t);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    ArgMin r = red[0];
    for (int w = 1; w < COR_THREADS / 32; w++) r = better(r, red[w]);
    return r.i;  // 0x7fffffff: nothing alive
}

__global__ void __launch_bounds__(COR_THREADS)
    dilate_batch_kernel(const double* __restrict__ obs, int n_obs, const double* __restrict__ segments, int n_seg, double bx, double by, double bz,
                        int* __restrict__ scratch_idx, unsigned char* __restrict__ scratch_alive, unsigned char* __restrict__ scratch_in0,
                        int* __restrict__ face_cnt, double* __restrict__ faces, double* __restrict__ ellipsoids) {
    const int s = blockIdx.x;
    if (s >= n_seg) return;
    __shared__ DilateState S;
    __shared__ double box[36], cur_face[6];
    __shared__ ArgMin red[COR_THREADS / 32];
    __shared__ int sh_n, sh_cnt, sh_any;
    const int tid = threadIdx.x;
    const double* p1 = segments + 6 * (size_t)s;
    const double* p2 = p1 + 3;
    int* list = scratch_idx + (size_t)s * n_obs;
    unsigned char* alive = scratch_alive + (size_t)s * n_obs;
    unsigned char* in0 = scratch_in0 + (size_t)s * n_obs;
    double* out = faces + (size_t)s * FRX_CORRIDOR_MAX_FACES * 6;
    const double bbox[3] = {bx, by, bz};
    const bool have_box = sqrt(bx * bx + by * by + bz * bz) != 0;
    if (tid == 0) {
        if (have_box) local_bbox(p1, p2, bbox, box);
        dilate_begin(S, p1, p2, 0.0);
        sh_n = 0, sh_cnt = 0;
    }
    __syncthreads();
    // ---- phase 0: points inside the local bounding box (order does not matter: every later choice is by (distance, cloud index)) ----
    for (int i0 = 0; i0 < n_obs; i0 += COR_THREADS) {
        const int i = i0 + tid;
        const bool in = i < n_obs && inside_faces(box, have_box ? 6 : 0, obs + 3 * (size_t)i);
        const unsigned m = __ballot_sync(0xffffffffu, in);
        int base = 0;
        if ((tid & 31) == 0 && m) base = atomicAdd(&sh_n, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (in) list[base + __popc(m & ((1u << (tid & 31)) - 1u))] = i;
    }
    __syncthreads();
    const int n = sh_n;
    // ---- phase 1a: points inside the initial ellipsoid; shrink the two short axes together ----
    for (int e = tid; e < n; e += COR_THREADS) {
        const unsigned char a = ell_dist(S.E, obs + 3 * (size_t)list[e]) <= 1;
        alive[e] = a, in0[e] = a;
    }
    __syncthreads();
    for (;;) {
        const int c = block_closest(S.E, obs, list, alive, n, red);
        if (c == 0x7fffffff) break;
        if (tid == 0) dilate_shrink_yz(S, obs + 3 * (size_t)c);
        __syncthreads();
        for (int e = tid; e < n; e += COR_THREADS)
            if (alive[e]) alive[e] = 1 - ell_dist(S.E, obs + 3 * (size_t)list[e]) > EPS;
        __syncthreads();
    }
    // ---- phase 1b: reset with the old third axis, shrink it ----
    if (tid == 0) dilate_reset_z(S);
    __syncthreads();
    for (int e = tid; e < n; e += COR_THREADS) alive[e] = in0[e] && ell_dist(S.E, obs + 3 * (size_t)list[e]) <= 1;
    __syncthreads();
    for (;;) {
        const int c = block_closest(S.E, obs, list, alive, n, red);
        if (c == 0x7fffffff) break;
        if (tid == 0) dilate_shrink_z(S, obs + 3 * (size_t)c);
        __syncthreads();
        for (int e = tid; e < n; e += COR_THREADS)
            if (alive[e]) alive[e] = 1 - ell_dist(S.E, obs + 3 * (size_t)list[e]) > EPS;
        __syncthreads();
    }
    // ---- phase 2: polyhedron ----
    for (int e = tid; e < n; e += COR_THREADS) alive[e] = 1;
    __syncthreads();
    for (;;) {
        const int c = block_closest(S.E, obs, list, alive, n, red);
        if (c == 0x7fffffff) break;
        if (tid == 0) {
            const double* pt = obs + 3 * (size_t)c;
            ell_hyperplane(S.E, pt, cur_face);
            cur_face[3] = pt[0], cur_face[4] = pt[1], cur_face[5] = pt[2];
            if (sh_cnt < FRX_CORRIDOR_MAX_FACES - 6)
                for (int k = 0; k < 6; k++) out[6 * sh_cnt + k] = cur_face[k];
            sh_cnt++;
        }
        __syncthreads();
        for (int e = tid; e < n; e += COR_THREADS)
            if (alive[e]) alive[e] = signed_dist(cur_face, obs + 3 * (size_t)list[e]) < 0;
        __syncthreads();
    }
    if (tid == 0) {
        int cnt = sh_cnt;
        if (cnt > FRX_CORRIDOR_MAX_FACES - 6)
            cnt = -1;  // more tangent planes than the output holds
        else if (have_box) {
            for (int k = 0; k < 36; k++) out[6 * cnt + k] = box[k];
            cnt += 6;
        }
        face_cnt[s] = cnt;
        if (ellipsoids) {
            for (int k = 0; k < 9; k++) ellipsoids[12 * (size_t)s + k] = S.E.C[k];
            for (int k = 0; k < 3; k++) ellipsoids[12 * (size_t)s + 9 + k] = S.E.d[k];
        }
    }
}

cudaError_t launch_dilate_batch(const double* obs, int n_obs, const double* segments, int n_seg, const double* bbox, int* scratch_idx,
                                unsigned char* scratch_alive, unsigned char* scratch_in0, int* face_cnt, double* faces, double* ellipsoids,
                                cudaStream_t st) {
    if (n_seg <= 0) return cudaSuccess;
    dilate_batch_kernel<<<n_seg, COR_THREADS, 0, st>>>(obs, n_obs, segments, n_seg, bbox[0], bbox[1], bbox[2], scratch_idx, scratch_alive, scratch_in0,
                                                       face_cnt, faces, ellipsoids);
    return cudaGetLastError();
}

}  // namespace frx
