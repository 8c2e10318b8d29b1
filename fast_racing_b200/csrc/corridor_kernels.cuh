// Launcher of the batched corridor step (corridor_kernels.cu).
#pragma once
#include <cuda_runtime.h>

namespace frx {
// scratch_idx: n_seg * n_obs ints; scratch_alive, scratch_in0: n_seg * n_obs bytes each.
cudaError_t launch_dilate_batch(const double* obs, int n_obs, const double* segments, int n_seg, const double* bbox, int* scratch_idx,
                                unsigned char* scratch_alive, unsigned char* scratch_in0, int* face_cnt, double* faces, double* ellipsoids,
                                cudaStream_t st);
}  // namespace frx
