// Column bookkeeping shared by the solver kernels (solver.cu, solver_step.cu).
#pragma once
#include "backend.cuh"

namespace lfvio {

#ifdef __CUDACC__
__device__ __forceinline__ double clamp_diag(double v) { return fmin(fmax(v, 1e-6), 1e32); }  // min/max_lm_diagonal

// local column c (0..19) of a projection factor -> column of the full (reduced | landmark) system
__device__ __forceinline__ int proj_col(const StateLayout &lay, int frame_i, int frame_j, int landmark, int c) {
    if (c < 6) return lay.col_pose(frame_i) + c;
    if (c < 12) return lay.col_pose(frame_j) + (c - 6);
    if (c < 18) return lay.col_ex() + (c - 12);
    if (c == 18) return lay.D() + landmark;
    return lay.col_td();
}
__device__ __forceinline__ int imu_col(const StateLayout &lay, int i, int c) {
    if (c < 6) return lay.col_pose(i) + c;
    if (c < 15) return lay.col_sb(i) + (c - 6);
    if (c < 21) return lay.col_pose(i + 1) + (c - 15);
    return lay.col_sb(i + 1) + (c - 21);
}
#endif

}  // namespace lfvio
