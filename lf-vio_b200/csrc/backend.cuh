// Back-end (Estimator::optimization) device data structures and small fp64 math.
//
// Reference: vins_estimator/src/estimator.cpp:676-1009 and vins_estimator/src/factor/*.
// All back-end arithmetic is fp64 (the IMU square-root information comes from inverting a
// covariance with entries 1e-9..1e-4 and the marginalisation thresholds eigenvalues at 1e-8).
#pragma once
#include "common.cuh"

namespace lfvio {

constexpr int kMaxFrames = 32;       // F = WINDOW_SIZE + 1 (11 in the reference; config 4 uses 21)
constexpr int kMaxPriorBlocks = 2 * kMaxFrames + 2;
constexpr int kMaxObs = kMaxFrames;  // observations of one landmark

// ---- device view of one problem ------------------------------------------------------------------
struct StateLayout {
    int F, L;
    __host__ __device__ int pose(int f) const { return 7 * f; }
    __host__ __device__ int sb(int f) const { return 7 * F + 9 * f; }
    __host__ __device__ int ex() const { return 16 * F; }
    __host__ __device__ int td() const { return 16 * F + 7; }
    __host__ __device__ int feat(int l) const { return 16 * F + 8 + l; }
    __host__ __device__ int size() const { return 16 * F + 8 + L; }
    // local (tangent) columns of the reduced system, then the landmarks.  Elimination order of the dense solve:
    // [sb_{F-1} .. sb_1 | sb_0 | pose_{F-1} .. pose_0 | ex 6 | td 1] — speed-bias blocks first, newest first, so that
    // their fill stays inside a pose prefix (chol_blocked.cuh: WindowProfile).
    __host__ __device__ int D() const { return 15 * F + 7; }
    __host__ __device__ int col_pose(int f) const { return 9 * F + 6 * (F - 1 - f); }
    __host__ __device__ int col_sb(int f) const { return 9 * (F - 1 - f); }
    __host__ __device__ int col_ex() const { return 15 * F; }
    __host__ __device__ int col_td() const { return 15 * F + 6; }
    __host__ __device__ int block_col(int id) const {
        if (id < 2 * F) return (id & 1) ? col_sb(id >> 1) : col_pose(id >> 1);
        return id == 2 * F ? col_ex() : col_td();
    }
    __host__ __device__ int block_state(int id) const {
        if (id < 2 * F) return (id & 1) ? sb(id >> 1) : pose(id >> 1);
        return id == 2 * F ? ex() : td();
    }
    __host__ __device__ int block_gsize(int id) const { return id < 2 * F ? ((id & 1) ? 9 : 7) : (id == 2 * F ? 7 : 1); }
    __host__ __device__ int block_lsize(int id) const { return id < 2 * F ? ((id & 1) ? 9 : 6) : (id == 2 * F ? 6 : 1); }
};

// Projection factors built on the device by the FeatureManager (fm.cu): what the host needs to know about a list it never sees.
struct DevFactorSrc {
    int n_proj, n_landmarks;
    unsigned frame_j_mask;  // bit j: a landmark hosted in frame 0 is observed from frame j (marginalisation plan)
    int n_marg_lm;          // landmarks hosted in frame 0; their indices are already in the MARGIN_OLD landmark list
};
struct FmBuildDst {
    lfvio_proj_factor *proj;
    int *lm_first;
    double *para_feature;
    int *marg_lm;
    int proj_cap, lm_cap;
};
int fm_build_factors_into(lfvio_fm *f, const FmBuildDst &dst, DevFactorSrc *src);
int fm_set_depth_from_dev(lfvio_fm *f, const double *x_dev);
int fm_window_size(const lfvio_fm *f);
int fm_device(const lfvio_fm *f);
int launch_triangulate_slots(const double *Ps, const double *Rs, const double *tic, const double *ric, const int *start_frame,
                             const int *n_obs, const double *obs, int rec, int n_slots, int n_frames, double init_depth, double *depth,
                             cudaStream_t s);

struct DevPrior {
    int n, n_blocks;
    int block_id[kMaxPriorBlocks];
    int block_idx[kMaxPriorBlocks];
    int block_x0[kMaxPriorBlocks];  // offset of the block's data in x0
    const double *x0, *J0, *r0;
    const double *H0;  // J0^T J0 (n x n), g-independent, built once per solve
};

struct DevProblem {
    StateLayout lay;
    lfvio_solver_opts o;
    int n_proj, n_imu;
    const lfvio_proj_factor *proj;
    const lfvio_imu_factor *imu;
    const double *imu_sqrt_info;  // n_imu x 225
    const int *lm_first;          // L + 1: landmark l owns proj factors [lm_first[l], lm_first[l+1])
    DevPrior prior;
    // The kernels read this struct from DEVICE memory (self_dev), so that a launch carries no per-frame value and the
    // whole launch sequence of a solve can be replayed from a CUDA graph.  Launch geometry comes from the g_* fields:
    // the counts rounded up to coarse buckets (host side, window.cu: upload), stable from frame to frame; CTAs beyond the
    // true counts exit.
    const DevProblem *self_dev;
    int g_nbp;    // CTAs of 128 projection factors (evaluate_kernel, colnorm_kernel)
    int g_L;      // landmark items (accumulate_kernel)
    int g_imu;    // IMU items = F - 1
    int cap_L;    // landmark capacity of the handle (sizes of the per-solve clears)
    int cap_n;    // prior capacity of the handle (15 F + 7)
};

// residuals + local, loss-corrected Jacobians of every factor at one state
struct EvalBuf {
    double *rp;  // n_proj x 2
    double *Jp;  // n_proj x 2 x 20  [pose_i 6 | pose_j 6 | ex 6 | lambda | td]
    double *ri;  // n_imu x 15
    double *Ji;  // n_imu x 15 x 30  [pose_i 6 | sb_i 9 | pose_j 6 | sb_j 9]
    double *rm;  // prior n
};

// Trust-region / dogleg control block, lives in device memory (no host round trip inside a solve).
struct SolverCtl {
    int cur;             // which EvalBuf / state buffer is current (0/1)
    int reuse;           // DoglegStrategy::reuse_
    int done;            // 0 running, 1 converged, 2 max iterations, -1 failure
    int iter;            // trust-region iterations performed
    int eval_pending;    // the step kernel produced a candidate that must be evaluated
    int step_valid;
    int successful;
    int invalid_in_a_row;
    unsigned int ticket; // last-block-done counter of the eval kernel
    int pad;
    double radius, mu, cost, cost_cand, x_norm, dogleg_step_norm, alpha, model_cost_change, step_norm;
    double gnorm, gn_norm, gdot;  // |gradient_|, |gauss_newton_step_|, gradient_ . gauss_newton_step_ (kept while reuse_)
    double jj, mu_solve;          // |J_s (gradient_/diagonal_)|^2 and the mu of the last linear solve (kept while reuse_)
    // per-iteration accumulators filled with atomics by the multi-CTA phases and consumed/zeroed by the phase that
    // needs the total: [0] sum gradient^2  [1] sum (J g)^2  [2] max |J^T r| (as ordered bits)  [3] sum gn^2
    // [4] sum gradient.gn  [5] model cost change  [6] |x - x_cand|^2
    double acc[8];
    unsigned int ticket2;  // last-block-done counter of cand_kernel
    int pad2;
    double cost_log[LFVIO_MAX_ITER_LOG], radius_log[LFVIO_MAX_ITER_LOG];
    int accepted_log[LFVIO_MAX_ITER_LOG];
    long long phase_clk[16];  // clock64() at the phase boundaries of the last step_kernel (LFVIO_DEBUG_STEP=1 prints them)
};

#ifdef __CUDACC__
// ---- small fp64 math -----------------------------------------------------------------------------
struct D3 {
    double x, y, z;
};
__device__ __forceinline__ D3 d3(double x, double y, double z) { return D3{x, y, z}; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return D3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return D3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 operator-(D3 a) { return D3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ D3 operator*(double s, D3 a) { return D3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return D3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ D3 operator/(D3 a, double s) { return D3{a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ double dot3(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 cross3(D3 a, D3 b) { return D3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double norm3(D3 a) { return sqrt(dot3(a, a)); }
__device__ __forceinline__ D3 unit3(D3 a) { return a / norm3(a); }
__device__ __forceinline__ double get3(const D3 &a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

struct DM3 {
    double m[3][3];
};
__device__ __forceinline__ DM3 dm3_zero() {
    DM3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = 0.0;
    return r;
}
__device__ __forceinline__ DM3 dm3_identity() {
    DM3 r = dm3_zero();
    r.m[0][0] = r.m[1][1] = r.m[2][2] = 1.0;
    return r;
}
__device__ __forceinline__ DM3 operator*(const DM3 &a, const DM3 &b) {
    DM3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
__device__ __forceinline__ D3 operator*(const DM3 &a, D3 v) {
    return D3{a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
              a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
__device__ __forceinline__ DM3 operator*(double s, const DM3 &a) {
    DM3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = s * a.m[i][j];
    return r;
}
__device__ __forceinline__ DM3 operator+(const DM3 &a, const DM3 &b) {
    DM3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] + b.m[i][j];
    return r;
}
__device__ __forceinline__ DM3 operator-(const DM3 &a, const DM3 &b) {
    DM3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] - b.m[i][j];
    return r;
}
__device__ __forceinline__ DM3 operator-(const DM3 &a) { return (-1.0) * a; }
__device__ __forceinline__ DM3 transpose3(const DM3 &a) {
    DM3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i];
    return r;
}
// Utility::skewSymmetric (utility/utility.h:30-38)
__device__ __forceinline__ DM3 skew3(D3 q) {
    DM3 r = dm3_zero();
    r.m[0][1] = -q.z;
    r.m[0][2] = q.y;
    r.m[1][0] = q.z;
    r.m[1][2] = -q.x;
    r.m[2][0] = -q.y;
    r.m[2][1] = q.x;
    return r;
}

struct DQ {  // Hamilton quaternion
    double w, x, y, z;
};
__device__ __forceinline__ DQ dq(double w, double x, double y, double z) { return DQ{w, x, y, z}; }
__device__ __forceinline__ D3 qvec(const DQ &q) { return D3{q.x, q.y, q.z}; }
__device__ __forceinline__ DQ operator*(const DQ &a, const DQ &b) {
    return DQ{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ DQ qinv(const DQ &q) {  // Eigen::Quaternion::inverse(): conjugate / squaredNorm
    const double n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
    return DQ{q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
}
__device__ __forceinline__ DQ qnormalized(const DQ &q) {
    const double n = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    return DQ{q.w / n, q.x / n, q.y / n, q.z / n};
}
__device__ __forceinline__ DM3 qrot(const DQ &q) {  // Eigen::Quaternion::toRotationMatrix()
    DM3 r;
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    r.m[0][0] = 1 - (tyy + tzz);
    r.m[0][1] = txy - twz;
    r.m[0][2] = txz + twy;
    r.m[1][0] = txy + twz;
    r.m[1][1] = 1 - (txx + tzz);
    r.m[1][2] = tyz - twx;
    r.m[2][0] = txz - twy;
    r.m[2][1] = tyz + twx;
    r.m[2][2] = 1 - (txx + tyy);
    return r;
}
__device__ __forceinline__ D3 qapply(const DQ &q, D3 v) {  // Eigen: q * v
    const D3 u = qvec(q);
    D3 uv = cross3(u, v);
    uv = uv + uv;
    return v + q.w * uv + cross3(u, uv);
}
// Utility::deltaQ (utility/utility.h:16-28): first order, not normalised
__device__ __forceinline__ DQ delta_q(D3 th) { return DQ{1.0, th.x / 2.0, th.y / 2.0, th.z / 2.0}; }
// bottom-right 3x3 of Utility::Qleft / Qright (utility/utility.h:47-64)
__device__ __forceinline__ DM3 qleft_br(const DQ &q) { return q.w * dm3_identity() + skew3(qvec(q)); }
__device__ __forceinline__ DM3 qright_br(const DQ &q) { return q.w * dm3_identity() - skew3(qvec(q)); }

__device__ __forceinline__ void load_pose(const double *x, D3 *p, DQ *q) {
    *p = D3{x[0], x[1], x[2]};
    *q = DQ{x[6], x[3], x[4], x[5]};
}
// PoseLocalParameterization::Plus (factor/pose_local_parameterization.cpp:3-19)
__device__ __forceinline__ void pose_plus(const double *x, const double *d, double *out) {
    out[0] = x[0] + d[0];
    out[1] = x[1] + d[1];
    out[2] = x[2] + d[2];
    const DQ nq = qnormalized(dq(x[6], x[3], x[4], x[5]) * delta_q(D3{d[3], d[4], d[5]}));
    out[3] = nq.x;
    out[4] = nq.y;
    out[5] = nq.z;
    out[6] = nq.w;
}
#endif  // __CUDACC__

// ---- launchers (window.cu drives them) -----------------------------------------------------------
struct SolverScratch {
    double *S;         // D x D reduced system (full storage, symmetric; upper triangle accumulated)
    double *rhs;       // D
    double *diag2;     // D : squared column norms of the scaled Jacobian (reduced columns)
    double *grad;      // D : J_s^T r (un-normalised)
    double *lm_c;      // L : sum j_lambda^2 (scaled)
    double *lm_g;      // L : j_lambda^T r
    double *lm_ws;     // L x 13 : [pose_i 6 | ex 6 | td] part of w_l
    double *proj_wj;   // n_proj x 6 : pose_j part of w_l
    int *lm_coli;      // L : reduced column of the host pose block (pose_i) of landmark l, -1 if it has no factor
    int *proj_colj;    // n_proj : reduced column of the pose_j block of factor k
    double *scale;     // D + L Jacobi scaling
    double *diagonal;  // D + L
    double *gradient;  // D + L (divided by diagonal)
    double *gn;        // D + L Gauss-Newton step (scaled space)
    double *step;      // D + L
    double *y;         // D work vector
    size_t zero_bytes; // bytes from S that must be cleared before every accumulation (S .. lm_g)
};

int launch_imu_prepare(const DevProblem &P, double *imu_sqrt_info, int *status, cudaStream_t s);
int launch_prior_prepare(const DevProblem &P, double *H0, cudaStream_t s);
int launch_evaluate(const DevProblem &P, const double *x0, const double *x1, const EvalBuf &e0, const EvalBuf &e1,
                    SolverCtl *ctl, int mode, int robust, cudaStream_t s);
int launch_jacobi_scale(const DevProblem &P, const EvalBuf &e0, const SolverScratch &w, cudaStream_t s);
int launch_accumulate(const DevProblem &P, const EvalBuf &e0, const EvalBuf &e1, const SolverScratch &w, SolverCtl *ctl,
                      cudaStream_t s, bool clear_done = false);
// true when launch_step's chol_kernel clears [S | rhs | diag2 | grad] itself while consuming them (fused step, S in smem)
bool step_zeroes_system(const DevProblem &P);
// One trust-region step up to (not including) the candidate evaluation: pre -> chol -> gnlm -> cand.
#ifdef LFVIO_CHOL_PROFILE
void chol_profile_read(long long *out);
#endif
int launch_step(const DevProblem &P, const EvalBuf &e0, const EvalBuf &e1, const SolverScratch &w, double *x0, double *x1,
                SolverCtl *ctl, cudaStream_t s);

}  // namespace lfvio
