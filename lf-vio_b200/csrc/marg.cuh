// Marginalisation plan (host-built index bookkeeping) and device buffers, shared by marg.cu and window.cu.
#pragma once
#include "backend.cuh"

namespace lfvio {

// Column bookkeeping of MarginalizationInfo::marginalize (marginalization_factor.cpp:174-205) in a canonical
// order: [dropped pose/speed-bias blocks (m_p columns) | kept blocks by ascending id (n columns)].  The dropped
// inverse depths are eliminated landmark by landmark and never get a column.
struct MargPlan {
    int flag;                       // LFVIO_MARGIN_OLD / LFVIO_MARGIN_SECOND_NEW
    int m_p, n, dim;                // dim = m_p + n
    int use_prior, imu_index;       // imu_index < 0: no IMU factor to marginalise
    int n_lm;                       // landmarks hosted in the dropped frame
    const int *lm_list;             // device: their indices
    int col[kMaxPriorBlocks];       // parameter-block id -> first column (-1: not involved)
    int n_keep;
    int keep_id[kMaxPriorBlocks];   // kept blocks in column order (ids BEFORE the addr_shift renaming)
    int keep_x0[kMaxPriorBlocks];   // offset of each kept block's data in x0_out
    int use_smem;                   // marg_finish_kernel's workspace fits in shared memory (decided on the handle's capacity)
    int g_items;                    // launch geometry: n_lm + 1 rounded up to a bucket (stable from frame to frame)
    int g_cap_n;                    // bound on n that sizes marg_finish_kernel's shared memory: 6 F + 10 (what the estimator's
                                    // own priors reach) or, for a larger hand-made prior, 15 F + 7
    const MargPlan *self_dev;       // device copy the kernels read
};
bool marg_workspace_in_smem(int cap_n);

struct MargBuffers {
    double *T;     // dim*dim + dim (system + right-hand side)
    double *work;  // global fallback workspace of marg_finish_kernel
    double *J0, *r0, *x0;
};

size_t marg_finish_workspace_doubles(int m_p, int n);
void marg_profile_read(long long *out);  // debug: phase boundaries (ns) of the last marg_finish_kernel
int launch_marginalize(const DevProblem &P, const EvalBuf &e0, const EvalBuf &e1, const double *x0, const double *x1,
                       const SolverCtl *ctl, MargPlan &plan, const MargBuffers &mb, cudaStream_t s);

}  // namespace lfvio
