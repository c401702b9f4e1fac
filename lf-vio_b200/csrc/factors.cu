// K7-K10 — batched factor evaluation: residuals + analytic local Jacobians of every factor of the window.
//
// Reference (ceres::CostFunction::Evaluate implementations):
//   ProjectionTdFactor::Evaluate   vins_estimator/src/factor/projection_td_factor.cpp:36-151 (ctor :8-34)
//   ProjectionFactor::Evaluate     vins_estimator/src/factor/projection_factor.cpp:21-121   (ctor :6-19)
//   IMUFactor::Evaluate            vins_estimator/src/factor/imu_factor.h:19-200
//   IntegrationBase::evaluate      vins_estimator/src/factor/integration_base.h:160-186
//   MarginalizationFactor::Evaluate vins_estimator/src/factor/marginalization_factor.cpp:333-381
// plus what Ceres does around them: the local parameterisation (PoseLocalParameterization's Jacobian is
// [I6; 0] => the first 6 columns of each 7-wide pose block) and the Corrector for CauchyLoss(1.0)
// (rho'' < 0 => residual and Jacobian are both scaled by sqrt(rho'), cf. marginalization_factor.cpp:37-68).
//
// One kernel evaluates everything at one state: projection factors one thread each (coalesced AoS reads of
// 144 B factor records, 42 doubles out), IMU factors one CTA each (raw 15x30 Jacobian assembled by 3x3
// blocks in shared memory, then the 15x15 square-root information is applied by the whole CTA), the prior
// one CTA (dense n x n mat-vec).  The square-root information is hoisted out of Evaluate: the covariance is
// constant during a solve, so it is factorised once per solve by imu_prepare_kernel.
// In candidate mode the last CTA to finish runs the trust-region accept/reject bookkeeping
// (ceres TrustRegionMinimizer + DoglegStrategy::StepAccepted/StepRejected/StepIsInvalid) on device.
#include "backend.cuh"

namespace lfvio {

namespace {

constexpr int O_P = 0, O_R = 3, O_V = 6, O_BA = 9, O_BG = 12;  // vins_estimator/src/parameters.h:50-57
constexpr int EV_THREADS = 128;

// Two threads per factor (part 0: residual, cost and the two pose blocks of the Jacobian, columns 0-11; part 1: extrinsic,
// inverse depth and td, columns 12-19): one thread per factor was the kernel's critical path (8.5 of 12.8 us, ~1000 dependent-ish
// fp64 instructions at one every 4.5 cycles per warp); the shared prelude is computed by both.
__device__ void eval_projection(const lfvio_proj_factor &f, const double *x, const StateLayout &lay,
                                const lfvio_solver_opts &o, int robust, double *r_out, double *J_out, double *cost, int part) {
    D3 Pi, Pj, tic;
    DQ Qi, Qj, qic;
    load_pose(x + lay.pose(f.frame_i), &Pi, &Qi);
    load_pose(x + lay.pose(f.frame_j), &Pj, &Qj);
    load_pose(x + lay.ex(), &tic, &qic);
    const double inv_dep_i = x[lay.feat(f.landmark)];
    const double td = x[lay.td()];
    const D3 pts_i = d3(f.pts_i[0], f.pts_i[1], f.pts_i[2]), pts_j = d3(f.pts_j[0], f.pts_j[1], f.pts_j[2]);
    const D3 vel_i = d3(f.vel_i[0], f.vel_i[1], f.vel_i[2]), vel_j = d3(f.vel_j[0], f.vel_j[1], f.vel_j[2]);
    // tangent basis of the un-shifted pts_j (ctor)
    const D3 a = unit3(pts_j);
    D3 tmp = d3(0, 0, 1);
    if (a.x == tmp.x && a.y == tmp.y && a.z == tmp.z) tmp = d3(1, 0, 0);
    const D3 b1 = unit3(tmp - a * dot3(a, tmp));
    const D3 b2 = cross3(a, b1);
    D3 pts_i_td = pts_i, pts_j_td = pts_j;
    if (o.estimate_td) {
        // ROW is a double in the reference (vins_estimator/src/parameters.h:39): ROW / 2 is a floating division
        const double row_i = f.row_i - (double)o.row / 2, row_j = f.row_j - (double)o.row / 2;
        pts_i_td = pts_i - (td - f.td_i + o.tr / o.row * row_i) * vel_i;
        pts_j_td = pts_j - (td - f.td_j + o.tr / o.row * row_j) * vel_j;
    }
    const D3 pts_camera_i = pts_i_td / inv_dep_i;
    const D3 pts_imu_i = qapply(qic, pts_camera_i) + tic;
    const D3 pts_w = qapply(Qi, pts_imu_i) + Pi;
    const D3 pts_imu_j = qapply(qinv(Qj), pts_w - Pj);
    const D3 pts_camera_j = qapply(qinv(qic), pts_imu_j - tic);
    const double sq = o.focal_length / 1.5;
    const D3 d = unit3(pts_camera_j) - unit3(pts_j_td);
    double r0 = sq * dot3(b1, d), r1 = sq * dot3(b2, d);
    const double sq_norm = r0 * r0 + r1 * r1;
    double sr = 1.0;
    if (robust) {
        const double b = o.cauchy_a * o.cauchy_a, c = 1.0 / b;
        const double sum = 1.0 + sq_norm * c;
        *cost = 0.5 * b * log(sum);
        sr = sqrt(1.0 / sum);
    } else {
        *cost = 0.5 * sq_norm;
    }
    if (part == 0) {
        r_out[0] = r0 * sr;
        r_out[1] = r1 * sr;
    } else {
        *cost = 0.0;
    }
    if (!J_out) return;
    const DM3 Ri = qrot(Qi), Rj = qrot(Qj), ric = qrot(qic);
    const double n = norm3(pts_camera_j);
    const double x1 = pts_camera_j.x, x2 = pts_camera_j.y, x3 = pts_camera_j.z;
    // norm_jaco = I / n - x x^T / n^3 (pow(norm, 3) and nine divisions in the original): one reciprocal, the rest products
    const double inv_n = 1.0 / n, inv_n3 = inv_n * inv_n * inv_n;
    DM3 nj;
    nj.m[0][0] = inv_n - x1 * x1 * inv_n3; nj.m[0][1] = -x1 * x2 * inv_n3;        nj.m[0][2] = -x1 * x3 * inv_n3;
    nj.m[1][0] = nj.m[0][1];               nj.m[1][1] = inv_n - x2 * x2 * inv_n3; nj.m[1][2] = -x2 * x3 * inv_n3;
    nj.m[2][0] = nj.m[0][2];               nj.m[2][1] = nj.m[1][2];               nj.m[2][2] = inv_n - x3 * x3 * inv_n3;
    double red[2][3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        red[0][c] = sr * sq * (b1.x * nj.m[0][c] + b1.y * nj.m[1][c] + b1.z * nj.m[2][c]);
        red[1][c] = sr * sq * (b2.x * nj.m[0][c] + b2.y * nj.m[1][c] + b2.z * nj.m[2][c]);
    }
    auto put = [&](int col0, const DM3 &A) {
#pragma unroll
        for (int rr = 0; rr < 2; rr++)
#pragma unroll
            for (int c = 0; c < 3; c++)
                J_out[rr * 20 + col0 + c] = red[rr][0] * A.m[0][c] + red[rr][1] * A.m[1][c] + red[rr][2] * A.m[2][c];
    };
    const DM3 ricT = transpose3(ric), RjT = transpose3(Rj);
    const DM3 A = ricT * RjT;
    if (part == 0) {
        put(0, A);
        put(3, A * Ri * (-skew3(pts_imu_i)));
        put(6, ricT * (-RjT));
        put(9, ricT * skew3(pts_imu_j));
        return;
    }
    const DM3 tmp_r = A * Ri * ric;
    if (o.estimate_extrinsic) {
        put(12, ricT * (RjT * Ri - dm3_identity()));
        put(15, (-tmp_r) * skew3(pts_camera_i) + skew3(tmp_r * pts_camera_i) +
                    skew3(ricT * (RjT * (Ri * tic + Pi - Pj) - tic)));
    } else {  // SetParameterBlockConstant(para_Ex_Pose): the block drops out of the program
#pragma unroll
        for (int c = 12; c < 18; c++) J_out[c] = J_out[20 + c] = 0.0;
    }
    const D3 dl = (tmp_r * pts_i_td) * (-1.0 / (inv_dep_i * inv_dep_i));
    const D3 dt = (tmp_r * vel_i) * (-1.0 / inv_dep_i);
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
        J_out[rr * 20 + 18] = red[rr][0] * dl.x + red[rr][1] * dl.y + red[rr][2] * dl.z;
        double jt = 0.0;
        if (o.estimate_td)  // reduce * R * vel_i / lambda * -1 + sqrt_info * velocity_j.head(2), as the reference has it
            jt = red[rr][0] * dt.x + red[rr][1] * dt.y + red[rr][2] * dt.z + sr * sq * (rr == 0 ? vel_j.x : vel_j.y);
        J_out[rr * 20 + 19] = jt;
    }
}

__device__ __forceinline__ DM3 jac_block(const double *jac, int r0, int c0) {
    DM3 m;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) m.m[i][j] = jac[(r0 + i) * 15 + c0 + j];
    return m;
}
__device__ __forceinline__ void put_block(double (*Jr)[30], int r0, int c0, const DM3 &A) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) Jr[r0 + i][c0 + j] = A.m[i][j];
}
// (Qleft(ql) * Qright(qr)).bottomRightCorner<3,3>()
__device__ DM3 qleft_qright_br(const DQ &ql, const DQ &qr) {
    double L[4][4], R[4][4];
    L[0][0] = ql.w; L[0][1] = -ql.x; L[0][2] = -ql.y; L[0][3] = -ql.z;
    L[1][0] = ql.x; L[2][0] = ql.y; L[3][0] = ql.z;
    R[0][0] = qr.w; R[0][1] = -qr.x; R[0][2] = -qr.y; R[0][3] = -qr.z;
    R[1][0] = qr.x; R[2][0] = qr.y; R[3][0] = qr.z;
    const DM3 lb = qleft_br(ql), rb = qright_br(qr);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            L[i + 1][j + 1] = lb.m[i][j];
            R[i + 1][j + 1] = rb.m[i][j];
        }
    DM3 out;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += L[i + 1][k] * R[k][j + 1];
            out.m[i][j] = s;
        }
    return out;
}

// One CTA per IMU factor.
__device__ void eval_imu_cta(const lfvio_imu_factor &f, const double *sqrt_info, const double *x, const StateLayout &lay,
                             const lfvio_solver_opts &o, double *r_out, double *J_out, double *cost_out,
                             double (*s_Jr)[30], double *s_raw, double *s_sq, double *s_res) {
    const int tid = threadIdx.x;
    for (int k = tid; k < 225; k += blockDim.x) s_sq[k] = sqrt_info[k];
    for (int k = tid; k < 450; k += blockDim.x) (&s_Jr[0][0])[k] = 0.0;
    __syncthreads();
    if (tid == 0) {
        const int i = f.frame_i, j = i + 1;
        D3 Pi, Pj;
        DQ Qi, Qj;
        load_pose(x + lay.pose(i), &Pi, &Qi);
        load_pose(x + lay.pose(j), &Pj, &Qj);
        const double *sbi = x + lay.sb(i), *sbj = x + lay.sb(j);
        const D3 Vi = d3(sbi[0], sbi[1], sbi[2]), Bai = d3(sbi[3], sbi[4], sbi[5]), Bgi = d3(sbi[6], sbi[7], sbi[8]);
        const D3 Vj = d3(sbj[0], sbj[1], sbj[2]), Baj = d3(sbj[3], sbj[4], sbj[5]), Bgj = d3(sbj[6], sbj[7], sbj[8]);
        const D3 G = d3(o.g[0], o.g[1], o.g[2]);
        const double sum_dt = f.sum_dt;
        const DM3 dp_dba = jac_block(f.jacobian, O_P, O_BA), dp_dbg = jac_block(f.jacobian, O_P, O_BG);
        const DM3 dq_dbg = jac_block(f.jacobian, O_R, O_BG);
        const DM3 dv_dba = jac_block(f.jacobian, O_V, O_BA), dv_dbg = jac_block(f.jacobian, O_V, O_BG);
        const D3 dba = Bai - d3(f.linearized_ba[0], f.linearized_ba[1], f.linearized_ba[2]);
        const D3 dbg = Bgi - d3(f.linearized_bg[0], f.linearized_bg[1], f.linearized_bg[2]);
        const DQ dQ = dq(f.delta_q[0], f.delta_q[1], f.delta_q[2], f.delta_q[3]);
        const D3 delta_p = d3(f.delta_p[0], f.delta_p[1], f.delta_p[2]), delta_v = d3(f.delta_v[0], f.delta_v[1], f.delta_v[2]);
        const DQ corrected_delta_q = dQ * delta_q(dq_dbg * dbg);
        const D3 corrected_delta_v = delta_v + dv_dba * dba + dv_dbg * dbg;
        const D3 corrected_delta_p = delta_p + dp_dba * dba + dp_dbg * dbg;
        const DQ Qi_inv = qinv(Qi);
        const D3 ap = qapply(Qi_inv, 0.5 * G * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt);
        const D3 av = qapply(Qi_inv, G * sum_dt + Vj - Vi);
        const D3 rp = ap - corrected_delta_p;
        const D3 rq = 2.0 * qvec(qinv(corrected_delta_q) * (Qi_inv * Qj));
        const D3 rv = av - corrected_delta_v;
        const D3 rba = Baj - Bai, rbg = Bgj - Bgi;
        for (int k = 0; k < 3; k++) {
            s_raw[O_P + k] = get3(rp, k);
            s_raw[O_R + k] = get3(rq, k);
            s_raw[O_V + k] = get3(rv, k);
            s_raw[O_BA + k] = get3(rba, k);
            s_raw[O_BG + k] = get3(rbg, k);
        }
        if (J_out) {
            const DM3 RiT = qrot(Qi_inv);
            put_block(s_Jr, O_P, 0, -RiT);
            put_block(s_Jr, O_P, 3, skew3(ap));
            put_block(s_Jr, O_R, 3, -qleft_qright_br(qinv(Qj) * Qi, corrected_delta_q));
            put_block(s_Jr, O_V, 3, skew3(av));
            put_block(s_Jr, O_P, 6, (-sum_dt) * RiT);
            put_block(s_Jr, O_P, 9, -dp_dba);
            put_block(s_Jr, O_P, 12, -dp_dbg);
            put_block(s_Jr, O_R, 12, -(qleft_br(qinv(Qj) * Qi * dQ) * dq_dbg));
            put_block(s_Jr, O_V, 6, -RiT);
            put_block(s_Jr, O_V, 9, -dv_dba);
            put_block(s_Jr, O_V, 12, -dv_dbg);
            put_block(s_Jr, O_BA, 9, -dm3_identity());
            put_block(s_Jr, O_BG, 12, -dm3_identity());
            put_block(s_Jr, O_P, 15, RiT);
            put_block(s_Jr, O_R, 18, qleft_br(qinv(corrected_delta_q) * Qi_inv * Qj));
            put_block(s_Jr, O_V, 21, RiT);
            put_block(s_Jr, O_BA, 24, dm3_identity());
            put_block(s_Jr, O_BG, 27, dm3_identity());
        }
    }
    __syncthreads();
    if (tid < 15) {
        double s = 0;
        for (int k = 0; k < 15; k++) s += s_sq[tid * 15 + k] * s_raw[k];
        r_out[tid] = s;
        s_res[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
        double c = 0;
        for (int k = 0; k < 15; k++) c += 0.5 * s_res[k] * s_res[k];
        *cost_out = c;
    }
    if (J_out) {
        for (int e = tid; e < 450; e += blockDim.x) {
            const int i = e / 30, c = e % 30;
            double s = 0;
#pragma unroll
            for (int k = 0; k < 15; k++) s += s_sq[i * 15 + k] * s_Jr[k][c];
            J_out[e] = s;
        }
    }
}

// MarginalizationFactor::Evaluate residual: r = r0 + J0 dx.  One CTA.
__device__ void eval_prior_cta(const DevPrior &p, const double *x, const StateLayout &lay, double *r_out, double *cost_out,
                               double *s_dx, double *s_red) {
    const int tid = threadIdx.x;
    for (int b = tid; b < p.n_blocks; b += blockDim.x) {
        const int id = p.block_id[b], size = lay.block_gsize(id), idx = p.block_idx[b];
        const double *xb = x + lay.block_state(id), *x0 = p.x0 + p.block_x0[b];
        if (size != 7) {
            for (int k = 0; k < size; k++) s_dx[idx + k] = xb[k] - x0[k];
        } else {
            for (int k = 0; k < 3; k++) s_dx[idx + k] = xb[k] - x0[k];
            const DQ dqq = qinv(dq(x0[6], x0[3], x0[4], x0[5])) * dq(xb[6], xb[3], xb[4], xb[5]);
            const double sgn = (dqq.w >= 0) ? 2.0 : -2.0;
            s_dx[idx + 3] = sgn * dqq.x;
            s_dx[idx + 4] = sgn * dqq.y;
            s_dx[idx + 5] = sgn * dqq.z;
        }
    }
    __syncthreads();
    // eight lanes per row (consecutive columns: 64-byte segments of J0), rows of a pass side by side; a thread per row read its
    // row with a stride of n doubles between lanes and summed 76 dependent terms: 6.8 us, the longest CTA of the kernel
    double c = 0;
    const int sub = tid & 7, grp = tid >> 3, ngrp = blockDim.x >> 3;
    for (int i0 = 0; i0 < p.n; i0 += ngrp) {
        const int i = i0 + grp;
        double s0 = 0, s1 = 0;
        if (i < p.n) {
            const double *row = p.J0 + (size_t)i * p.n;
            int k = sub;
            for (; k + 8 < p.n; k += 16) {
                s0 = fma(row[k], s_dx[k], s0);
                s1 = fma(row[k + 8], s_dx[k + 8], s1);
            }
            if (k < p.n) s0 = fma(row[k], s_dx[k], s0);
        }
        double s = s0 + s1;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (i < p.n && sub == 0) {
            s += p.r0[i];
            r_out[i] = s;
            c += 0.5 * s * s;
        }
    }
    c = warp_sum(c);
    if ((tid & 31) == 0) s_red[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
        double t = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += s_red[k];
        *cost_out = t;
    }
}

// Trust-region bookkeeping after a candidate evaluation (ceres TrustRegionMinimizer loop body, DoglegStrategy
// StepAccepted / StepRejected / StepIsInvalid).  Single thread.
__device__ void accept_logic(SolverCtl *ctl, const lfvio_solver_opts &o) {
    ctl->iter += 1;
    const int it = ctl->iter;
    int accepted = 0;
    if (ctl->step_valid) {
        const double cost_cand = atomicAdd(&ctl->cost_cand, 0.0);
        ctl->invalid_in_a_row = 0;
        const double cost_change = ctl->cost - cost_cand;
        if (!isfinite(cost_cand)) {
            ctl->step_valid = 0;
        } else if (!o.fixed_iterations && fabs(cost_change) <= o.function_tolerance * ctl->cost) {
            ctl->done = 1;  // "Function tolerance reached": returns without taking the step
        } else {
            const double rel = cost_change / ctl->model_cost_change;
            if (rel > o.min_relative_decrease) {
                if (rel < 0.25) ctl->radius *= 0.5;
                if (rel > 0.75) ctl->radius = fmax(ctl->radius, 3.0 * ctl->dogleg_step_norm);
                ctl->radius = fmin(ctl->radius, o.max_trust_region_radius);
                ctl->mu = fmax(1e-8, 2.0 * ctl->mu / 10.0);
                ctl->reuse = 0;
                ctl->cur ^= 1;
                ctl->cost = cost_cand;
                ctl->successful += 1;
                accepted = 1;
            } else {
                ctl->radius *= 0.5;
                ctl->reuse = 1;
            }
        }
    }
    if (!ctl->step_valid) {
        ctl->mu *= 10.0;
        ctl->reuse = 0;
        if (++ctl->invalid_in_a_row >= 5 && ctl->done == 0) ctl->done = -1;
    }
    if (it < LFVIO_MAX_ITER_LOG) {
        ctl->cost_log[it] = ctl->cost;
        ctl->radius_log[it] = ctl->radius;
        ctl->accepted_log[it] = accepted;
    }
    if (ctl->done == 0 && ctl->radius < o.min_trust_region_radius) ctl->done = 1;
    if (ctl->done == 0 && it >= o.max_num_iterations) ctl->done = 2;
    ctl->eval_pending = 0;
}

// mode 0: evaluate the current state unconditionally (initial point).  mode 1: evaluate the candidate if one
// is pending, then run accept_logic in the last CTA.  mode 2: plain evaluation of buffer 0 (debug entry point).
CTA_PROF_DECL(g_eval_prof);
__global__ void __launch_bounds__(EV_THREADS) evaluate_kernel(const DevProblem *__restrict__ Pd, const double *x0, const double *x1, EvalBuf e0,
                                                              EvalBuf e1, SolverCtl *ctl, int mode, int robust, int nbp, int n_imu_cta) {
    const DevProblem &P = *Pd;
    __shared__ double s_Jr[15][30];
    __shared__ double s_raw[16];
    __shared__ double s_res[16];
    __shared__ double s_sq[225];
    __shared__ double s_dx[15 * kMaxFrames + 8];
    __shared__ double s_red[EV_THREADS / 32];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    CTA_PROF(g_eval_prof, 0);
    pdl_prologue();
    CTA_PROF(g_eval_prof, 1);
    bool active = true;
    int target = 0;
    if (mode == 0) {
        target = ctl->cur;
    } else if (mode == 1) {
        if (ctl->done != 0) return;
        active = ctl->eval_pending != 0;
        target = ctl->cur ^ 1;
    }
    if (active) {
        const double *x = target ? x1 : x0;
        const EvalBuf &e = target ? e1 : e0;
        double *cost_acc = (mode == 1) ? &ctl->cost_cand : (mode == 2 ? &ctl->acc[7] : &ctl->cost);  // mode 2 never touches the solve's cost
        const int b = blockIdx.x;
        if (b < nbp) {
            const int k = (b * EV_THREADS + tid) >> 1;
            double c = 0;
            if (k < P.n_proj) eval_projection(P.proj[k], x, P.lay, P.o, robust, e.rp + 2 * (size_t)k, e.Jp + 40 * (size_t)k, &c, tid & 1);
            c = warp_sum(c);
            if ((tid & 31) == 0) s_red[tid >> 5] = c;
            __syncthreads();
            if (tid == 0) {
                double t = 0;
                for (int q = 0; q < EV_THREADS / 32; q++) t += s_red[q];
                atomicAdd(cost_acc, t);
            }
        } else if (b < nbp + n_imu_cta) {
            const int k = b - nbp;
            if (k >= P.n_imu) {
                // (launch geometry is sized for F - 1 IMU factors; fewer are present)
            } else if (P.imu[k].sum_dt <= 10.0) {  // estimator.cpp:720-721
                double c = 0;
                eval_imu_cta(P.imu[k], P.imu_sqrt_info + 225 * (size_t)k, x, P.lay, P.o, e.ri + 15 * (size_t)k,
                             e.Ji + 450 * (size_t)k, &c, s_Jr, s_raw, s_sq, s_res);
                if (tid == 0) atomicAdd(cost_acc, c);
            } else {
                for (int q = tid; q < 15; q += blockDim.x) e.ri[15 * (size_t)k + q] = 0.0;
                for (int q = tid; q < 450; q += blockDim.x) e.Ji[450 * (size_t)k + q] = 0.0;
            }
        } else if (P.prior.n > 0) {
            double c = 0;
            eval_prior_cta(P.prior, x, P.lay, e.rm, &c, s_dx, s_red);
            if (tid == 0) atomicAdd(cost_acc, c);
        }
    }
    if (mode == 2) return;
    // ---- last CTA done: bookkeeping ----
    __syncthreads();
    CTA_PROF(g_eval_prof, 2);
    if (tid == 0) {
        __threadfence();
        const unsigned t = atomicAdd(&ctl->ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && tid == 0) {
        __threadfence();
        ctl->ticket = 0;
        if (mode == 0) {
            const double c0 = atomicAdd(&ctl->cost, 0.0);
            ctl->cost_log[0] = c0;
            ctl->radius_log[0] = ctl->radius;
            ctl->accepted_log[0] = 1;
            if (!isfinite(c0)) ctl->done = -1;
        } else {
            accept_logic(ctl, P.o);
        }
    }
    CTA_PROF(g_eval_prof, 3);
}

// Eigen::LLT<Matrix15d>(covariance.inverse()).matrixL().transpose()  (imu_factor.h:64), once per solve.
// One warp per factor on a shared-memory 15 x 30 augmented matrix: Gauss-Jordan with partial pivoting (same pivot
// rule and operation order as the oracle's invert()), then Cholesky of the inverse.
__global__ void __launch_bounds__(32) imu_prepare_kernel(const DevProblem *__restrict__ Pd, double *out, int *status) {
    const lfvio_imu_factor *imu = Pd->imu;
    const int k = blockIdx.x, lane = threadIdx.x;
    if (k >= Pd->n_imu) return;
    __shared__ double w[15][31];
    __shared__ double Lm[15][16];
    if (imu[k].sum_dt > 10.0) {  // inactive factor: leave zeros
        for (int q = lane; q < 225; q += 32) out[225 * (size_t)k + q] = 0.0;
        return;
    }
    // Gauss-Jordan with the augmented matrix in REGISTERS, lane j = column j (0-14 covariance, 15-29 identity), 15 rows each: the
    // pivot search is a scan of lane `col`'s own registers, a row swap is a register swap in every lane, the multipliers
    // w[i][col] are broadcast from lane `col` — no shared-memory round trip and no warp-wide argmax per pivot (the shared-memory
    // version took ~2.3 k cycles per pivot, this one ~0.5 k).  Same pivot rule and the same operation per element as before.
    double a[15];
#pragma unroll
    for (int i = 0; i < 15; i++) a[i] = lane < 15 ? imu[k].covariance[i * 15 + lane] : ((lane - 15 == i) ? 1.0 : 0.0);
    bool ok = true;
#pragma unroll
    for (int col = 0; col < 15; col++) {
        // partial pivoting: first row (>= col) with the largest |w[i][col]| — all candidates live in lane `col`
        int piv = col;
        double best = fabs(a[col]);
#pragma unroll
        for (int i = col + 1; i < 15; i++) {
            const double v = fabs(a[i]);
            if (v > best) {
                best = v;
                piv = i;
            }
        }
        piv = __shfl_sync(0xffffffffu, piv, col);
        best = __shfl_sync(0xffffffffu, best, col);
        if (!(best > 0.0)) {
            ok = false;
            break;
        }
        if (piv != col) {  // swap rows col <-> piv (piv is uniform across the warp)
            const double t = a[col];
#pragma unroll
            for (int i = col + 1; i < 15; i++)
                if (i == piv) {
                    a[col] = a[i];
                    a[i] = t;
                }
        }
        const double d = __shfl_sync(0xffffffffu, a[col], col);
        a[col] /= d;
#pragma unroll
        for (int i = 0; i < 15; i++) {
            if (i == col) continue;
            const double f = __shfl_sync(0xffffffffu, a[i], col);  // w[i][col]
            if (lane == col) a[i] = 0.0;                           // w[i][col] -= f * 1
            else if (f != 0.0) a[i] -= f * a[col];
        }
    }
    if (ok && lane >= 15 && lane < 30) {
#pragma unroll
        for (int i = 0; i < 15; i++) w[i][lane] = a[i];
    }
    __syncwarp();
    // Cholesky (lower) of the inverse held in w[:, 15:30]
    if (ok) {
        for (int j = 0; j < 15; j++) {
            double d = w[j][15 + j];
            for (int q = 0; q < j; q++) d -= Lm[j][q] * Lm[j][q];
            if (!(d > 0.0)) {
                ok = false;
                break;
            }
            d = sqrt(d);
            __syncwarp();
            if (lane == 0) Lm[j][j] = d;
            if (lane > j && lane < 15) {
                double s = w[lane][15 + j];
                for (int q = 0; q < j; q++) s -= Lm[lane][q] * Lm[j][q];
                Lm[lane][j] = s / d;
            }
            __syncwarp();
        }
    }
    if (!ok) {
        if (lane == 0) *status = 1;
        return;
    }
    for (int q = lane; q < 225; q += 32) {
        const int i = q / 15, j = q % 15;
        out[225 * (size_t)k + q] = (j >= i) ? Lm[j][i] : 0.0;  // L^T
    }
}

// H0 = J0^T J0 of the prior (constant during a solve).
__global__ void prior_prepare_kernel(const DevProblem *__restrict__ Pd, double *H0) {
    const double *J0 = Pd->prior.J0;
    const int n = Pd->prior.n;
    const int a = blockIdx.x, b = threadIdx.x + blockIdx.y * blockDim.x;
    if (a >= n || b >= n) return;
    double s = 0;
    for (int i = 0; i < n; i++) s += J0[(size_t)i * n + a] * J0[(size_t)i * n + b];
    H0[(size_t)a * n + b] = s;
}

}  // namespace
CTA_PROF_READER(lfvio_debug_eval_prof, g_eval_prof)

int launch_imu_prepare(const DevProblem &P, double *imu_sqrt_info, int *status, cudaStream_t s) {
    imu_prepare_kernel<<<std::max(P.g_imu, 1), 32, 0, s>>>(P.self_dev, imu_sqrt_info, status);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

int launch_prior_prepare(const DevProblem &P, double *H0, cudaStream_t s) {
    prior_prepare_kernel<<<dim3(std::max(P.cap_n, 1), cdiv(std::max(P.cap_n, 1), 128)), 128, 0, s>>>(P.self_dev, H0);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

int launch_evaluate(const DevProblem &P, const double *x0, const double *x1, const EvalBuf &e0, const EvalBuf &e1,
                    SolverCtl *ctl, int mode, int robust, cudaStream_t s) {
    const int nbp = 2 * P.g_nbp;  // two threads per projection factor: >= cdiv(2 * P.n_proj, EV_THREADS)
    const int grid = nbp + P.g_imu + 1;
    LFVIO_CUDA_TRY(launch_pdl(evaluate_kernel, dim3(grid), dim3(EV_THREADS), 0, s, P.self_dev, x0, x1, e0, e1, ctl, mode, robust, nbp,
                              P.g_imu));
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

}  // namespace lfvio
