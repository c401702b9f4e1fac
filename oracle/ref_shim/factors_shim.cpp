// TEST INFRASTRUCTURE ONLY (oracle/): C entry points over the reference's OWN factor code, compiled UNMODIFIED from
// /root/reference/vins_estimator/src/factor/*.{h,cpp} + utility/utility.{h,cpp} (see oracle/Makefile, target _ref) against the
// stand-in headers of oracle/ref_stubs (Eigen / Ceres interface / ROS macros are third-party and absent from the image).
// What is the reference's here: every residual, Jacobian, pre-integration step, Plus(), and the whole of
// MarginalizationInfo::{addResidualBlockInfo, preMarginalize, marginalize, getParameterBlocks} / MarginalizationFactor::Evaluate.
// What is restated here (and says so): the few lines of Estimator::optimization() that pick the factors to marginalise
// (vins_estimator/src/estimator.cpp:833-1005) and the globals parameters.cpp would have read from the YAML.
// The structs are the interface structs of include/lfvio.h so one problem description feeds the product, the oracle and this.
#include <cstring>
#include <unordered_map>
#include <vector>

#include "factor/imu_factor.h"
#include "factor/marginalization_factor.h"
#include "factor/pose_local_parameterization.h"
#include "factor/projection_factor.h"
#include "factor/projection_td_factor.h"
#include "lfvio.h"

// ---- globals of vins_estimator/src/parameters.cpp:3-36 (that file needs cv::FileStorage and ros::NodeHandle: not compiled)
double INIT_DEPTH = 5.0, MIN_PARALLAX = 10.0 / 160.0;
double ACC_N = 0.02, ACC_W = 0.04, GYR_N = 0.01, GYR_W = 0.001;
std::vector<Eigen::Matrix3d> RIC;
std::vector<Eigen::Vector3d> TIC;
Eigen::Vector3d G{0.0, 0.0, 9.81007};
double BIAS_ACC_THRESHOLD = 0.1, BIAS_GYR_THRESHOLD = 0.1, SOLVER_TIME = 0.04;
int NUM_ITERATIONS = 8, ESTIMATE_EXTRINSIC = 1, ESTIMATE_TD = 1, ROLLING_SHUTTER = 0;
std::string EX_CALIB_RESULT_PATH, VINS_RESULT_PATH, IMU_TOPIC;
double ROW = 960, COL = 1280, TD = 0, TR = 0;
extern "C" int lfvio_ref_warn_count = 0;

namespace {

struct Params {  // the para_* arrays of estimator.h:107-113, contiguous so that block ids map to addresses
    int F, L;
    std::vector<double> pose, sb, ex, feat, td;
    explicit Params(const lfvio_window_state &s) : F(s.n_frames), L(s.n_landmarks) {
        pose.assign(s.para_Pose, s.para_Pose + 7 * F);
        sb.assign(s.para_SpeedBias, s.para_SpeedBias + 9 * F);
        ex.assign(s.para_Ex_Pose, s.para_Ex_Pose + 7);
        feat.assign(s.para_Feature, s.para_Feature + L);
        td.assign(s.para_Td, s.para_Td + 1);
    }
    double *block(int id) {  // include/lfvio.h: pose f -> 2f, speed-bias f -> 2f+1, extrinsic -> 2F, td -> 2F+1
        if (id == 2 * F) return ex.data();
        if (id == 2 * F + 1) return td.data();
        return (id & 1) ? &sb[9 * (id >> 1)] : &pose[7 * (id >> 1)];
    }
    int id_of(const double *p) {
        if (p == ex.data()) return 2 * F;
        if (p == td.data()) return 2 * F + 1;
        if (p >= pose.data() && p < pose.data() + 7 * F) return 2 * (int)((p - pose.data()) / 7);
        if (p >= sb.data() && p < sb.data() + 9 * F) return 2 * (int)((p - sb.data()) / 9) + 1;
        return -1;
    }
};

void set_globals(const lfvio_solver_opts &o) {
    G = Eigen::Vector3d(o.g[0], o.g[1], o.g[2]);
    ROW = (double)o.row;
    TR = o.tr;
    ESTIMATE_TD = o.estimate_td;
    ProjectionFactor::sqrt_info = o.focal_length / 1.5 * Eigen::Matrix2d::Identity();    // estimator.cpp:18-19
    ProjectionTdFactor::sqrt_info = o.focal_length / 1.5 * Eigen::Matrix2d::Identity();
}

IntegrationBase *make_preintegration(const lfvio_imu_factor &f) {  // an IntegrationBase holding the given members
    Eigen::Vector3d z = Eigen::Vector3d::Zero();
    IntegrationBase *ib = new IntegrationBase(z, z, Eigen::Vector3d(f.linearized_ba[0], f.linearized_ba[1], f.linearized_ba[2]),
                                              Eigen::Vector3d(f.linearized_bg[0], f.linearized_bg[1], f.linearized_bg[2]));
    ib->sum_dt = f.sum_dt;
    ib->delta_p = Eigen::Vector3d(f.delta_p[0], f.delta_p[1], f.delta_p[2]);
    ib->delta_q = Eigen::Quaterniond(f.delta_q[0], f.delta_q[1], f.delta_q[2], f.delta_q[3]);
    ib->delta_v = Eigen::Vector3d(f.delta_v[0], f.delta_v[1], f.delta_v[2]);
    for (int i = 0; i < 15; i++)
        for (int j = 0; j < 15; j++) {
            ib->jacobian(i, j) = f.jacobian[15 * i + j];
            ib->covariance(i, j) = f.covariance[15 * i + j];
        }
    return ib;
}

ceres::CostFunction *make_projection(const lfvio_proj_factor &p, const lfvio_solver_opts &o) {
    const Eigen::Vector3d pi(p.pts_i[0], p.pts_i[1], p.pts_i[2]), pj(p.pts_j[0], p.pts_j[1], p.pts_j[2]);
    if (!o.estimate_td) return new ProjectionFactor(pi, pj);
    return new ProjectionTdFactor(pi, pj, Eigen::Vector3d(p.vel_i[0], p.vel_i[1], p.vel_i[2]),
                                  Eigen::Vector3d(p.vel_j[0], p.vel_j[1], p.vel_j[2]), p.td_i, p.td_j, p.row_i, p.row_j, 0);
}

// a MarginalizationInfo in the state marginalize() + getParameterBlocks() leave it in, rebuilt from the interface struct
MarginalizationInfo *make_prior_info(const lfvio_prior &pr, Params &P, std::vector<double *> &blocks) {
    MarginalizationInfo *info = new MarginalizationInfo();
    info->m = 0;
    info->n = pr.n;
    int off = 0;
    blocks.clear();
    for (int b = 0; b < pr.n_blocks; b++) {
        const int id = pr.block_id[b];
        const int gs = (id == 2 * P.F + 1) ? 1 : (id == 2 * P.F ? 7 : ((id & 1) ? 9 : 7));
        double *data = new double[gs];
        std::memcpy(data, pr.x0 + off, sizeof(double) * gs);
        off += gs;
        info->keep_block_size.push_back(gs);
        info->keep_block_idx.push_back(pr.block_idx[b]);
        info->keep_block_data.push_back(data);
        info->parameter_block_data[(long)(1000 + b)] = data;  // owner, so that ~MarginalizationInfo frees it
        blocks.push_back(P.block(id));
    }
    info->linearized_jacobians.resize(pr.n, pr.n);
    info->linearized_residuals.resize(pr.n);
    for (int i = 0; i < pr.n; i++) {
        info->linearized_residuals(i) = pr.r0[i];
        for (int j = 0; j < pr.n; j++) info->linearized_jacobians(i, j) = pr.J0[(size_t)i * pr.n + j];
    }
    return info;
}

int local_size(int s) { return s == 7 ? 6 : s; }

}  // namespace

extern "C" {

// ceres::CostFunction::Evaluate of every factor (+ the loss correction of ResidualBlockInfo::Evaluate when robust != 0);
// output layout as lfvio_window_evaluate (include/lfvio.h).
int ref_evaluate(const lfvio_window_state *s, const lfvio_factor_list *fl, const lfvio_solver_opts *o, int robust,
                 double *res_proj, double *jac_proj, double *res_imu, double *jac_imu, double *res_prior, double *jac_prior,
                 double *cost) {
    set_globals(*o);
    Params P(*s);
    ceres::CauchyLoss loss(o->cauchy_a);
    double c = 0;
    for (int k = 0; k < fl->n_proj; k++) {
        const lfvio_proj_factor &p = fl->proj[k];
        std::vector<double *> pb{P.block(2 * p.frame_i), P.block(2 * p.frame_j), P.ex.data(), &P.feat[p.landmark]};
        if (o->estimate_td) pb.push_back(P.td.data());
        ResidualBlockInfo rbi(make_projection(p, *o), robust ? &loss : nullptr, pb, std::vector<int>{});
        if (robust) {
            double r[2], rho[3];
            rbi.cost_function->Evaluate(pb.data(), r, nullptr);
            loss.Evaluate(r[0] * r[0] + r[1] * r[1], rho);
            c += 0.5 * rho[0];
        }
        rbi.Evaluate();
        if (!robust) c += 0.5 * rbi.residuals.squaredNorm();
        if (res_proj) res_proj[2 * k] = rbi.residuals(0), res_proj[2 * k + 1] = rbi.residuals(1);
        if (jac_proj) {
            double *J = jac_proj + 40 * k;
            std::memset(J, 0, 40 * sizeof(double));
            const int col0[5] = {0, 6, 12, 18, 19};
            for (size_t b = 0; b < pb.size(); b++)
                for (int r = 0; r < 2; r++)
                    for (int cidx = 0; cidx < local_size((int)rbi.jacobians[b].cols()); cidx++)
                        J[20 * r + col0[b] + cidx] = rbi.jacobians[b](r, cidx);
        }
        delete[] rbi.raw_jacobians;
        delete rbi.cost_function;
    }
    for (int k = 0; k < fl->n_imu; k++) {
        const lfvio_imu_factor &f = fl->imu[k];
        IntegrationBase *ib = make_preintegration(f);
        std::vector<double *> pb{P.block(2 * f.frame_i), P.block(2 * f.frame_i + 1), P.block(2 * f.frame_i + 2), P.block(2 * f.frame_i + 3)};
        ResidualBlockInfo rbi(new IMUFactor(ib), nullptr, pb, std::vector<int>{});
        rbi.Evaluate();
        c += 0.5 * rbi.residuals.squaredNorm();
        if (res_imu)
            for (int r = 0; r < 15; r++) res_imu[15 * k + r] = rbi.residuals(r);
        if (jac_imu) {
            double *J = jac_imu + 450 * k;
            const int col0[4] = {0, 6, 15, 21};
            for (size_t b = 0; b < 4; b++)
                for (int r = 0; r < 15; r++)
                    for (int cidx = 0; cidx < local_size((int)rbi.jacobians[b].cols()); cidx++)
                        J[30 * r + col0[b] + cidx] = rbi.jacobians[b](r, cidx);
        }
        delete[] rbi.raw_jacobians;
        delete rbi.cost_function;
        delete ib;
    }
    if (fl->prior && fl->prior->n > 0) {
        std::vector<double *> pb;
        MarginalizationInfo *info = make_prior_info(*fl->prior, P, pb);
        const int n = info->n;
        ResidualBlockInfo rbi(new MarginalizationFactor(info), nullptr, pb, std::vector<int>{});
        rbi.Evaluate();
        c += 0.5 * rbi.residuals.squaredNorm();
        if (res_prior)
            for (int r = 0; r < n; r++) res_prior[r] = rbi.residuals(r);
        if (jac_prior) {
            std::memset(jac_prior, 0, sizeof(double) * n * n);
            for (size_t b = 0; b < pb.size(); b++)
                for (int r = 0; r < n; r++)
                    for (int cidx = 0; cidx < local_size((int)rbi.jacobians[b].cols()); cidx++)
                        jac_prior[(size_t)r * n + info->keep_block_idx[b] + cidx] = rbi.jacobians[b](r, cidx);
        }
        delete[] rbi.raw_jacobians;
        delete rbi.cost_function;
        delete info;
    }
    if (cost) *cost = c;
    return 0;
}

// PoseLocalParameterization::Plus (pose_local_parameterization.cpp:3-19); its methods are private in the class, public in the base
void ref_pose_plus(const double *x, const double *delta, double *out) {
    PoseLocalParameterization plp;
    static_cast<ceres::LocalParameterization &>(plp).Plus(x, delta, out);
}
void ref_pose_plus_jacobian(const double *x, double *jac42) {
    PoseLocalParameterization plp;
    static_cast<ceres::LocalParameterization &>(plp).ComputeJacobian(x, jac42);
}

// The next prior: factor selection restated from estimator.cpp:833-1005, everything else is MarginalizationInfo itself.
// Block order in `out` is whatever getParameterBlocks() yields (an unordered_map keyed by addresses).
int ref_marginalize(const lfvio_window_state *s, const lfvio_factor_list *fl, const lfvio_solver_opts *o, int flag,
                    lfvio_prior *out, int *m_out) {
    set_globals(*o);
    Params P(*s);
    const int F = P.F, W = F - 1;
    ceres::CauchyLoss *loss = new ceres::CauchyLoss(o->cauchy_a);
    std::vector<IntegrationBase *> ibs;
    MarginalizationInfo *info = new MarginalizationInfo();
    MarginalizationInfo *last = nullptr;
    std::vector<double *> last_blocks;
    if (fl->prior && fl->prior->n > 0) last = make_prior_info(*fl->prior, P, last_blocks);
    out->n = 0;
    out->n_blocks = 0;
    bool any = false;
    if (flag == LFVIO_MARGIN_OLD) {
        if (last) {
            std::vector<int> drop;
            for (int i = 0; i < (int)last_blocks.size(); i++)
                if (last_blocks[i] == P.block(0) || last_blocks[i] == P.block(1)) drop.push_back(i);
            info->addResidualBlockInfo(new ResidualBlockInfo(new MarginalizationFactor(last), nullptr, last_blocks, drop));
            any = true;
        }
        for (int k = 0; k < fl->n_imu; k++)
            if (fl->imu[k].frame_i == 0 && fl->imu[k].sum_dt < 10.0) {
                ibs.push_back(make_preintegration(fl->imu[k]));
                info->addResidualBlockInfo(new ResidualBlockInfo(new IMUFactor(ibs.back()), nullptr,
                                                                 std::vector<double *>{P.block(0), P.block(1), P.block(2), P.block(3)},
                                                                 std::vector<int>{0, 1}));
                any = true;
            }
        for (int k = 0; k < fl->n_proj; k++) {
            const lfvio_proj_factor &p = fl->proj[k];
            if (p.frame_i != 0) continue;
            std::vector<double *> pb{P.block(0), P.block(2 * p.frame_j), P.ex.data(), &P.feat[p.landmark]};
            if (o->estimate_td) pb.push_back(P.td.data());
            info->addResidualBlockInfo(new ResidualBlockInfo(make_projection(p, *o), loss, pb, std::vector<int>{0, 3}));
            any = true;
        }
    } else {
        bool has = false;
        for (double *b : last_blocks) has |= (b == P.block(2 * (W - 1)));
        if (last && has) {
            std::vector<int> drop;
            for (int i = 0; i < (int)last_blocks.size(); i++)
                if (last_blocks[i] == P.block(2 * (W - 1))) drop.push_back(i);
            info->addResidualBlockInfo(new ResidualBlockInfo(new MarginalizationFactor(last), nullptr, last_blocks, drop));
            any = true;
        }
    }
    int rc = 0;
    if (any) {
        info->preMarginalize();
        info->marginalize();
        std::unordered_map<long, double *> addr_shift;
        if (flag == LFVIO_MARGIN_OLD) {
            for (int i = 1; i <= W; i++) {
                addr_shift[reinterpret_cast<long>(P.block(2 * i))] = P.block(2 * (i - 1));
                addr_shift[reinterpret_cast<long>(P.block(2 * i + 1))] = P.block(2 * (i - 1) + 1);
            }
        } else {
            for (int i = 0; i <= W; i++) {
                if (i == W - 1) continue;
                const int t = (i == W) ? i - 1 : i;
                addr_shift[reinterpret_cast<long>(P.block(2 * i))] = P.block(2 * t);
                addr_shift[reinterpret_cast<long>(P.block(2 * i + 1))] = P.block(2 * t + 1);
            }
        }
        addr_shift[reinterpret_cast<long>(P.ex.data())] = P.ex.data();
        if (o->estimate_td) addr_shift[reinterpret_cast<long>(P.td.data())] = P.td.data();
        std::vector<double *> kept = info->getParameterBlocks(addr_shift);
        const int n = info->n;
        if (m_out) *m_out = info->m;
        if (n > out->cap_n || (int)kept.size() > out->cap_blocks)
            rc = -3;
        else {
            out->n = n;
            out->n_blocks = (int)kept.size();
            int off = 0;
            for (size_t b = 0; b < kept.size(); b++) {
                out->block_id[b] = P.id_of(kept[b]);
                out->block_idx[b] = info->keep_block_idx[b] - info->m;
                std::memcpy(out->x0 + off, info->keep_block_data[b], sizeof(double) * info->keep_block_size[b]);
                off += info->keep_block_size[b];
            }
            for (int i = 0; i < n; i++) {
                out->r0[i] = info->linearized_residuals(i);
                for (int j = 0; j < n; j++) out->J0[(size_t)i * n + j] = info->linearized_jacobians(i, j);
            }
        }
    }
    delete info;  // frees its factors' cost functions, hence `last` through the MarginalizationFactor? no: see below
    // MarginalizationInfo::~MarginalizationInfo deletes each factor's cost_function; MarginalizationFactor does not own `last`
    delete last;
    for (IntegrationBase *ib : ibs) delete ib;
    delete loss;
    return rc;
}

// IntegrationBase::push_back over each interval (integration_base.h:30-157); out[i] = the members an IMUFactor reads.
int ref_preintegrate(const lfvio_imu_interval *iv, int n_intervals, const double *dt, const double *acc, const double *gyr,
                     int n_samples, const lfvio_imu_noise *noise, const double *g, lfvio_imu_factor *out) {
    ACC_N = noise->acc_n;
    GYR_N = noise->gyr_n;
    ACC_W = noise->acc_w;
    GYR_W = noise->gyr_w;
    G = Eigen::Vector3d(g[0], g[1], g[2]);
    for (int i = 0; i < n_intervals; i++) {
        const lfvio_imu_interval &v = iv[i];
        if (v.first < 0 || v.first + v.count > n_samples) return -1;
        IntegrationBase ib(Eigen::Vector3d(v.acc_0[0], v.acc_0[1], v.acc_0[2]), Eigen::Vector3d(v.gyr_0[0], v.gyr_0[1], v.gyr_0[2]),
                           Eigen::Vector3d(v.ba[0], v.ba[1], v.ba[2]), Eigen::Vector3d(v.bg[0], v.bg[1], v.bg[2]));
        for (int k = v.first; k < v.first + v.count; k++)
            ib.push_back(dt[k], Eigen::Vector3d(acc[3 * k], acc[3 * k + 1], acc[3 * k + 2]),
                         Eigen::Vector3d(gyr[3 * k], gyr[3 * k + 1], gyr[3 * k + 2]));
        lfvio_imu_factor &f = out[i];
        std::memset(&f, 0, sizeof(f));
        f.frame_i = v.frame_i;
        f.sum_dt = ib.sum_dt;
        for (int a = 0; a < 3; a++) {
            f.delta_p[a] = ib.delta_p(a);
            f.delta_v[a] = ib.delta_v(a);
            f.linearized_ba[a] = ib.linearized_ba(a);
            f.linearized_bg[a] = ib.linearized_bg(a);
        }
        f.delta_q[0] = ib.delta_q.w();
        f.delta_q[1] = ib.delta_q.x();
        f.delta_q[2] = ib.delta_q.y();
        f.delta_q[3] = ib.delta_q.z();
        for (int a = 0; a < 15; a++)
            for (int b = 0; b < 15; b++) {
                f.jacobian[15 * a + b] = ib.jacobian(a, b);
                f.covariance[15 * a + b] = ib.covariance(a, b);
            }
    }
    return 0;
}

// utility/utility.h:66-108, utility.cpp:3-12
void ref_R2ypr(const double *R9, double *ypr) {
    Eigen::Matrix3d R;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R(i, j) = R9[3 * i + j];
    Eigen::Vector3d y = Utility::R2ypr(R);
    for (int i = 0; i < 3; i++) ypr[i] = y(i);
}
void ref_ypr2R(const double *ypr, double *R9) {
    Eigen::Matrix3d R = Utility::ypr2R(Eigen::Vector3d(ypr[0], ypr[1], ypr[2]));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R9[3 * i + j] = R(i, j);
}
void ref_g2R(const double *g, double *R9) {
    Eigen::Matrix3d R = Utility::g2R(Eigen::Vector3d(g[0], g[1], g[2]));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R9[3 * i + j] = R(i, j);
}
int ref_warn_count(void) { return lfvio_ref_warn_count; }

}  // extern "C"
