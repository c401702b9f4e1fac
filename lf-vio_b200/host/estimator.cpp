// Estimator::optimization() host mirror: packs the problem, calls the device solver, applies the gauge fix,
// builds the next prior.  No arithmetic of the hot path happens here (only the 3x3 gauge re-anchoring of a15).
#include "estimator.h"

#include <cmath>
#include <cstring>

namespace lfvio_host {

int ESTIMATE_TD = 1, ESTIMATE_EXTRINSIC = 1, NUM_ITERATIONS = 8, EST_ROW = 960;
double TR = 0.0, TD = -0.008, FOCAL_LENGTH = 160.0;
Vec3 G{0.0, 0.0, 9.81007};

namespace {

Mat3 matmul(const Mat3 &a, const Mat3 &b) {
    Mat3 r{};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
    return r;
}
Vec3 matvec(const Mat3 &a, const Vec3 &v) {
    return {a[0] * v[0] + a[1] * v[1] + a[2] * v[2], a[3] * v[0] + a[4] * v[1] + a[5] * v[2], a[6] * v[0] + a[7] * v[1] + a[8] * v[2]};
}
Mat3 transpose(const Mat3 &a) { return {a[0], a[3], a[6], a[1], a[4], a[7], a[2], a[5], a[8]}; }

// Eigen::Quaterniond q{R}: returns (x, y, z, w)
std::array<double, 4> rot_to_quat(const Mat3 &m) {
    auto M = [&](int r, int c) { return m[3 * r + c]; };
    double q[4];  // x y z w
    double t = M(0, 0) + M(1, 1) + M(2, 2);
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (M(2, 1) - M(1, 2)) * t;
        q[1] = (M(0, 2) - M(2, 0)) * t;
        q[2] = (M(1, 0) - M(0, 1)) * t;
    } else {
        int i = 0;
        if (M(1, 1) > M(0, 0)) i = 1;
        if (M(2, 2) > M(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (M(k, j) - M(j, k)) * t;
        q[j] = (M(j, i) + M(i, j)) * t;
        q[k] = (M(k, i) + M(i, k)) * t;
    }
    return {q[0], q[1], q[2], q[3]};
}
// Quaterniond(w, x, y, z)[.normalized()].toRotationMatrix()
Mat3 quat_to_rot(double w, double x, double y, double z, bool normalize) {
    if (normalize) {
        const double n = std::sqrt(w * w + x * x + y * y + z * z);
        w /= n;
        x /= n;
        y /= n;
        z /= n;
    }
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    return {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
}

}  // namespace

Vec3 R2ypr(const Mat3 &R) {
    const Vec3 n{R[0], R[3], R[6]}, o{R[1], R[4], R[7]}, a{R[2], R[5], R[8]};
    const double y = std::atan2(n[1], n[0]);
    const double p = std::atan2(-n[2], n[0] * std::cos(y) + n[1] * std::sin(y));
    const double r = std::atan2(a[0] * std::sin(y) - a[1] * std::cos(y), -o[0] * std::sin(y) + o[1] * std::cos(y));
    return {y / M_PI * 180.0, p / M_PI * 180.0, r / M_PI * 180.0};
}
Mat3 ypr2R(const Vec3 &ypr) {
    const double y = ypr[0] / 180.0 * M_PI, p = ypr[1] / 180.0 * M_PI, r = ypr[2] / 180.0 * M_PI;
    const Mat3 Rz{std::cos(y), -std::sin(y), 0, std::sin(y), std::cos(y), 0, 0, 0, 1};
    const Mat3 Ry{std::cos(p), 0, std::sin(p), 0, 1, 0, -std::sin(p), 0, std::cos(p)};
    const Mat3 Rx{1, 0, 0, 0, std::cos(r), -std::sin(r), 0, std::sin(r), std::cos(r)};
    return matmul(matmul(Rz, Ry), Rx);
}

// ---- FeatureManager (feature_manager.cpp:36-43,145-197) --------------------------------------------------
int FeatureManager::getFeatureCount() {
    int cnt = 0;
    for (auto &it : feature) {
        it.used_num = (int)it.feature_per_frame.size();
        if (it.used_num >= 2 && it.start_frame < window_size - 2) cnt++;
    }
    return cnt;
}
std::vector<double> FeatureManager::getDepthVector() {
    std::vector<double> dep(getFeatureCount());
    int idx = -1;
    for (auto &it : feature) {
        it.used_num = (int)it.feature_per_frame.size();
        if (!(it.used_num >= 2 && it.start_frame < window_size - 2)) continue;
        dep[++idx] = 1. / it.estimated_depth;
    }
    return dep;
}
void FeatureManager::setDepth(const std::vector<double> &x) {
    int idx = -1;
    for (auto &it : feature) {
        it.used_num = (int)it.feature_per_frame.size();
        if (!(it.used_num >= 2 && it.start_frame < window_size - 2)) continue;
        it.estimated_depth = 1.0 / x[++idx];
        it.solve_flag = 1;  // both branches of the reference set 1 (feature_manager.cpp:151-157)
    }
}

// ---- Estimator ---------------------------------------------------------------------------------------
Estimator::Estimator(int window_size, int max_features, int device)
    : WINDOW_SIZE(window_size), Ps(window_size + 1), Vs(window_size + 1), Bas(window_size + 1), Bgs(window_size + 1),
      Rs(window_size + 1, Mat3{1, 0, 0, 0, 1, 0, 0, 0, 1}), pre_integrations(window_size + 1, nullptr),
      para_Pose(7 * (window_size + 1)), para_SpeedBias(9 * (window_size + 1)), para_Feature(max_features), m_device(device) {
    tic[0] = Vec3{0, 0, 0};
    ric[0] = Mat3{1, 0, 0, 0, 1, 0, 0, 0, 1};
    f_manager.window_size = window_size;
    last_status = lfvio_window_create(window_size, max_features, 16 * max_features, device, &m_window);
}
Estimator::~Estimator() {
    if (m_window) lfvio_window_destroy(m_window);
}
const char *Estimator::lastError() const { return lfvio_last_error(); }

void Estimator::vector2double() {
    for (int i = 0; i <= WINDOW_SIZE; i++) {
        const auto q = rot_to_quat(Rs[i]);
        double *p = &para_Pose[7 * i];
        p[0] = Ps[i][0]; p[1] = Ps[i][1]; p[2] = Ps[i][2];
        p[3] = q[0]; p[4] = q[1]; p[5] = q[2]; p[6] = q[3];
        double *s = &para_SpeedBias[9 * i];
        for (int k = 0; k < 3; k++) {
            s[k] = Vs[i][k];
            s[3 + k] = Bas[i][k];
            s[6 + k] = Bgs[i][k];
        }
    }
    const auto q = rot_to_quat(ric[0]);
    for (int k = 0; k < 3; k++) para_Ex_Pose[0][k] = tic[0][k];
    para_Ex_Pose[0][3] = q[0]; para_Ex_Pose[0][4] = q[1]; para_Ex_Pose[0][5] = q[2]; para_Ex_Pose[0][6] = q[3];
    const std::vector<double> dep = f_manager.getDepthVector();
    if (para_Feature.size() < dep.size()) para_Feature.resize(dep.size());
    for (size_t i = 0; i < dep.size(); i++) para_Feature[i] = dep[i];
    if (ESTIMATE_TD) para_Td[0][0] = td;
}

void Estimator::double2vector() {
    Vec3 origin_R0 = R2ypr(Rs[0]);
    Vec3 origin_P0 = Ps[0];
    if (failure_occur) {
        origin_R0 = R2ypr(last_R0);
        origin_P0 = last_P0;
        failure_occur = false;
    }
    const double *p0 = &para_Pose[0];
    const Mat3 R00 = quat_to_rot(p0[6], p0[3], p0[4], p0[5], false);
    const Vec3 origin_R00 = R2ypr(R00);
    const double y_diff = origin_R0[0] - origin_R00[0];
    Mat3 rot_diff = ypr2R(Vec3{y_diff, 0, 0});
    if (std::fabs(std::fabs(origin_R0[1]) - 90) < 1.0 || std::fabs(std::fabs(origin_R00[1]) - 90) < 1.0)
        rot_diff = matmul(Rs[0], transpose(R00));  // euler singular point
    for (int i = 0; i <= WINDOW_SIZE; i++) {
        const double *p = &para_Pose[7 * i], *s = &para_SpeedBias[9 * i];
        Rs[i] = matmul(rot_diff, quat_to_rot(p[6], p[3], p[4], p[5], true));
        const Vec3 dp = matvec(rot_diff, Vec3{p[0] - p0[0], p[1] - p0[1], p[2] - p0[2]});
        Ps[i] = Vec3{dp[0] + origin_P0[0], dp[1] + origin_P0[1], dp[2] + origin_P0[2]};
        Vs[i] = matvec(rot_diff, Vec3{s[0], s[1], s[2]});
        Bas[i] = Vec3{s[3], s[4], s[5]};
        Bgs[i] = Vec3{s[6], s[7], s[8]};
    }
    tic[0] = Vec3{para_Ex_Pose[0][0], para_Ex_Pose[0][1], para_Ex_Pose[0][2]};
    ric[0] = quat_to_rot(para_Ex_Pose[0][6], para_Ex_Pose[0][3], para_Ex_Pose[0][4], para_Ex_Pose[0][5], false);
    std::vector<double> dep(f_manager.getFeatureCount());
    for (size_t i = 0; i < dep.size(); i++) dep[i] = para_Feature[i];
    f_manager.setDepth(dep);
    if (ESTIMATE_TD) td = para_Td[0][0];
}

// Residual blocks in the reference's insertion order (estimator.cpp:717-772).
void Estimator::buildFactors() {
    m_imu.clear();
    for (int i = 0; i < WINDOW_SIZE; i++) {
        const int j = i + 1;
        if (!pre_integrations[j]) continue;
        lfvio_imu_factor f = *pre_integrations[j];
        f.frame_i = i;
        m_imu.push_back(f);  // the > 10 s rule (estimator.cpp:720-721) is applied on device
    }
    m_proj.clear();
    int feature_index = -1;
    for (auto &it_per_id : f_manager.feature) {
        it_per_id.used_num = (int)it_per_id.feature_per_frame.size();
        if (!(it_per_id.used_num >= 2 && it_per_id.start_frame < WINDOW_SIZE - 2)) continue;
        ++feature_index;
        const int imu_i = it_per_id.start_frame;
        int imu_j = imu_i - 1;
        const FeaturePerFrame &f0 = it_per_id.feature_per_frame[0];
        for (auto &it_per_frame : it_per_id.feature_per_frame) {
            imu_j++;
            if (imu_i == imu_j) continue;
            lfvio_proj_factor f;
            memset(&f, 0, sizeof(f));
            f.frame_i = imu_i;
            f.frame_j = imu_j;
            f.landmark = feature_index;
            for (int k = 0; k < 3; k++) {
                f.pts_i[k] = f0.point[k];
                f.pts_j[k] = it_per_frame.point[k];
                f.vel_i[k] = f0.velocity[k];
                f.vel_j[k] = it_per_frame.velocity[k];
            }
            f.td_i = f0.cur_td;
            f.td_j = it_per_frame.cur_td;
            f.row_i = f0.uv[1];
            f.row_j = it_per_frame.uv[1];
            m_proj.push_back(f);
        }
    }
}

void Estimator::optimization() {
    if (!m_window) return;
    vector2double();
    buildFactors();
    lfvio_solver_opts o;
    lfvio_solver_default_opts(&o);
    o.max_num_iterations = NUM_ITERATIONS;
    o.estimate_td = ESTIMATE_TD;
    o.estimate_extrinsic = ESTIMATE_EXTRINSIC ? 1 : 0;
    o.row = EST_ROW;
    o.focal_length = FOCAL_LENGTH;
    o.tr = TR;
    for (int k = 0; k < 3; k++) o.g[k] = G[k];
    lfvio_prior prior;
    memset(&prior, 0, sizeof(prior));
    MarginalizationInfo &lm = last_marginalization_info;
    if (lm.valid()) {
        prior.n = lm.n;
        prior.n_blocks = lm.n_blocks;
        prior.block_id = lm.block_id.data();
        prior.block_idx = lm.block_idx.data();
        prior.x0 = lm.x0.data();
        prior.J0 = lm.J0.data();
        prior.r0 = lm.r0.data();
    }
    const int L = f_manager.getFeatureCount();
    lfvio_window_state st{WINDOW_SIZE + 1, L, para_Pose.data(), para_SpeedBias.data(), para_Ex_Pose[0], para_Feature.data(), para_Td[0]};
    lfvio_factor_list fl{(int)m_proj.size(), m_proj.data(), (int)m_imu.size(), m_imu.data(), lm.valid() ? &prior : nullptr};
    last_status = lfvio_window_optimize(m_window, &st, &fl, &o, &summary);  // ceres::Solve (estimator.cpp:825)
    if (last_status != LFVIO_OK) return;
    double2vector();
    // ---- next prior (estimator.cpp:833-1005): the dropped factors are evaluated at the re-anchored state ----
    vector2double();
    const int nmax = 15 * (WINDOW_SIZE + 1) + 7, bmax = 2 * (WINDOW_SIZE + 1) + 2;
    MarginalizationInfo next;
    next.block_id.resize(bmax);
    next.block_idx.resize(bmax);
    next.x0.resize(9 * (size_t)bmax);
    next.J0.resize((size_t)nmax * nmax);
    next.r0.resize(nmax);
    lfvio_prior out;
    memset(&out, 0, sizeof(out));
    out.block_id = next.block_id.data();
    out.block_idx = next.block_idx.data();
    out.x0 = next.x0.data();
    out.J0 = next.J0.data();
    out.r0 = next.r0.data();
    out.cap_n = nmax;
    out.cap_blocks = bmax;
    last_status = lfvio_window_marginalize(m_window, &st, &fl, &o, marginalization_flag == MARGIN_OLD ? LFVIO_MARGIN_OLD : LFVIO_MARGIN_SECOND_NEW, &out);
    if (last_status != LFVIO_OK) return;
    if (out.n > 0) {  // MARGIN_SECOND_NEW without a matching prior leaves the old one in place (estimator.cpp:942-943)
        next.n = out.n;
        next.n_blocks = out.n_blocks;
        last_marginalization_info = std::move(next);
    }
}

}  // namespace lfvio_host

// ---- C entry point used by the tests to drive the C++ class from Python (ctypes) --------------------------
// Loads a flat problem into an Estimator (state + FeatureManager + pre-integrations), runs optimization() and
// returns the re-anchored state and the new prior.  Not part of include/lfvio.h: it is a test hook of the host mirror.
extern "C" int lfvio_host_estimator_roundtrip(const lfvio_window_state *st, const lfvio_factor_list *fl, int estimate_td,
                                              int margin_flag, double *pose_out, double *sb_out, double *ex_out,
                                              double *feat_out, double *td_out, lfvio_prior *prior_out, int *iterations) {
    using namespace lfvio_host;
    const int F = st->n_frames;
    ESTIMATE_TD = estimate_td;
    Estimator est(F - 1, st->n_landmarks > 0 ? st->n_landmarks : 1, 0);
    if (est.last_status != LFVIO_OK) return est.last_status;
    auto q2r = [](const double *p) {
        const double w = p[6], x = p[3], y = p[4], z = p[5];
        const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
        const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
        return Mat3{1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
    };
    for (int i = 0; i < F; i++) {
        const double *p = st->para_Pose + 7 * i, *s = st->para_SpeedBias + 9 * i;
        est.Ps[i] = Vec3{p[0], p[1], p[2]};
        est.Rs[i] = q2r(p);
        est.Vs[i] = Vec3{s[0], s[1], s[2]};
        est.Bas[i] = Vec3{s[3], s[4], s[5]};
        est.Bgs[i] = Vec3{s[6], s[7], s[8]};
    }
    est.tic[0] = Vec3{st->para_Ex_Pose[0], st->para_Ex_Pose[1], st->para_Ex_Pose[2]};
    est.ric[0] = q2r(st->para_Ex_Pose);
    est.td = st->para_Td[0];
    std::vector<lfvio_imu_factor> imu(fl->imu, fl->imu + fl->n_imu);
    for (auto &f : imu) est.pre_integrations[f.frame_i + 1] = &f;
    // FeatureManager from the factor list: one FeaturePerId per landmark, observations in frame order
    int k = 0;
    for (int l = 0; l < st->n_landmarks; l++) {
        FeaturePerId fp;
        fp.feature_id = l;
        fp.estimated_depth = 1.0 / st->para_Feature[l];
        bool first = true;
        while (k < fl->n_proj && fl->proj[k].landmark == l) {
            const lfvio_proj_factor &f = fl->proj[k];
            if (first) {
                fp.start_frame = f.frame_i;
                FeaturePerFrame o;
                o.point = Vec3{f.pts_i[0], f.pts_i[1], f.pts_i[2]};
                o.velocity = Vec3{f.vel_i[0], f.vel_i[1], f.vel_i[2]};
                o.cur_td = f.td_i;
                o.uv = {0.0, f.row_i};
                fp.feature_per_frame.push_back(o);
                first = false;
            }
            FeaturePerFrame o;
            o.point = Vec3{f.pts_j[0], f.pts_j[1], f.pts_j[2]};
            o.velocity = Vec3{f.vel_j[0], f.vel_j[1], f.vel_j[2]};
            o.cur_td = f.td_j;
            o.uv = {0.0, f.row_j};
            fp.feature_per_frame.push_back(o);
            k++;
        }
        est.f_manager.feature.push_back(fp);
    }
    if (fl->prior && fl->prior->n > 0) {
        MarginalizationInfo &m = est.last_marginalization_info;
        const lfvio_prior &p = *fl->prior;
        m.n = p.n;
        m.n_blocks = p.n_blocks;
        m.block_id.assign(p.block_id, p.block_id + p.n_blocks);
        m.block_idx.assign(p.block_idx, p.block_idx + p.n_blocks);
        int x0_len = 0;
        for (int b = 0; b < p.n_blocks; b++) x0_len += p.block_id[b] < 2 * F ? ((p.block_id[b] & 1) ? 9 : 7) : (p.block_id[b] == 2 * F ? 7 : 1);
        m.x0.assign(p.x0, p.x0 + x0_len);
        m.J0.assign(p.J0, p.J0 + (size_t)p.n * p.n);
        m.r0.assign(p.r0, p.r0 + p.n);
    }
    est.marginalization_flag = margin_flag == 0 ? Estimator::MARGIN_OLD : Estimator::MARGIN_SECOND_NEW;
    est.optimization();
    if (est.last_status != LFVIO_OK) return est.last_status;
    est.vector2double();
    memcpy(pose_out, est.para_Pose.data(), sizeof(double) * 7 * F);
    memcpy(sb_out, est.para_SpeedBias.data(), sizeof(double) * 9 * F);
    memcpy(ex_out, est.para_Ex_Pose[0], sizeof(double) * 7);
    for (int l = 0; l < st->n_landmarks; l++) feat_out[l] = est.para_Feature[l];
    td_out[0] = est.td;
    *iterations = est.summary.iterations;
    const MarginalizationInfo &m = est.last_marginalization_info;
    prior_out->n = 0;
    prior_out->n_blocks = 0;
    if (m.valid() && m.n <= prior_out->cap_n && m.n_blocks <= prior_out->cap_blocks) {
        prior_out->n = m.n;
        prior_out->n_blocks = m.n_blocks;
        memcpy(prior_out->block_id, m.block_id.data(), sizeof(int) * m.n_blocks);
        memcpy(prior_out->block_idx, m.block_idx.data(), sizeof(int) * m.n_blocks);
        memcpy(prior_out->J0, m.J0.data(), sizeof(double) * (size_t)m.n * m.n);
        memcpy(prior_out->r0, m.r0.data(), sizeof(double) * m.n);
        memcpy(prior_out->x0, m.x0.data(), sizeof(double) * 9 * m.n_blocks);
    }
    return LFVIO_OK;
}
