// Estimator::initialStructure / visualInitialAlign / relativePose host mirror (vins_estimator/src/estimator.cpp:221-484)
// over the lfvio C-ABI.  The control flow and the bookkeeping (all_image_frame, key-frame matching, the state rescaling
// and gravity alignment of visualInitialAlign — a handful of 3x3 products) stay on the host as in the reference; every
// numeric step with a loop over features, frames or iterations is a device call:
//   relativePose           -> lfvio_solve_relative_rt   (solve_5pts.cpp:538-575)
//   GlobalSFM::construct   -> lfvio_global_sfm          (initial_sfm.cpp:127-316)
//   PnP of every frame     -> lfvio_epnp, ONE call for all non-keyframes (estimator.cpp:320-357)
//   solveGyroscopeBias     -> lfvio_solve_gyro_bias + lfvio_preintegrate (repropagation of every frame, one call)
//   LinearAlignment        -> lfvio_linear_alignment    (initial_aligment.cpp:56-205)
//   triangulate            -> lfvio_triangulate         (feature_manager.cpp:199-253)
// Not mirrored: the IMU-excitation check (:224-249) only logs in the reference (its `return false` is commented out).
#include <cmath>

#include "estimator.h"

namespace lfvio_host {

using namespace detail;

namespace {

double norm3(const Vec3 &v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

// Eigen::Quaterniond::FromTwoVectors(a, b).toRotationMatrix() for unit vectors
Mat3 from_two_vectors(const Vec3 &v0, const Vec3 &v1) {
    const double c = v1[0] * v0[0] + v1[1] * v0[1] + v1[2] * v0[2];
    if (c < -1.0 + 1e-12) return Mat3{1, 0, 0, 0, -1, 0, 0, 0, -1};  // antiparallel: any half turn (Eigen takes one from an SVD)
    const Vec3 axis{v0[1] * v1[2] - v0[2] * v1[1], v0[2] * v1[0] - v0[0] * v1[2], v0[0] * v1[1] - v0[1] * v1[0]};
    const double s = std::sqrt((1.0 + c) * 2.0), invs = 1.0 / s;
    return quat_to_rot(s * 0.5, axis[0] * invs, axis[1] * invs, axis[2] * invs, false);
}

Mat3 g2R(const Vec3 &g) {  // utility/utility.cpp:3-12
    const double n = norm3(g);
    Mat3 R0 = from_two_vectors(Vec3{g[0] / n, g[1] / n, g[2] / n}, Vec3{0, 0, 1.0});
    const double yaw = R2ypr(R0)[0];
    return matmul(ypr2R(Vec3{-yaw, 0, 0}), R0);
}

}  // namespace

std::vector<std::pair<Vec3, Vec3>> FeatureManager::getCorresponding(int frame_count_l, int frame_count_r) {
    std::vector<std::pair<Vec3, Vec3>> corres;
    for (auto &it : feature) {
        const int end_frame = it.start_frame + (int)it.feature_per_frame.size() - 1;
        if (it.start_frame <= frame_count_l && end_frame >= frame_count_r)
            corres.emplace_back(it.feature_per_frame[frame_count_l - it.start_frame].point,
                                it.feature_per_frame[frame_count_r - it.start_frame].point);
    }
    return corres;
}

void FeatureManager::clearDepth() {
    for (auto &it : feature) {
        it.used_num = (int)it.feature_per_frame.size();
        if (!(it.used_num >= 2 && it.start_frame < window_size - 2)) continue;
        it.estimated_depth = 1.0 / -1.0;
    }
}

// (Re)integrate the pre_integration of the frames of all_image_frame in one lfvio_preintegrate call: all of them with
// new_bg (solveGyroscopeBias' repropagate(Vector3d::Zero(), Bgs[0]), initial_aligment.cpp:31-37) or, with all == false,
// only those that have no record yet.
int Estimator::integrateImageFrames(bool all, const Vec3 *new_bg) {
    std::vector<lfvio_imu_interval> iv;
    std::vector<ImageFrame *> who;
    std::vector<double> dt, acc, gyr;
    for (auto &kv : all_image_frame) {
        ImageFrame &f = kv.second;
        if (!f.has_pre) continue;
        if (new_bg) {
            f.ba = Vec3{0, 0, 0};
            f.bg = *new_bg;
        } else if (!all && f.pre.sum_dt > 0) {
            continue;
        }
        lfvio_imu_interval I{};
        I.first = (int)dt.size();
        I.count = (int)f.dt.size();
        I.frame_i = 0;
        I.propagate = 0;
        for (int k = 0; k < 3; k++) {
            I.acc_0[k] = f.acc_start[k], I.gyr_0[k] = f.gyr_start[k];
            I.ba[k] = f.ba[k], I.bg[k] = f.bg[k];
        }
        I.R[0] = I.R[4] = I.R[8] = 1.0;
        for (size_t s = 0; s < f.dt.size(); s++) {
            dt.push_back(f.dt[s]);
            acc.insert(acc.end(), f.acc[s].begin(), f.acc[s].end());
            gyr.insert(gyr.end(), f.gyr[s].begin(), f.gyr[s].end());
        }
        iv.push_back(I);
        who.push_back(&f);
    }
    if (iv.empty()) return LFVIO_OK;
    std::vector<lfvio_imu_factor> out(iv.size());
    acc.resize(acc.size() + 3);
    gyr.resize(gyr.size() + 3);
    dt.resize(dt.size() + 1);
    const int st = lfvio_preintegrate(iv.data(), (int)iv.size(), dt.data(), acc.data(), gyr.data(), (int)dt.size() - 1, &imu_noise, G.data(),
                                      out.data(), nullptr, m_device);
    if (st != LFVIO_OK) return st;
    for (size_t i = 0; i < who.size(); i++) who[i]->pre = out[i];
    return LFVIO_OK;
}

bool InitialEXRotation::CalibrationExRotation(const std::vector<std::pair<Vec3, Vec3>> &corres, const std::array<double, 4> &dq,
                                              Mat3 &calib_ric_result) {
    frame_count++;
    std::vector<double> flat;
    flat.reserve(6 * corres.size());
    for (const auto &c : corres) {
        flat.insert(flat.end(), c.first.begin(), c.first.end());
        flat.insert(flat.end(), c.second.begin(), c.second.end());
    }
    Mat3 rc{};
    status = lfvio_solve_relative_r(flat.data(), (int)corres.size(), ransac_seed, (unsigned)frame_count, rc.data(), device);
    if (status != LFVIO_OK) return false;
    const Mat3 ri = quat_to_rot(dq[0], dq[1], dq[2], dq[3], false);
    const Mat3 rg = matmul(matmul(transpose(ric), ri), ric);  // ric.inverse() * delta_q_imu * ric
    Rc.insert(Rc.end(), rc.begin(), rc.end());
    Rimu.insert(Rimu.end(), ri.begin(), ri.end());
    Rc_g.insert(Rc_g.end(), rg.begin(), rg.end());
    status = lfvio_calibrate_ex_rotation(Rc.data(), Rimu.data(), Rc_g.data(), frame_count, ric.data(), sv, device);
    if (status != LFVIO_OK) return false;
    if (frame_count >= window_size && sv[2] > 0.25) {  // ric_cov(1) > 0.25 (:62-68)
        calib_ric_result = ric;
        return true;
    }
    return false;
}

bool Estimator::relativePose(Mat3 &relative_R, Vec3 &relative_T, int &l) {
    for (int i = 0; i < WINDOW_SIZE; i++) {  // estimator.cpp:459-483
        const auto corres = f_manager.getCorresponding(i, WINDOW_SIZE);
        if (corres.size() <= 20) continue;
        double sum_parallax = 0;
        std::vector<double> flat;
        flat.reserve(6 * corres.size());
        for (const auto &c : corres) {
            const double dx = c.first[0] / c.first[2] - c.second[0] / c.second[2], dy = c.first[1] / c.first[2] - c.second[1] / c.second[2];
            sum_parallax += std::sqrt(dx * dx + dy * dy);
            flat.insert(flat.end(), c.first.begin(), c.first.end());
            flat.insert(flat.end(), c.second.begin(), c.second.end());
        }
        const double average_parallax = 1.0 * sum_parallax / (int)corres.size();
        if (!(average_parallax * 160 > 30)) continue;
        int inliers = 0;
        last_status = lfvio_solve_relative_rt(flat.data(), (int)corres.size(), ransac_seed, init_attempts * 64u + (unsigned)i, relative_R.data(),
                                              relative_T.data(), &inliers, nullptr, m_device);
        if (last_status != LFVIO_OK) return false;
        if (inliers > 12) {
            l = i;
            return true;
        }
    }
    return false;
}

bool Estimator::initialStructure() {
    init_attempts++;
    init_fail_stage = 0;
    if (flushIMU() != LFVIO_OK) return false;
    if ((last_status = integrateImageFrames(false, nullptr)) != LFVIO_OK) return false;
    // global sfm (estimator.cpp:250-288)
    const int F = frame_count + 1;
    std::vector<int> obs_first{0}, obs_frame, ids;
    std::vector<double> obs_pts;
    for (auto &it : f_manager.feature) {
        int imu_j = it.start_frame - 1;
        for (auto &o : it.feature_per_frame) {
            imu_j++;
            obs_frame.push_back(imu_j);
            obs_pts.insert(obs_pts.end(), o.point.begin(), o.point.end());
        }
        obs_first.push_back((int)obs_frame.size());
        ids.push_back(it.feature_id);
    }
    Mat3 relative_R{};
    Vec3 relative_T{};
    int l = 0;
    if (!relativePose(relative_R, relative_T, l)) {
        init_fail_stage = 1;
        return false;
    }
    const int L = (int)ids.size();
    std::vector<double> Q(4 * F), T(3 * F), points(3 * (size_t)L);
    std::vector<unsigned char> state(L);
    last_status = lfvio_global_sfm(F, l, relative_R.data(), relative_T.data(), obs_first.data(), obs_frame.data(), obs_pts.data(), L, 1,
                                   Q.data(), T.data(), points.data(), state.data(), &sfm_summary, m_device);
    if (last_status != LFVIO_OK || !sfm_summary.ok) {
        marginalization_flag = MARGIN_OLD;
        init_fail_stage = 2;
        return false;
    }
    std::map<int, Vec3> sfm_tracked_points;
    for (int j = 0; j < L; j++)
        if (state[j]) sfm_tracked_points[ids[j]] = Vec3{points[3 * j], points[3 * j + 1], points[3 * j + 2]};
    // solve pnp for all frame (:290-369): key frames take the SfM pose, the others go to ONE batched EPnP call
    const Mat3 RIC_T = transpose(ric[0]);
    std::vector<ImageFrame *> pnp_frames;
    std::vector<int> first{0};
    std::vector<double> pws, us;
    {
        int i = 0;
        for (auto &kv : all_image_frame) {
            ImageFrame &f = kv.second;
            if (kv.first == Headers[i]) {
                f.is_key_frame = true;
                f.R = matmul(quat_to_rot(Q[4 * i], Q[4 * i + 1], Q[4 * i + 2], Q[4 * i + 3], false), RIC_T);
                f.T = Vec3{T[3 * i], T[3 * i + 1], T[3 * i + 2]};
                i++;
                continue;
            }
            if (kv.first > Headers[i]) i++;
            f.is_key_frame = false;
            int n = 0;
            for (auto &id_pts : f.points) {
                const auto it = sfm_tracked_points.find(id_pts.first);
                if (it == sfm_tracked_points.end()) continue;
                for (auto &i_p : id_pts.second) {
                    us.insert(us.end(), i_p.second.begin(), i_p.second.begin() + 3);
                    pws.insert(pws.end(), it->second.begin(), it->second.end());
                    n++;
                }
            }
            if (n < 6) {  // :333-338
                init_fail_stage = 3;
                return false;
            }
            first.push_back(first.back() + n);
            pnp_frames.push_back(&f);
        }
    }
    if (!pnp_frames.empty()) {
        const int np = (int)pnp_frames.size();
        std::vector<double> Rp(9 * (size_t)np), Tp(3 * (size_t)np);
        last_status = lfvio_epnp(pws.data(), us.data(), first.data(), np, Rp.data(), Tp.data(), nullptr, m_device);
        if (last_status != LFVIO_OK) return false;
        for (int p = 0; p < np; p++) {  // :351-356
            Mat3 tmp{};
            for (int k = 0; k < 9; k++) tmp[k] = Rp[9 * p + k];
            const Mat3 R_pnp = transpose(tmp);
            const Vec3 t = matvec(R_pnp, Vec3{-Tp[3 * p], -Tp[3 * p + 1], -Tp[3 * p + 2]});
            pnp_frames[p]->R = matmul(R_pnp, RIC_T);
            pnp_frames[p]->T = t;
        }
    }
    if (visualInitialAlign()) return true;
    init_fail_stage = 4;
    return false;
}

bool Estimator::visualInitialAlign() {
    // VisualIMUAlignment (initial_aligment.cpp:207-215)
    const int n = (int)all_image_frame.size();
    if (n < 2 || n > LFVIO_INIT_MAX_FRAMES) return false;
    std::vector<double> R(9 * (size_t)n), T(3 * (size_t)n), dq, Jrg, dp, dv, dts;
    {
        int i = 0;
        for (auto &kv : all_image_frame) {
            for (int k = 0; k < 9; k++) R[9 * i + k] = kv.second.R[k];
            for (int k = 0; k < 3; k++) T[3 * i + k] = kv.second.T[k];
            if (i > 0) {
                if (!kv.second.has_pre) return false;
                const IntegrationBase &p = kv.second.pre;
                dq.insert(dq.end(), p.delta_q, p.delta_q + 4);
                for (int r = 0; r < 3; r++)
                    for (int c = 0; c < 3; c++) Jrg.push_back(p.jacobian[15 * (3 + r) + 12 + c]);  // block<3,3>(O_R, O_BG)
            }
            i++;
        }
    }
    Vec3 delta_bg{};
    if ((last_status = lfvio_solve_gyro_bias(R.data(), dq.data(), Jrg.data(), n, delta_bg.data(), m_device)) != LFVIO_OK) return false;
    for (int i = 0; i <= WINDOW_SIZE; i++)
        for (int k = 0; k < 3; k++) Bgs[i][k] += delta_bg[k];
    if ((last_status = integrateImageFrames(true, &Bgs[0])) != LFVIO_OK) return false;
    {
        int i = 0;
        for (auto &kv : all_image_frame) {
            if (i > 0) {
                const IntegrationBase &p = kv.second.pre;
                dp.insert(dp.end(), p.delta_p, p.delta_p + 3);
                dv.insert(dv.end(), p.delta_v, p.delta_v + 3);
                dts.push_back(p.sum_dt);
            }
            i++;
        }
    }
    std::vector<double> x(3 * (size_t)n + 4);
    int ok = 0;
    const double g_norm = norm3(G);
    last_status = lfvio_linear_alignment(R.data(), T.data(), dp.data(), dv.data(), dts.data(), n, tic[0].data(), g_norm, g.data(), x.data(),
                                         &ok, m_device);
    if (last_status != LFVIO_OK || !ok) return false;
    // change state (estimator.cpp:390-455)
    for (int i = 0; i <= frame_count; i++) {
        ImageFrame &f = all_image_frame[Headers[i]];
        Ps[i] = f.T;
        Rs[i] = f.R;
        f.is_key_frame = true;
    }
    f_manager.clearDepth();
    const Vec3 TIC_TMP[1] = {Vec3{0, 0, 0}};
    if ((last_status = f_manager.triangulate(Ps, Rs, TIC_TMP, ric, m_device)) != LFVIO_OK) return false;
    const double s = x[3 * (size_t)n + 2];
    for (int i = 0; i <= WINDOW_SIZE; i++)
        if (repropagate(i, Vec3{0, 0, 0}, Bgs[i]) != LFVIO_OK) return false;
    {
        const Vec3 r0 = matvec(Rs[0], tic[0]);
        const Vec3 base{s * Ps[0][0] - r0[0], s * Ps[0][1] - r0[1], s * Ps[0][2] - r0[2]};
        for (int i = frame_count; i >= 0; i--) {
            const Vec3 ri = matvec(Rs[i], tic[0]);
            for (int k = 0; k < 3; k++) Ps[i][k] = s * Ps[i][k] - ri[k] - base[k];
        }
    }
    int kv = -1;
    for (auto &fr : all_image_frame)
        if (fr.second.is_key_frame) {
            kv++;  // (x.segment<3>(kv * 3): the reference indexes the velocities by key-frame count, :427-435)
            Vs[kv] = matvec(fr.second.R, Vec3{x[3 * kv], x[3 * kv + 1], x[3 * kv + 2]});
        }
    for (auto &it : f_manager.feature) {
        it.used_num = (int)it.feature_per_frame.size();
        if (!(it.used_num >= 2 && it.start_frame < WINDOW_SIZE - 2)) continue;
        it.estimated_depth *= s;
    }
    Mat3 R0 = g2R(g);
    const double yaw = R2ypr(matmul(R0, Rs[0]))[0];
    R0 = matmul(ypr2R(Vec3{-yaw, 0, 0}), R0);
    g = matvec(R0, g);
    for (int i = 0; i <= frame_count; i++) {
        Ps[i] = matvec(R0, Ps[i]);
        Rs[i] = matmul(R0, Rs[i]);
        Vs[i] = matvec(R0, Vs[i]);
    }
    return true;
}

}  // namespace lfvio_host

// ---- C entry points for the Python tests ---------------------------------------------------------------------------
extern "C" void *lfvio_host_ex_rotation_create(int window_size, unsigned long long seed, int device) {
    auto *e = new lfvio_host::InitialEXRotation();
    e->window_size = window_size;
    e->ransac_seed = seed;
    e->device = device;
    return e;
}
extern "C" void lfvio_host_ex_rotation_destroy(void *h) { delete static_cast<lfvio_host::InitialEXRotation *>(h); }
// corres: n x 6, dq: (w, x, y, z); ric_out (9) receives calib_ric_result when the function returns true (1); sv_out (4)
extern "C" int lfvio_host_ex_rotation_step(void *h, const double *corres, int n, const double *dq, double *ric_out, double *sv_out) {
    using namespace lfvio_host;
    auto *e = static_cast<InitialEXRotation *>(h);
    std::vector<std::pair<Vec3, Vec3>> c;
    for (int i = 0; i < n; i++)
        c.emplace_back(Vec3{corres[6 * i], corres[6 * i + 1], corres[6 * i + 2]}, Vec3{corres[6 * i + 3], corres[6 * i + 4], corres[6 * i + 5]});
    Mat3 out{};
    const bool ok = e->CalibrationExRotation(c, {dq[0], dq[1], dq[2], dq[3]}, out);
    if (e->status != LFVIO_OK) return -1;
    for (int k = 0; k < 4; k++) sv_out[k] = e->sv[k];
    if (ok)
        for (int k = 0; k < 9; k++) ric_out[k] = out[k];
    return ok ? 1 : 0;
}
