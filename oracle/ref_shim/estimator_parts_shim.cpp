// TEST INFRASTRUCTURE ONLY (oracle/): C entry points over more of the reference's own code, compiled UNMODIFIED from
// /root/reference/vins_estimator/src/{feature_manager.cpp, initial/initial_aligment.cpp} against oracle/ref_stubs:
//   * class FeatureManager (addFeatureCheckParallax, getCorresponding, triangulate, removeBack / removeFront /
//     removeBackShiftDepth / removeFailures, get/setDepthVector)               feature_manager.cpp:1-369
//   * solveGyroscopeBias / LinearAlignment / RefineGravity / VisualIMUAlignment  initial/initial_aligment.cpp:3-216
// The entry points mirror the test hooks of the C++ host mirror (lfvio_host_fm_*, lf-vio_b200/host/estimator.cpp) so one test
// drives both. WINDOW_SIZE is the reference's compile-time 10.
#include <cstring>
#include <iostream>
#include <map>

#include "feature_manager.h"
#include "initial/initial_alignment.h"
#include "lfvio.h"

extern double ACC_N, ACC_W, GYR_N, GYR_W, MIN_PARALLAX;
// defined in initial_aligment.cpp, declared only there (initial_alignment.h exports VisualIMUAlignment alone)
void solveGyroscopeBias(std::map<double, ImageFrame> &all_image_frame, Eigen::Vector3d *Bgs);
bool LinearAlignment(std::map<double, ImageFrame> &all_image_frame, Eigen::Vector3d &g, Eigen::VectorXd &x);

namespace {
struct FM {
    Eigen::Matrix3d Rs[WINDOW_SIZE + 1];
    FeatureManager f;
    FM() : f(Rs) {
        for (auto &R : Rs) R.setIdentity();
    }
};
Eigen::Matrix3d mat3(const double *p) {
    Eigen::Matrix3d R;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R(i, j) = p[3 * i + j];
    return R;
}
}  // namespace

extern "C" {

void *ref_fm_create(double min_parallax) {
    MIN_PARALLAX = min_parallax;
    return new FM();
}
void ref_fm_destroy(void *h) { delete static_cast<FM *>(h); }
// points8: n x 8 = [bearing 3 | u v | velocity 3] (the Matrix<double,8,1> of estimator_node.cpp:292-312)
int ref_fm_add(void *h, int frame_count, const int *ids, const double *points8, int n, double td) {
    std::map<int, std::vector<std::pair<int, Eigen::Matrix<double, 8, 1>>>> image;
    for (int i = 0; i < n; i++) {
        Eigen::Matrix<double, 8, 1> p;
        for (int k = 0; k < 8; k++) p(k) = points8[8 * i + k];
        image[ids[i]].emplace_back(0, p);
    }
    return static_cast<FM *>(h)->f.addFeatureCheckParallax(frame_count, image, td) ? 1 : 0;
}
int ref_fm_last_track_num(void *h) { return static_cast<FM *>(h)->f.last_track_num; }
void ref_fm_set_rs(void *h, const double *Rs) {  // (WINDOW_SIZE + 1) x 9: the Rs[] the manager reads through its pointer
    for (int i = 0; i <= WINDOW_SIZE; i++) static_cast<FM *>(h)->Rs[i] = mat3(Rs + 9 * i);
}
int ref_fm_corresponding(void *h, int l, int r, double *out, int cap) {
    const auto c = static_cast<FM *>(h)->f.getCorresponding(l, r);
    const int n = (int)c.size() < cap ? (int)c.size() : cap;
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) out[6 * i + k] = c[i].first(k), out[6 * i + 3 + k] = c[i].second(k);
    return (int)c.size();
}
// op: 0 removeBack, 1 removeFront(arg), 2 removeBackShiftDepth(pose: marg_R 9 | marg_P 3 | new_R 9 | new_P 3), 3 removeFailures
void ref_fm_slide(void *h, int op, int arg, const double *pose) {
    FeatureManager &f = static_cast<FM *>(h)->f;
    if (op == 0)
        f.removeBack();
    else if (op == 1)
        f.removeFront(arg);
    else if (op == 2)
        f.removeBackShiftDepth(mat3(pose), Eigen::Vector3d(pose[9], pose[10], pose[11]), mat3(pose + 12),
                               Eigen::Vector3d(pose[21], pose[22], pose[23]));
    else
        f.removeFailures();
}
// per feature [id, start_frame, n_obs, solve_flag] and estimated_depth; set_depth != NULL first assigns estimated_depth in
// list order. Returns the number of features; *count_out = getFeatureCount().
int ref_fm_snapshot(void *h, int *meta4, double *depth, int cap, const double *set_depth, int *count_out) {
    FeatureManager &f = static_cast<FM *>(h)->f;
    int i = 0;
    for (auto &it : f.feature) {
        if (set_depth) it.estimated_depth = set_depth[i];
        if (i < cap) {
            meta4[4 * i] = it.feature_id, meta4[4 * i + 1] = it.start_frame, meta4[4 * i + 2] = (int)it.feature_per_frame.size();
            meta4[4 * i + 3] = it.solve_flag;
            depth[i] = it.estimated_depth;
        }
        i++;
    }
    if (count_out) *count_out = f.getFeatureCount();
    return i;
}
// FeatureManager::triangulate(Ps, tic, ric) with Rs taken from ref_fm_set_rs
void ref_fm_triangulate(void *h, const double *Ps, const double *tic, const double *ric) {
    FM *m = static_cast<FM *>(h);
    Eigen::Vector3d P[WINDOW_SIZE + 1], t[1]{Eigen::Vector3d(tic[0], tic[1], tic[2])};
    Eigen::Matrix3d r[1]{mat3(ric)};
    for (int i = 0; i <= WINDOW_SIZE; i++) P[i] = Eigen::Vector3d(Ps[3 * i], Ps[3 * i + 1], Ps[3 * i + 2]);
    m->f.setRic(r);
    m->f.triangulate(P, t, r);
}
int ref_fm_get_depth_vector(void *h, double *out, int cap) {  // getDepthVector(): 1 / estimated_depth of the used features
    Eigen::VectorXd v = static_cast<FM *>(h)->f.getDepthVector();
    for (int i = 0; i < (int)v.size() && i < cap; i++) out[i] = v(i);
    return (int)v.size();
}
void ref_fm_set_depth(void *h, const double *x, int n) {  // setDepth(x): estimated_depth = 1 / x, solve_flag by sign
    Eigen::VectorXd v(n);
    for (int i = 0; i < n; i++) v(i) = x[i];
    static_cast<FM *>(h)->f.setDepth(v);
}

// VisualIMUAlignment pieces on n frames (all_image_frame): R (n x 9), T (n x 3) = ImageFrame::R / T; intervals[i] (n - 1 of
// them) + the sample arrays build frame i+1's pre_integration by push_back. mode 0: solveGyroscopeBias only; 1: LinearAlignment
// only; 2: VisualIMUAlignment. bg: in = Bgs[*] before (all equal), out = Bgs[0] after. x: caller provides 3 n + 4 doubles.
int ref_visual_imu_alignment(int n, const double *R, const double *T, const lfvio_imu_interval *iv, const double *dt,
                             const double *acc, const double *gyr, const lfvio_imu_noise *noise, const double *tic, double g_norm,
                             int mode, double *bg, double *g_out, double *x_out, int *x_size, int *ok, lfvio_imu_factor *pre_out) {
    ACC_N = noise->acc_n, GYR_N = noise->gyr_n, ACC_W = noise->acc_w, GYR_W = noise->gyr_w;
    G = Eigen::Vector3d(0, 0, g_norm);
    TIC.assign(1, Eigen::Vector3d(tic[0], tic[1], tic[2]));
    std::map<double, ImageFrame> frames;
    std::vector<IntegrationBase *> owned;
    const Eigen::Vector3d bg0(bg[0], bg[1], bg[2]);
    for (int i = 0; i < n; i++) {
        ImageFrame f;
        f.t = (double)i;
        f.R = mat3(R + 9 * i);
        f.T = Eigen::Vector3d(T[3 * i], T[3 * i + 1], T[3 * i + 2]);
        f.pre_integration = nullptr;
        if (i > 0) {
            const lfvio_imu_interval &v = iv[i - 1];
            IntegrationBase *ib = new IntegrationBase(Eigen::Vector3d(v.acc_0[0], v.acc_0[1], v.acc_0[2]),
                                                      Eigen::Vector3d(v.gyr_0[0], v.gyr_0[1], v.gyr_0[2]),
                                                      Eigen::Vector3d(v.ba[0], v.ba[1], v.ba[2]), bg0);
            for (int k = v.first; k < v.first + v.count; k++)
                ib->push_back(dt[k], Eigen::Vector3d(acc[3 * k], acc[3 * k + 1], acc[3 * k + 2]),
                              Eigen::Vector3d(gyr[3 * k], gyr[3 * k + 1], gyr[3 * k + 2]));
            f.pre_integration = ib;
            owned.push_back(ib);
        }
        frames[f.t] = f;
    }
    Eigen::Vector3d Bgs[WINDOW_SIZE + 1];
    for (auto &b : Bgs) b = bg0;
    Eigen::Vector3d g = Eigen::Vector3d::Zero();
    Eigen::VectorXd x;
    std::cout.setstate(std::ios_base::failbit);  // LinearAlignment prints its blocks to stdout
    bool r = true;
    if (mode == 0)
        solveGyroscopeBias(frames, Bgs);
    else if (mode == 1)
        r = LinearAlignment(frames, g, x);
    else
        r = VisualIMUAlignment(frames, Bgs, g, x);
    std::cout.clear();
    for (int k = 0; k < 3; k++) bg[k] = Bgs[0](k), g_out[k] = g(k);
    *x_size = (int)x.size();
    for (int k = 0; k < (int)x.size(); k++) x_out[k] = x(k);
    *ok = r ? 1 : 0;
    if (pre_out)
        for (int i = 0; i + 1 < n; i++) {
            IntegrationBase &ib = *owned[i];
            lfvio_imu_factor &f = pre_out[i];
            std::memset(&f, 0, sizeof(f));
            f.sum_dt = ib.sum_dt;
            for (int a = 0; a < 3; a++) {
                f.delta_p[a] = ib.delta_p(a), f.delta_v[a] = ib.delta_v(a);
                f.linearized_ba[a] = ib.linearized_ba(a), f.linearized_bg[a] = ib.linearized_bg(a);
            }
            f.delta_q[0] = ib.delta_q.w(), f.delta_q[1] = ib.delta_q.x(), f.delta_q[2] = ib.delta_q.y(), f.delta_q[3] = ib.delta_q.z();
            for (int a = 0; a < 15; a++)
                for (int b = 0; b < 15; b++) f.jacobian[15 * a + b] = ib.jacobian(a, b), f.covariance[15 * a + b] = ib.covariance(a, b);
        }
    for (IntegrationBase *ib : owned) delete ib;
    return 0;
}

}  // extern "C"
