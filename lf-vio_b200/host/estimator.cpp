// Estimator::optimization() host mirror: packs the problem, calls the device solver, applies the gauge fix,
// builds the next prior.  No arithmetic of the hot path happens here (only the 3x3 gauge re-anchoring of a15).
#include "estimator.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <cstring>

namespace lfvio_host {

int ESTIMATE_TD = 1, ESTIMATE_EXTRINSIC = 1, NUM_ITERATIONS = 8, EST_ROW = 960;
double TR = 0.0, TD = -0.008, FOCAL_LENGTH = 160.0, MIN_PARALLAX = 10.0 / 160.0;
Vec3 G{0.0, 0.0, 9.81007};

// LFVIO_DEBUG_HOST=1: where the host time of processImage() goes (printed when an Estimator is destroyed)
namespace {
struct HostTimers {
    const bool on = getenv("LFVIO_DEBUG_HOST") != nullptr;
    double acc[12] = {0};
    long calls[12] = {0};
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
    }
    void report() const {
        if (!on) return;
        static const char *nm[12] = {"addFeatureCheckParallax", "flushIMU (pre-integration)", "triangulate", "vector2double + buildFactors", "lfvio_window_optimize(_full)", "double2vector", "lfvio_window_marginalize", "slideWindow + removeFailures", "", "", "", ""};
        for (int i = 0; i < 8; i++)
            if (calls[i]) fprintf(stderr, "Estimator host %-30s %9.1f us x %ld\n", nm[i], acc[i] / calls[i], calls[i]);
    }
} g_ht;
struct Lap {
    int i;
    double t0;
    explicit Lap(int i_) : i(i_), t0(g_ht.on ? HostTimers::now() : 0.0) {}
    ~Lap() {
        if (g_ht.on) {
            g_ht.acc[i] += HostTimers::now() - t0;
            g_ht.calls[i]++;
        }
    }
};
}  // namespace

namespace detail {

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

}  // namespace detail
using namespace detail;

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

bool FeatureManager::addFeatureCheckParallax(int frame_count, const ImageFeatures &image, double td) {
    double parallax_sum = 0;
    int parallax_num = 0;
    last_track_num = 0;
    for (const auto &id_pts : image) {
        const std::array<double, 8> &p = id_pts.second[0].second;
        FeaturePerFrame f;  // FeaturePerFrame(_point, td)  (feature_manager.h:21-35)
        f.point = Vec3{p[0], p[1], p[2]};
        f.uv = {p[3], p[4]};
        f.velocity = Vec3{p[5], p[6], p[7]};
        f.cur_td = td;
        const int feature_id = id_pts.first;
        auto it = feature.begin();
        for (; it != feature.end(); ++it)
            if (it->feature_id == feature_id) break;
        if (it == feature.end()) {
            FeaturePerId fp;
            fp.feature_id = feature_id;
            fp.start_frame = frame_count;
            fp.feature_per_frame.push_back(f);
            feature.push_back(fp);
        } else {
            it->feature_per_frame.push_back(f);
            last_track_num++;
        }
    }
    if (frame_count < 2 || last_track_num < 20) return true;
    for (const auto &it : feature) {
        if (it.start_frame <= frame_count - 2 && it.start_frame + (int)it.feature_per_frame.size() - 1 >= frame_count - 1) {
            const Vec3 &pi = it.feature_per_frame[frame_count - 2 - it.start_frame].point;
            const Vec3 &pj = it.feature_per_frame[frame_count - 1 - it.start_frame].point;
            parallax_sum += std::acos(pi[0] * pj[0] + pi[1] * pj[1] + pi[2] * pj[2]) * 10;  // compensatedParallax2
            parallax_num++;
        }
    }
    if (parallax_num == 0) return true;
    return parallax_sum / parallax_num >= MIN_PARALLAX;
}

void FeatureManager::removeBackShiftDepth(const Mat3 &marg_R, const Vec3 &marg_P, const Mat3 &new_R, const Vec3 &new_P) {
    for (auto it = feature.begin(), it_next = feature.begin(); it != feature.end(); it = it_next) {
        it_next++;
        if (it->start_frame != 0) {
            it->start_frame--;
            continue;
        }
        const Vec3 uv_i = it->feature_per_frame[0].point;
        it->feature_per_frame.erase(it->feature_per_frame.begin());
        if (it->feature_per_frame.size() < 2) {
            feature.erase(it);
            continue;
        }
        const Vec3 pts_i{uv_i[0] * it->estimated_depth, uv_i[1] * it->estimated_depth, uv_i[2] * it->estimated_depth};
        const Vec3 w = matvec(marg_R, pts_i);
        const Vec3 d{w[0] + marg_P[0] - new_P[0], w[1] + marg_P[1] - new_P[1], w[2] + marg_P[2] - new_P[2]};
        const Vec3 pts_j = matvec(transpose(new_R), d);
        const double dep_j = std::sqrt(pts_j[0] * pts_j[0] + pts_j[1] * pts_j[1] + pts_j[2] * pts_j[2]);
        it->estimated_depth = dep_j > 0 ? dep_j : init_depth;
    }
}

void FeatureManager::removeBack() {
    for (auto it = feature.begin(), it_next = feature.begin(); it != feature.end(); it = it_next) {
        it_next++;
        if (it->start_frame != 0) {
            it->start_frame--;
        } else {
            it->feature_per_frame.erase(it->feature_per_frame.begin());
            if (it->feature_per_frame.empty()) feature.erase(it);
        }
    }
}

void FeatureManager::removeFront(int frame_count) {
    for (auto it = feature.begin(), it_next = feature.begin(); it != feature.end(); it = it_next) {
        it_next++;
        if (it->start_frame == frame_count) {
            it->start_frame--;
        } else {
            const int j = window_size - 1 - it->start_frame;
            if (it->start_frame + (int)it->feature_per_frame.size() - 1 < frame_count - 1) continue;
            it->feature_per_frame.erase(it->feature_per_frame.begin() + j);
            if (it->feature_per_frame.empty()) feature.erase(it);
        }
    }
}

void FeatureManager::removeFailures() {
    for (auto it = feature.begin(), it_next = feature.begin(); it != feature.end(); it = it_next) {
        it_next++;
        if (it->solve_flag == 2) feature.erase(it);
    }
}

int FeatureManager::triangulate(const std::vector<Vec3> &Ps, const std::vector<Mat3> &Rs, const Vec3 tic[], const Mat3 ric[], int device) {
    std::vector<int> start, first{0};
    std::vector<double> pts, depth;
    for (auto &it : feature) {
        it.used_num = (int)it.feature_per_frame.size();
        start.push_back(it.start_frame);
        for (const auto &o : it.feature_per_frame) pts.insert(pts.end(), o.point.begin(), o.point.end());
        first.push_back((int)(pts.size() / 3));
        depth.push_back(it.estimated_depth);
    }
    if (start.empty()) return LFVIO_OK;
    std::vector<double> P, R;
    for (const auto &p : Ps) P.insert(P.end(), p.begin(), p.end());
    for (const auto &r : Rs) R.insert(R.end(), r.begin(), r.end());
    const int st = lfvio_triangulate(P.data(), R.data(), (int)Ps.size(), tic[0].data(), ric[0].data(), start.data(), first.data(),
                                     pts.data(), (int)start.size(), init_depth, depth.data(), device);
    if (st != LFVIO_OK) return st;
    size_t i = 0;
    for (auto &it : feature) it.estimated_depth = depth[i++];
    return LFVIO_OK;
}

// ---- Estimator ---------------------------------------------------------------------------------------
Estimator::Estimator(int window_size, int max_features, int device)
    : WINDOW_SIZE(window_size), Ps(window_size + 1), Vs(window_size + 1), Bas(window_size + 1), Bgs(window_size + 1),
      Rs(window_size + 1, Mat3{1, 0, 0, 0, 1, 0, 0, 0, 1}), pre_integrations(window_size + 1, nullptr),
      dt_buf(window_size + 1), linear_acceleration_buf(window_size + 1), angular_velocity_buf(window_size + 1),
      para_Pose(7 * (window_size + 1)), para_SpeedBias(9 * (window_size + 1)), para_Feature(max_features), m_device(device),
      m_iv(window_size + 1), m_pre(window_size + 1) {
    tic[0] = Vec3{0, 0, 0};
    ric[0] = Mat3{1, 0, 0, 0, 1, 0, 0, 0, 1};
    f_manager.window_size = window_size;
    m_max_features = max_features;
    last_status = lfvio_window_create(window_size, max_features, 16 * max_features, device, &m_window);
}
Estimator::~Estimator() {
    g_ht.report();
    if (m_store) lfvio_fm_destroy(m_store);
    if (m_window) lfvio_window_destroy(m_window);
}

// ---- FeatureManager on the device ---------------------------------------------------------------------
// The host-side list (what the initialisation worked on) becomes the content of the device store.
int Estimator::storeLoad() {
    if (!m_store) {
        const int st = lfvio_fm_create(WINDOW_SIZE, 2 * m_max_features, 16 * m_max_features, MIN_PARALLAX, f_manager.init_depth, m_device, &m_store);
        if (st != LFVIO_OK) return st;
    }
    std::vector<int> ids, start, first{0}, flag;
    std::vector<double> depth, obs;
    for (const auto &it : f_manager.feature) {
        ids.push_back(it.feature_id);
        start.push_back(it.start_frame);
        flag.push_back(it.solve_flag);
        depth.push_back(it.estimated_depth);
        for (const auto &o : it.feature_per_frame) {
            const double rec[9] = {o.point[0], o.point[1], o.point[2], o.velocity[0], o.velocity[1], o.velocity[2], o.uv[0], o.uv[1], o.cur_td};
            obs.insert(obs.end(), rec, rec + 9);
        }
        first.push_back((int)(obs.size() / 9));
    }
    const int st = lfvio_fm_load(m_store, (int)ids.size(), ids.data(), start.data(), first.data(), depth.data(), flag.data(), obs.data());
    if (st == LFVIO_OK) m_store_live = true;
    return st;
}
int Estimator::storeTriangulate() {
    std::vector<double> P, R;
    for (const auto &p : Ps) P.insert(P.end(), p.begin(), p.end());
    for (const auto &r : Rs) R.insert(R.end(), r.begin(), r.end());
    return lfvio_fm_triangulate(m_store, P.data(), R.data(), tic[0].data(), ric[0].data());
}
// The two-call path (after a detected failure the gauge is anchored on the host) needs the list on the host after all.
int Estimator::storeDownloadFactors() {
    int R = 0, L = 0;
    int st = lfvio_fm_build_factors(m_store, &R, &L);
    if (st != LFVIO_OK) return st;
    m_proj.resize(R > 0 ? R : 1);
    std::vector<int> lm_first(L + 1);
    if ((int)para_Feature.size() < L) para_Feature.resize(L);
    st = lfvio_fm_download_factors(m_store, m_proj.data(), (int)m_proj.size(), lm_first.data(), para_Feature.data(), L);
    m_proj.resize(R);
    m_store_L = L;
    return st;
}
int Estimator::storeSize() {
    if (!m_store_live) return (int)f_manager.feature.size();
    std::vector<int> meta(3 * 2 * (size_t)m_max_features);
    std::vector<double> depth(2 * (size_t)m_max_features);
    return lfvio_fm_snapshot(m_store, meta.data(), depth.data(), 2 * m_max_features);
}
const char *Estimator::lastError() const { return lfvio_last_error(); }

void Estimator::processIMU(double dt, const Vec3 &linear_acceleration, const Vec3 &angular_velocity) {
    if (!first_imu) {
        first_imu = true;
        acc_0 = linear_acceleration;
        gyr_0 = angular_velocity;
    }
    const int j = frame_count;
    Interval &iv = m_iv[j];
    if (!iv.open) {  // new IntegrationBase{acc_0, gyr_0, Bas[frame_count], Bgs[frame_count]}  (estimator.cpp:95-98)
        iv.open = true;
        iv.acc_start = acc_0;
        iv.gyr_start = gyr_0;
        iv.ba = Bas[j];
        iv.bg = Bgs[j];
        iv.P = Ps[j];
        iv.V = Vs[j];
        iv.R = Rs[j];
        iv.dirty = true;
        iv.predict = true;
    }
    if (j != 0) {
        dt_buf[j].push_back(dt);
        linear_acceleration_buf[j].push_back(linear_acceleration);
        angular_velocity_buf[j].push_back(angular_velocity);
        iv.dirty = true;
        if (m_tmp_valid) {  // tmp_pre_integration->push_back (estimator.cpp:103)
            if (m_tmp.dt.empty()) {  // (acc_0 / gyr_0 are still the values its constructor saw, unless the caller overrode them)
                m_tmp.acc_start = acc_0;
                m_tmp.gyr_start = gyr_0;
            }
            m_tmp.dt.push_back(dt);
            m_tmp.acc.push_back(linear_acceleration);
            m_tmp.gyr.push_back(angular_velocity);
        }
    }
    acc_0 = linear_acceleration;
    gyr_0 = angular_velocity;
}

int Estimator::integrate(int j, bool propagate) {
    Interval &iv = m_iv[j];
    const int n = (int)dt_buf[j].size();
    lfvio_imu_interval I{};
    I.first = 0;
    I.count = n;
    I.frame_i = j - 1;
    I.propagate = propagate ? 1 : 0;
    for (int k = 0; k < 3; k++) {
        I.acc_0[k] = iv.acc_start[k];
        I.gyr_0[k] = iv.gyr_start[k];
        I.ba[k] = iv.ba[k];
        I.bg[k] = iv.bg[k];
        I.P[k] = iv.P[k];
        I.V[k] = iv.V[k];
    }
    for (int k = 0; k < 9; k++) I.R[k] = iv.R[k];
    std::vector<double> acc(3 * (size_t)n + 3), gyr(3 * (size_t)n + 3);
    for (int k = 0; k < n; k++)
        for (int a = 0; a < 3; a++) {
            acc[3 * k + a] = linear_acceleration_buf[j][k][a];
            gyr[3 * k + a] = angular_velocity_buf[j][k][a];
        }
    double state[15];
    // One gravity source for every propagation (processIMU uses the member g, estimator.cpp:118-131; EstimatorNode::predict too):
    // the aligned gravity once the estimator runs, the configured G before (and after a state injection / clearState, which set g = G).
    const Vec3 &grav = (solver_flag == NON_LINEAR) ? g : G;
    const int st = lfvio_preintegrate(&I, 1, dt_buf[j].data(), acc.data(), gyr.data(), n, &imu_noise, grav.data(), &m_pre[j], state, m_device);
    if (st != LFVIO_OK) return st;
    pre_integrations[j] = &m_pre[j];
    if (propagate) {
        Ps[j] = Vec3{state[0], state[1], state[2]};
        Vs[j] = Vec3{state[3], state[4], state[5]};
        for (int k = 0; k < 9; k++) Rs[j][k] = state[6 + k];
    }
    iv.dirty = false;
    return LFVIO_OK;
}

int Estimator::flushIMU() {
    for (int j = 0; j <= WINDOW_SIZE; j++)
        if (m_iv[j].open && m_iv[j].dirty) {
            last_status = integrate(j, m_iv[j].predict);
            m_iv[j].predict = false;
            if (last_status != LFVIO_OK) return last_status;
        }
    return LFVIO_OK;
}

int Estimator::repropagate(int j, const Vec3 &ba, const Vec3 &bg) {
    if (!m_iv[j].open) return LFVIO_OK;
    m_iv[j].ba = ba;
    m_iv[j].bg = bg;
    return last_status = integrate(j, false);  // repropagate() re-integrates the deltas only (integration_base.h:39-53)
}

void Estimator::processImage(const ImageFeatures &image, double header_stamp) {
    if (use_device_store && !m_store_live && solver_flag == NON_LINEAR && (last_status = storeLoad()) != LFVIO_OK) return;
    {
        Lap lap(0);
        if (m_store_live) {
            m_ids.clear();
            m_feat8.clear();
            for (const auto &id_pts : image) {  // std::map: ascending ids
                m_ids.push_back(id_pts.first);
                const std::array<double, 8> &p = id_pts.second[0].second;
                m_feat8.insert(m_feat8.end(), p.begin(), p.end());
            }
            int key = 1;
            last_status = lfvio_fm_add_frame(m_store, frame_count, m_ids.data(), m_feat8.data(), (int)m_ids.size(), td, &key, &f_manager.last_track_num);
            if (last_status != LFVIO_OK) return;
            marginalization_flag = key ? MARGIN_OLD : MARGIN_SECOND_NEW;
        } else {
            marginalization_flag = f_manager.addFeatureCheckParallax(frame_count, image, td) ? MARGIN_OLD : MARGIN_SECOND_NEW;
        }
    }
    if ((int)Headers.size() <= WINDOW_SIZE) Headers.resize(WINDOW_SIZE + 1, 0.0);
    Headers[frame_count] = header_stamp;
    {  // estimator.cpp:136-140: the frame joins all_image_frame with the pre-integration since the previous image
        ImageFrame imageframe = std::move(m_tmp);
        imageframe.points = image;
        imageframe.t = header_stamp;
        imageframe.is_key_frame = false;
        imageframe.has_pre = m_tmp_valid;
        all_image_frame[header_stamp] = std::move(imageframe);
        m_tmp = ImageFrame{};
        m_tmp.acc_start = acc_0;
        m_tmp.gyr_start = gyr_0;
        m_tmp.ba = Bas[frame_count];
        m_tmp.bg = Bgs[frame_count];
        m_tmp_valid = true;
    }
    if (ESTIMATE_EXTRINSIC == 2 && frame_count != 0) {  // estimator.cpp:142-159: calibrate the extrinsic rotation online
        if (flushIMU() != LFVIO_OK) return;
        if (pre_integrations[frame_count]) {
            const auto corres = f_manager.getCorresponding(frame_count - 1, frame_count);
            const double *dq = pre_integrations[frame_count]->delta_q;
            Mat3 calib_ric{};
            initial_ex_rotation.window_size = WINDOW_SIZE;
            initial_ex_rotation.device = m_device;
            if (initial_ex_rotation.CalibrationExRotation(corres, {dq[0], dq[1], dq[2], dq[3]}, calib_ric)) {
                ric[0] = calib_ric;
                RIC0 = calib_ric;
                ESTIMATE_EXTRINSIC = 1;
            }
            if (initial_ex_rotation.status != LFVIO_OK) {
                last_status = initial_ex_rotation.status;
                return;
            }
        }
    }
    if (solver_flag == INITIAL) {
        // With run_initial_structure == false the window fills, then slides until the caller has initialised the states and
        // set solver_flag = NON_LINEAR (what the reference does while its initialisation fails, :177-178).
        if (frame_count == WINDOW_SIZE) {
            bool result = false;
            if (run_initial_structure && ESTIMATE_EXTRINSIC != 2 && (header_stamp - initial_timestamp) > 0.1) {
                result = initialStructure();
                initial_timestamp = header_stamp;
            }
            if (result) {
                solver_flag = NON_LINEAR;
                if (stop_after_initial_structure) return;  // parity hook: expose the states initialStructure leaves
                solveOdometry();
                if (last_status != LFVIO_OK) return;
                slideWindow();
                f_manager.removeFailures();
                last_R = Rs[WINDOW_SIZE];
                last_P = Ps[WINDOW_SIZE];
                last_R0 = Rs[0];
                last_P0 = Ps[0];
            } else {
                if (flushIMU() != LFVIO_OK) return;
                slideWindow();
            }
        } else {
            frame_count++;
        }
        return;
    }
    solveOdometry();
    if (last_status != LFVIO_OK) return;
    if (failureDetection()) {
        failure_occur = true;
        if (run_initial_structure) {  // estimator.cpp:196-204: system reboot; the window refills and initialStructure() runs again
            clearState();              // (this also clears failure_occur, as in the reference)
            setParameter();
            reboot_count++;
        }  // otherwise (states injected by the caller) the reboot is the caller's: failure_occur stays set
        return;
    }
    {
        Lap lap(7);
        slideWindow();
        if (!m_store_live) f_manager.removeFailures();  // (a no-op in LF-VIO: nothing ever sets solve_flag = 2)
    }
    key_poses.assign(Ps.begin(), Ps.end());
    last_R = Rs[WINDOW_SIZE];
    last_P = Ps[WINDOW_SIZE];
    last_R0 = Rs[0];
    last_P0 = Ps[0];
}

void Estimator::pokeCovariance(int j) {
    if (j >= 0 && j < (int)m_pre.size()) memset(m_pre[j].covariance, 0, sizeof(m_pre[j].covariance));
}

void Estimator::setParameter() {
    tic[0] = TIC0;
    ric[0] = RIC0;
    td = TD;
}

void Estimator::clearState() {
    const Mat3 I{1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i <= WINDOW_SIZE; i++) {
        Rs[i] = I;
        Ps[i] = Vs[i] = Bas[i] = Bgs[i] = Vec3{0, 0, 0};
        dt_buf[i].clear();
        linear_acceleration_buf[i].clear();
        angular_velocity_buf[i].clear();
        pre_integrations[i] = nullptr;
        m_iv[i] = Interval{};
    }
    tic[0] = Vec3{0, 0, 0};
    ric[0] = I;
    solver_flag = INITIAL;
    first_imu = false;
    sum_of_back = sum_of_front = 0;
    frame_count = 0;
    initial_timestamp = 0;
    all_image_frame.clear();
    td = TD;
    m_tmp = ImageFrame{};
    m_tmp_valid = false;
    last_marginalization_info = MarginalizationInfo{};
    f_manager.clearState();
    m_store_live = false;  // the next initialisation works on the host-side list again
    failure_occur = false;
    g = G;  // (the reference leaves g as it was; the IMU-rate prediction must not run gravity-free until the next alignment)
}

void Estimator::solveOdometry() {
    if (frame_count < WINDOW_SIZE) return;
    if (solver_flag == NON_LINEAR) {
        {
            Lap lap(1);
            if (flushIMU() != LFVIO_OK) return;  // Ps/Vs/Rs[WINDOW_SIZE] predicted before triangulating against them
        }
        {
            Lap lap(2);
            last_status = m_store_live ? storeTriangulate() : f_manager.triangulate(Ps, Rs, tic, ric, m_device);
        }
        if (last_status != LFVIO_OK) return;
        optimization();
    }
}

bool Estimator::failureDetection() {
    auto norm = [](const Vec3 &v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
    if (norm(Bgs[WINDOW_SIZE]) > 1.0) return true;
    const Vec3 &p = Ps[WINDOW_SIZE];
    if (norm(Vec3{p[0] - last_P[0], p[1] - last_P[1], p[2] - last_P[2]}) > 5) return true;
    if (std::fabs(p[2] - last_P[2]) > 1) return true;
    return false;  // (the little-feature, big-acc-bias and big-angle tests only log in the reference)
}

void Estimator::slideWindow() {
    if (marginalization_flag == MARGIN_OLD) {
        const double t_0 = Headers.empty() ? 0.0 : Headers[0];
        back_R0 = Rs[0];
        back_P0 = Ps[0];
        if (frame_count == WINDOW_SIZE) {
            // estimator.cpp:1051-1066 (`true || solver_flag == INITIAL`): frames up to the old frame 0 leave all_image_frame
            all_image_frame.erase(all_image_frame.begin(), all_image_frame.upper_bound(t_0));
            for (int i = 0; i < WINDOW_SIZE; i++) {
                std::swap(Rs[i], Rs[i + 1]);
                std::swap(m_pre[i], m_pre[i + 1]);
                std::swap(m_iv[i], m_iv[i + 1]);
                dt_buf[i].swap(dt_buf[i + 1]);
                linear_acceleration_buf[i].swap(linear_acceleration_buf[i + 1]);
                angular_velocity_buf[i].swap(angular_velocity_buf[i + 1]);
                Headers[i] = Headers[i + 1];
                std::swap(Ps[i], Ps[i + 1]);
                std::swap(Vs[i], Vs[i + 1]);
                std::swap(Bas[i], Bas[i + 1]);
                std::swap(Bgs[i], Bgs[i + 1]);
            }
            Headers[WINDOW_SIZE] = Headers[WINDOW_SIZE - 1];
            Ps[WINDOW_SIZE] = Ps[WINDOW_SIZE - 1];
            Vs[WINDOW_SIZE] = Vs[WINDOW_SIZE - 1];
            Rs[WINDOW_SIZE] = Rs[WINDOW_SIZE - 1];
            Bas[WINDOW_SIZE] = Bas[WINDOW_SIZE - 1];
            Bgs[WINDOW_SIZE] = Bgs[WINDOW_SIZE - 1];
            // new IntegrationBase{acc_0, gyr_0, Bas[WINDOW_SIZE], Bgs[WINDOW_SIZE]}: opened by the next processIMU
            m_iv[WINDOW_SIZE] = Interval{};
            dt_buf[WINDOW_SIZE].clear();
            linear_acceleration_buf[WINDOW_SIZE].clear();
            angular_velocity_buf[WINDOW_SIZE].clear();
            for (int i = 0; i <= WINDOW_SIZE; i++) pre_integrations[i] = m_iv[i].open ? &m_pre[i] : nullptr;
            slideWindowOld();
        }
    } else if (frame_count == WINDOW_SIZE) {
        const int fc = frame_count;
        for (size_t i = 0; i < dt_buf[fc].size(); i++) {  // the newest interval is appended to the previous one
            dt_buf[fc - 1].push_back(dt_buf[fc][i]);
            linear_acceleration_buf[fc - 1].push_back(linear_acceleration_buf[fc][i]);
            angular_velocity_buf[fc - 1].push_back(angular_velocity_buf[fc][i]);
        }
        if (!dt_buf[fc].empty()) m_iv[fc - 1].dirty = true;  // re-integrated from its own start by the next flushIMU
        Headers[fc - 1] = Headers[fc];
        Ps[fc - 1] = Ps[fc];
        Vs[fc - 1] = Vs[fc];
        Rs[fc - 1] = Rs[fc];
        Bas[fc - 1] = Bas[fc];
        Bgs[fc - 1] = Bgs[fc];
        m_iv[WINDOW_SIZE] = Interval{};
        pre_integrations[WINDOW_SIZE] = nullptr;
        dt_buf[WINDOW_SIZE].clear();
        linear_acceleration_buf[WINDOW_SIZE].clear();
        angular_velocity_buf[WINDOW_SIZE].clear();
        slideWindowNew();
    }
}

void Estimator::slideWindowNew() {
    sum_of_front++;
    if (m_store_live) {
        const int st = lfvio_fm_remove_front(m_store, frame_count);
        if (st != LFVIO_OK) last_status = st;
        return;
    }
    f_manager.removeFront(frame_count);
}

void Estimator::slideWindowOld() {
    sum_of_back++;
    if (solver_flag == NON_LINEAR) {
        const Mat3 R0 = matmul(back_R0, ric[0]), R1 = matmul(Rs[0], ric[0]);
        const Vec3 a = matvec(back_R0, tic[0]), b = matvec(Rs[0], tic[0]);
        const Vec3 P0{back_P0[0] + a[0], back_P0[1] + a[1], back_P0[2] + a[2]}, P1{Ps[0][0] + b[0], Ps[0][1] + b[1], Ps[0][2] + b[2]};
        if (m_store_live) {
            const int st = lfvio_fm_remove_back_shift_depth(m_store, R0.data(), P0.data(), R1.data(), P1.data());
            if (st != LFVIO_OK) last_status = st;
            return;
        }
        f_manager.removeBackShiftDepth(R0, P0, R1, P1);
    } else {
        f_manager.removeBack();
    }
}

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
    if (!m_store_live) {  // (with the device store para_Feature is written on the device, straight into the solver's state)
        const std::vector<double> dep = f_manager.getDepthVector();
        if (para_Feature.size() < dep.size()) para_Feature.resize(dep.size());
        for (size_t i = 0; i < dep.size(); i++) para_Feature[i] = dep[i];
    }
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
    if (m_store_live) {
        if (!m_depth_on_device && m_store_L > 0) {  // (the one-call path has applied setDepth on the device already)
            int n = 0;
            const int st = lfvio_fm_depth(m_store, 0, para_Feature.data(), m_store_L, &n);
            if (st != LFVIO_OK) last_status = st;
        }
        m_depth_on_device = true;
    } else {
        std::vector<double> dep(f_manager.getFeatureCount());
        for (size_t i = 0; i < dep.size(); i++) dep[i] = para_Feature[i];
        f_manager.setDepth(dep);
    }
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
    if (m_store_live) return;  // the projection factors are built on the device (lfvio_fm_build_factors)
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
    if (flushIMU() != LFVIO_OK) return;  // pre_integrations[] and the predicted newest state up to date
    double t_lap = g_ht.on ? HostTimers::now() : 0.0;
    auto lap = [&](int i) {
        if (g_ht.on) {
            const double t = HostTimers::now();
            g_ht.acc[i] += t - t_lap;
            g_ht.calls[i]++;
            t_lap = t;
        }
    };
    vector2double();
    buildFactors();
    lap(3);
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
    const bool one_call = !failure_occur;
    m_depth_on_device = false;
    m_store_L = 0;
    if (m_store_live && !one_call && (last_status = storeDownloadFactors()) != LFVIO_OK) return;  // the list on the host after all
    if (m_store_live && (int)para_Feature.size() < m_max_features) para_Feature.resize(m_max_features);
    const int L = m_store_live ? m_store_L : f_manager.getFeatureCount();
    lfvio_window_state st{WINDOW_SIZE + 1, L, para_Pose.data(), para_SpeedBias.data(), para_Ex_Pose[0], para_Feature.data(), para_Td[0]};
    lfvio_factor_list fl{(int)m_proj.size(), m_proj.data(), (int)m_imu.size(), m_imu.data(), lm.valid() ? &prior : nullptr};
    // A failed solve is not fatal in the reference: ceres::Solve (estimator.cpp:825) leaves the parameter blocks at the last
    // accepted state, optimization() goes on to the marginalisation, processImage() to failureDetection() / slideWindow().
    // Same here for the numeric and capacity failures of the device solve; only CUDA / argument errors stop the estimator.
    const auto pose0 = para_Pose;
    const auto sb0 = para_SpeedBias;
    const auto feat0 = para_Feature;
    double ex0[7];
    memcpy(ex0, para_Ex_Pose[0], sizeof(ex0));
    const double td0 = para_Td[0][0];
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
    const int mflag = marginalization_flag == MARGIN_OLD ? LFVIO_MARGIN_OLD : LFVIO_MARGIN_SECOND_NEW;
    // One call does solve + gauge re-anchoring + next prior on the device (two graph launches, one synchronisation) whenever the
    // re-anchoring is the ordinary one; after a detected failure double2vector() anchors at last_R0 / last_P0 instead
    // (estimator.cpp:1059-1064), which the device does not know: then the solve and the marginalisation are two calls.
    if (one_call && m_store_live) {
        last_status = lfvio_window_optimize_full_fm(m_window, m_store, &st, &fl, &o, mflag, &summary, &out);
        m_store_L = st.n_landmarks;
        m_depth_on_device = last_status == LFVIO_OK;
    } else if (one_call) {
        last_status = lfvio_window_optimize_full(m_window, &st, &fl, &o, mflag, &summary, &out);  // ceres::Solve (estimator.cpp:825) + :833-1005
    } else {
        last_status = lfvio_window_optimize(m_window, &st, &fl, &o, &summary);  // ceres::Solve (estimator.cpp:825)
    }
    lap(4);
    if (last_status == LFVIO_ERR_NUMERIC || last_status == LFVIO_ERR_CAPACITY) {
        solve_failures++;
        last_solve_failure = last_status;
        last_status = LFVIO_OK;
        if (m_store_live && one_call) m_depth_on_device = true;  // the store keeps the depths the failed solve started from
        bool finite = true;
        for (double v : para_Pose) finite = finite && std::isfinite(v);
        for (double v : para_SpeedBias) finite = finite && std::isfinite(v);
        if (!finite || last_solve_failure == LFVIO_ERR_CAPACITY) {  // nothing usable came back: the state the solve started from
            para_Pose = pose0;
            para_SpeedBias = sb0;
            memcpy(para_Ex_Pose[0], ex0, sizeof(ex0));
            para_Feature = feat0;
            para_Td[0][0] = td0;
        }
        if (last_solve_failure == LFVIO_ERR_CAPACITY || one_call) {
            // capacity: the marginalisation would hit the same limit; one call: its prior was not produced.  Go on without a prior.
            double2vector();
            last_marginalization_info = MarginalizationInfo{};
            return;
        }
    }
    if (last_status != LFVIO_OK) return;
    double2vector();  // (after the one-call path the state is already re-anchored: this is then the copy into Ps / Rs / ..)
    if (!one_call) {
        // ---- next prior (estimator.cpp:833-1005): the dropped factors are evaluated at the re-anchored state ----
        vector2double();
        lap(5);
        last_status = lfvio_window_marginalize(m_window, &st, &fl, &o, mflag, &out);
        lap(6);
    }
    if (last_status == LFVIO_ERR_NUMERIC || last_status == LFVIO_ERR_CAPACITY) {  // no usable prior: drop it (its block ids would be stale after the slide)
        solve_failures++;
        last_solve_failure = last_status;
        last_status = LFVIO_OK;
        last_marginalization_info = MarginalizationInfo{};
        return;
    }
    if (last_status != LFVIO_OK) return;
    if (out.n > 0) {  // MARGIN_SECOND_NEW without a matching prior leaves the old one in place (estimator.cpp:942-943)
        next.n = out.n;
        next.n_blocks = out.n_blocks;
        last_marginalization_info = std::move(next);
    } else if (marginalization_flag == MARGIN_OLD) {
        // the reference replaces the prior unconditionally (estimator.cpp:933-936); the window is about to shift, so a kept
        // prior would name the wrong frames
        last_marginalization_info = MarginalizationInfo{};
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

// Test hook: drives Estimator::processIMU / flushIMU and FeatureManager::triangulate of the host mirror from Python.
extern "C" int lfvio_host_process_imu(int frame, const double *P, const double *V, const double *R, const double *ba, const double *bg,
                                      const double *g, const double *noise, const double *acc0, const double *gyr0, const double *dt,
                                      const double *acc, const double *gyr, int n, lfvio_imu_factor *out, double *state_out) {
    using namespace lfvio_host;
    Estimator est(10, 8, 0);
    if (est.last_status != LFVIO_OK) return est.last_status;
    G = Vec3{g[0], g[1], g[2]};
    est.imu_noise = lfvio_imu_noise{noise[0], noise[1], noise[2], noise[3]};
    est.frame_count = frame;
    est.Ps[frame] = Vec3{P[0], P[1], P[2]};
    est.Vs[frame] = Vec3{V[0], V[1], V[2]};
    for (int k = 0; k < 9; k++) est.Rs[frame][k] = R[k];
    est.Bas[frame] = Vec3{ba[0], ba[1], ba[2]};
    est.Bgs[frame] = Vec3{bg[0], bg[1], bg[2]};
    est.first_imu = true;  // the previous frame's last sample
    est.acc_0 = Vec3{acc0[0], acc0[1], acc0[2]};
    est.gyr_0 = Vec3{gyr0[0], gyr0[1], gyr0[2]};
    for (int k = 0; k < n; k++)
        est.processIMU(dt[k], Vec3{acc[3 * k], acc[3 * k + 1], acc[3 * k + 2]}, Vec3{gyr[3 * k], gyr[3 * k + 1], gyr[3 * k + 2]});
    const int st = est.flushIMU();
    if (st != LFVIO_OK) return st;
    if (!est.pre_integrations[frame]) return LFVIO_ERR_ARG;
    *out = *est.pre_integrations[frame];
    for (int k = 0; k < 3; k++) {
        state_out[k] = est.Ps[frame][k];
        state_out[3 + k] = est.Vs[frame][k];
    }
    for (int k = 0; k < 9; k++) state_out[6 + k] = est.Rs[frame][k];
    return LFVIO_OK;
}

extern "C" int lfvio_host_triangulate(const double *Ps, const double *Rs, int n_frames, const double *tic, const double *ric,
                                      const int *start_frame, const int *obs_first, const double *obs_pts, int n, double *depth) {
    using namespace lfvio_host;
    FeatureManager fm;
    fm.window_size = n_frames - 1;
    for (int l = 0; l < n; l++) {
        FeaturePerId fp;
        fp.feature_id = l;
        fp.start_frame = start_frame[l];
        fp.estimated_depth = depth[l];
        for (int k = obs_first[l]; k < obs_first[l + 1]; k++) {
            FeaturePerFrame o;
            o.point = Vec3{obs_pts[3 * k], obs_pts[3 * k + 1], obs_pts[3 * k + 2]};
            fp.feature_per_frame.push_back(o);
        }
        fm.feature.push_back(fp);
    }
    std::vector<Vec3> P(n_frames);
    std::vector<Mat3> R(n_frames);
    for (int f = 0; f < n_frames; f++) {
        P[f] = Vec3{Ps[3 * f], Ps[3 * f + 1], Ps[3 * f + 2]};
        for (int k = 0; k < 9; k++) R[f][k] = Rs[9 * f + k];
    }
    Vec3 t[1] = {Vec3{tic[0], tic[1], tic[2]}};
    Mat3 r[1];
    for (int k = 0; k < 9; k++) r[0][k] = ric[k];
    const int st = fm.triangulate(P, R, t, r, 0);
    if (st != LFVIO_OK) return st;
    int l = 0;
    for (const auto &it : fm.feature) depth[l++] = it.estimated_depth;
    return LFVIO_OK;
}

// ---- test hooks: the C++ Estimator driven frame by frame from Python (closed loop, tests/test_host_classes_gpu.py) ----
extern "C" void *lfvio_host_estimator_create(int window_size, int max_features, int device) {
    auto *e = new lfvio_host::Estimator(window_size, max_features, device);
    if (e->last_status != LFVIO_OK) {
        delete e;
        return nullptr;
    }
    return e;
}
extern "C" void lfvio_host_estimator_destroy(void *h) { delete static_cast<lfvio_host::Estimator *>(h); }

extern "C" int lfvio_host_estimator_set_params(void *h, const double *noise4, const double *g3, const double *tic3, const double *ric9,
                                               double td, double min_parallax, int estimate_td, const double *ba3, const double *bg3) {
    using namespace lfvio_host;
    auto *e = static_cast<Estimator *>(h);
    e->imu_noise = lfvio_imu_noise{noise4[0], noise4[1], noise4[2], noise4[3]};
    G = Vec3{g3[0], g3[1], g3[2]};
    e->TIC0 = Vec3{tic3[0], tic3[1], tic3[2]};
    for (int k = 0; k < 9; k++) e->RIC0[k] = ric9[k];
    TD = td;
    e->setParameter();
    MIN_PARALLAX = min_parallax;
    ESTIMATE_TD = estimate_td;
    for (int i = 0; i <= e->WINDOW_SIZE; i++) {
        e->Bas[i] = Vec3{ba3[0], ba3[1], ba3[2]};
        e->Bgs[i] = Vec3{bg3[0], bg3[1], bg3[2]};
    }
    return LFVIO_OK;
}

// acc0/gyr0 (may be NULL): override acc_0 / gyr_0 before the interval, for streams whose intervals do not share a sample
extern "C" int lfvio_host_estimator_process_imu(void *h, const double *acc0, const double *gyr0, const double *dt, const double *acc,
                                                const double *gyr, int n) {
    using namespace lfvio_host;
    auto *e = static_cast<Estimator *>(h);
    if (acc0 && gyr0) {
        e->first_imu = true;
        e->acc_0 = Vec3{acc0[0], acc0[1], acc0[2]};
        e->gyr_0 = Vec3{gyr0[0], gyr0[1], gyr0[2]};
    }
    for (int k = 0; k < n; k++)
        e->processIMU(dt[k], Vec3{acc[3 * k], acc[3 * k + 1], acc[3 * k + 2]}, Vec3{gyr[3 * k], gyr[3 * k + 1], gyr[3 * k + 2]});
    return LFVIO_OK;
}

// states of all window frames (replaces initialStructure); pending intervals are integrated first so that their
// state prediction cannot overwrite what is set here
extern "C" int lfvio_host_estimator_set_states(void *h, const double *P, const double *R, const double *V, int nonlinear) {
    using namespace lfvio_host;
    auto *e = static_cast<Estimator *>(h);
    const int st = e->flushIMU();
    if (st != LFVIO_OK) return st;
    for (int i = 0; i <= e->WINDOW_SIZE; i++) {
        e->Ps[i] = Vec3{P[3 * i], P[3 * i + 1], P[3 * i + 2]};
        e->Vs[i] = Vec3{V[3 * i], V[3 * i + 1], V[3 * i + 2]};
        for (int k = 0; k < 9; k++) e->Rs[i][k] = R[9 * i + k];
    }
    e->last_P = e->Ps[e->WINDOW_SIZE];
    e->last_R = e->Rs[e->WINDOW_SIZE];
    e->last_P0 = e->Ps[0];
    e->last_R0 = e->Rs[0];
    if (nonlinear) e->solver_flag = Estimator::NON_LINEAR;
    e->g = G;  // injected states live in the configured gravity frame
    return LFVIO_OK;
}

extern "C" int lfvio_host_estimator_use_device_store(void *h, int on) {
    auto *e = static_cast<lfvio_host::Estimator *>(h);
    e->use_device_store = on != 0;
    return e->storeLive() ? 1 : 0;
}

// switch the mirror to the reference's own start-up: processImage() runs initialStructure() once the window is full
extern "C" int lfvio_host_estimator_enable_initialization(void *h, unsigned long long ransac_seed, int stop_after) {
    auto *e = static_cast<lfvio_host::Estimator *>(h);
    e->run_initial_structure = true;
    e->stop_after_initial_structure = stop_after != 0;
    e->ransac_seed = ransac_seed;
    return LFVIO_OK;
}
// info: [0] solver_flag, [1] init_attempts, [2] init_fail_stage, [3] sfm ok, [4] BA iterations, [5] all_image_frame.size(),
// [6] reboot_count;
// g: the aligned gravity
extern "C" int lfvio_host_estimator_init_info(void *h, int *info, double *g) {
    auto *e = static_cast<lfvio_host::Estimator *>(h);
    info[0] = (int)e->solver_flag;
    info[1] = (int)e->init_attempts;
    info[2] = e->init_fail_stage;
    info[3] = e->sfm_summary.ok;
    info[4] = e->sfm_summary.ba_iterations;
    info[5] = (int)e->all_image_frame.size();
    info[6] = e->reboot_count;
    for (int k = 0; k < 3; k++) g[k] = e->g[k];
    return LFVIO_OK;
}

// test hook: move last_P so that the next processImage() trips failureDetection() (|P - last_P| > 5 m, estimator.cpp:645-650)
extern "C" int lfvio_host_estimator_poke_last_p(void *h, double dx) {
    auto *e = static_cast<lfvio_host::Estimator *>(h);
    e->last_P[0] += dx;
    return e->reboot_count;
}

// test hook: make the covariance of pre_integrations[j] singular, so that the next solves fail numerically (LFVIO_ERR_NUMERIC)
// until that interval has left the window; returns how many failed solves the estimator has survived so far
extern "C" int lfvio_host_estimator_poke_covariance(void *h, int j) {
    auto *e = static_cast<lfvio_host::Estimator *>(h);
    e->pokeCovariance(j);
    return e->solve_failures;
}

extern "C" int lfvio_host_estimator_process_image(void *h, const int *ids, const double *feat8, int n, double stamp) {
    using namespace lfvio_host;
    auto *e = static_cast<Estimator *>(h);
    ImageFeatures image;
    for (int i = 0; i < n; i++) {
        std::array<double, 8> p;
        for (int k = 0; k < 8; k++) p[k] = feat8[8 * i + k];
        image[ids[i]].emplace_back(0, p);
    }
    e->processImage(image, stamp);
    return e->last_status;
}

extern "C" int lfvio_host_estimator_get(void *h, double *P, double *R, double *V, double *ba, double *bg, double *td, int *info) {
    using namespace lfvio_host;
    auto *e = static_cast<Estimator *>(h);
    for (int i = 0; i <= e->WINDOW_SIZE; i++)
        for (int k = 0; k < 3; k++) {
            P[3 * i + k] = e->Ps[i][k];
            V[3 * i + k] = e->Vs[i][k];
            ba[3 * i + k] = e->Bas[i][k];
            bg[3 * i + k] = e->Bgs[i][k];
        }
    for (int i = 0; i <= e->WINDOW_SIZE; i++)
        for (int k = 0; k < 9; k++) R[9 * i + k] = e->Rs[i][k];
    *td = e->td;
    info[0] = (int)e->marginalization_flag;
    info[1] = e->storeSize();
    info[2] = e->summary.iterations;
    info[3] = e->frame_count;
    info[4] = e->failure_occur ? 1 : 0;
    info[5] = e->sum_of_back;
    info[6] = e->sum_of_front;
    return LFVIO_OK;
}

// ---- FeatureManager alone (pure host bookkeeping: usable and testable without a CUDA device) ------------------------------
extern "C" void *lfvio_host_fm_create(int window_size, double min_parallax) {
    auto *f = new lfvio_host::FeatureManager();
    f->window_size = window_size;
    lfvio_host::MIN_PARALLAX = min_parallax;
    return f;
}
extern "C" void lfvio_host_fm_destroy(void *h) { delete static_cast<lfvio_host::FeatureManager *>(h); }
// ids (n) / bearings (n x 3): one image; returns addFeatureCheckParallax (1 = key frame)
extern "C" int lfvio_host_fm_add(void *h, int frame_count, const int *ids, const double *bearings, int n, double td) {
    using namespace lfvio_host;
    ImageFeatures image;
    for (int i = 0; i < n; i++)
        image[ids[i]].emplace_back(0, std::array<double, 8>{bearings[3 * i], bearings[3 * i + 1], bearings[3 * i + 2], 0, 0, 0, 0, 0});
    return static_cast<FeatureManager *>(h)->addFeatureCheckParallax(frame_count, image, td) ? 1 : 0;
}
extern "C" int lfvio_host_fm_corresponding(void *h, int l, int r, double *out, int cap) {
    const auto c = static_cast<lfvio_host::FeatureManager *>(h)->getCorresponding(l, r);
    const int n = (int)c.size() < cap ? (int)c.size() : cap;
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) out[6 * i + k] = c[i].first[k], out[6 * i + 3 + k] = c[i].second[k];
    return (int)c.size();
}
// op: 0 removeBack, 1 removeFront(arg), 2 removeBackShiftDepth(pose: marg_R 9 | marg_P 3 | new_R 9 | new_P 3)
extern "C" void lfvio_host_fm_slide(void *h, int op, int arg, const double *pose) {
    using namespace lfvio_host;
    auto *f = static_cast<FeatureManager *>(h);
    if (op == 0) f->removeBack();
    else if (op == 1) f->removeFront(arg);
    else {
        Mat3 mR{}, nR{};
        for (int k = 0; k < 9; k++) mR[k] = pose[k], nR[k] = pose[12 + k];
        f->removeBackShiftDepth(mR, Vec3{pose[9], pose[10], pose[11]}, nR, Vec3{pose[21], pose[22], pose[23]});
    }
}
// snapshot: per feature [id, start_frame, n_obs] (ints) and estimated_depth; set_depth != NULL first assigns estimated_depth in
// list order.  Returns the number of features (getFeatureCount() in *count_out).
extern "C" int lfvio_host_fm_snapshot(void *h, int *meta, double *depth, int cap, const double *set_depth, int *count_out) {
    auto *f = static_cast<lfvio_host::FeatureManager *>(h);
    int i = 0;
    for (auto &it : f->feature) {
        if (set_depth) it.estimated_depth = set_depth[i];
        if (i < cap) {
            meta[3 * i] = it.feature_id, meta[3 * i + 1] = it.start_frame, meta[3 * i + 2] = (int)it.feature_per_frame.size();
            depth[i] = it.estimated_depth;
        }
        i++;
    }
    if (count_out) *count_out = f->getFeatureCount();
    return i;
}

// ---- vector2double / double2vector alone (pure host math: the gauge re-anchoring of estimator.cpp:532-583) -----------------
// P, R (row-major), V, Ba, Bg: the F = window_size + 1 states before the solve; pose7 (x y z qx qy qz qw) / sb9: the optimised
// parameter arrays as ceres leaves them; ex7.  On return P / R / V / Ba / Bg hold what double2vector() writes, pose7_rt / sb9_rt
// what the following vector2double() packs again.  Needs no device (the estimator is built without a device window).
extern "C" int lfvio_host_double2vector(int window_size, double *P, double *R, double *V, double *Ba, double *Bg, const double *pose7,
                                        const double *sb9, const double *ex7, double *pose7_rt, double *sb9_rt) {
    using namespace lfvio_host;
    Estimator e(window_size, 16, -1);  // (window creation fails without a device and is not needed here)
    const int F = window_size + 1;
    for (int i = 0; i < F; i++) {
        e.Ps[i] = Vec3{P[3 * i], P[3 * i + 1], P[3 * i + 2]};
        e.Vs[i] = Vec3{V[3 * i], V[3 * i + 1], V[3 * i + 2]};
        e.Bas[i] = Vec3{Ba[3 * i], Ba[3 * i + 1], Ba[3 * i + 2]};
        e.Bgs[i] = Vec3{Bg[3 * i], Bg[3 * i + 1], Bg[3 * i + 2]};
        for (int k = 0; k < 9; k++) e.Rs[i][k] = R[9 * i + k];
    }
    for (int k = 0; k < 7 * F; k++) e.para_Pose[k] = pose7[k];
    for (int k = 0; k < 9 * F; k++) e.para_SpeedBias[k] = sb9[k];
    for (int k = 0; k < 7; k++) e.para_Ex_Pose[0][k] = ex7[k];
    e.double2vector();
    for (int i = 0; i < F; i++)
        for (int k = 0; k < 3; k++) {
            P[3 * i + k] = e.Ps[i][k], V[3 * i + k] = e.Vs[i][k], Ba[3 * i + k] = e.Bas[i][k], Bg[3 * i + k] = e.Bgs[i][k];
        }
    for (int i = 0; i < F; i++)
        for (int k = 0; k < 9; k++) R[9 * i + k] = e.Rs[i][k];
    e.vector2double();
    for (int k = 0; k < 7 * F; k++) pose7_rt[k] = e.para_Pose[k];
    for (int k = 0; k < 9 * F; k++) sb9_rt[k] = e.para_SpeedBias[k];
    return 0;
}
