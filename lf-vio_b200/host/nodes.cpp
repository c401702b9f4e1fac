#include "nodes.h"

#include <cmath>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <sstream>

namespace lfvio_host {

using namespace detail;

// ------------------------------------------------------------------------------------------------------------------
// feature_tracker_node.cpp
// ------------------------------------------------------------------------------------------------------------------
bool FeatureTrackerNode::img_callback(const ImageView &img, double stamp, PointCloudMsg *out, bool *restart) {
    if (restart) *restart = false;
    if (first_image_flag) {  // :29-35
        first_image_flag = false;
        first_image_time = stamp;
        last_image_time = stamp;
        return false;
    }
    if (stamp - last_image_time > 1.0 || stamp < last_image_time) {  // :36-47 unstable camera stream
        first_image_flag = true;
        last_image_time = 0;
        pub_count = 1;
        if (restart) *restart = true;
        return false;
    }
    last_image_time = stamp;
    if (std::round(1.0 * pub_count / (stamp - first_image_time)) <= FREQ) {  // :49-62 frequency control
        PUB_THIS_FRAME = true;
        if (std::fabs(1.0 * pub_count / (stamp - first_image_time) - FREQ) < 0.01 * FREQ) {
            first_image_time = stamp;
            pub_count = 0;
        }
    } else {
        PUB_THIS_FRAME = false;
    }
    trackerData[0].readImage(img, stamp);  // :82-88
    for (unsigned int i = 0;; i++)          // :100-108
        if (!trackerData[0].updateID(i)) break;
    if (!PUB_THIS_FRAME) return false;
    pub_count++;
    PointCloudMsg msg;
    msg.stamp = stamp;
    const FeatureTracker &t = trackerData[0];
    for (size_t j = 0; j < t.ids.size(); j++) {
        if (t.track_cnt[j] <= 1) continue;  // :142
        msg.points.push_back(t.cur_un_pts_3d[j]);
        msg.channels[0].push_back((float)(t.ids[j] * 1 + 0));  // p_id * NUM_OF_CAM + i, as float32 (:154)
        msg.channels[1].push_back(t.cur_pts[j].x);
        msg.channels[2].push_back(t.cur_pts[j].y);
        msg.channels[3].push_back(t.pts_velocity_3d[j].x);
        msg.channels[4].push_back(t.pts_velocity_3d[j].y);
        msg.channels[5].push_back(t.pts_velocity_3d[j].z);
    }
    if (!init_pub) {  // :173-178: the very first publish is suppressed (no optical-flow speed yet)
        init_pub = true;
        return false;
    }
    if (out) *out = std::move(msg);
    return true;
}

// ------------------------------------------------------------------------------------------------------------------
// estimator_node.cpp
// ------------------------------------------------------------------------------------------------------------------
namespace {
std::array<double, 4> qmul(const std::array<double, 4> &a, const std::array<double, 4> &b) {  // w x y z
    return {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3], a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]};
}
Vec3 qrotv(const std::array<double, 4> &q, const Vec3 &v) { return matvec(quat_to_rot(q[0], q[1], q[2], q[3], false), v); }
}  // namespace

void EstimatorNode::predict(const ImuMsg &m) {  // :40-77
    const double t = m.stamp;
    if (init_imu) {
        latest_time = t;
        init_imu = false;
        return;
    }
    const double dt = t - latest_time;
    latest_time = t;
    const Vec3 &g = estimator.g;
    auto sub = [](const Vec3 &a, const Vec3 &b) { return Vec3{a[0] - b[0], a[1] - b[1], a[2] - b[2]}; };
    const Vec3 un_acc_0 = sub(qrotv(tmp_Q, sub(acc_0, tmp_Ba)), g);
    Vec3 un_gyr;
    for (int k = 0; k < 3; k++) un_gyr[k] = 0.5 * (gyr_0[k] + m.angular_velocity[k]) - tmp_Bg[k];
    tmp_Q = qmul(tmp_Q, {1.0, un_gyr[0] * dt / 2, un_gyr[1] * dt / 2, un_gyr[2] * dt / 2});  // Utility::deltaQ (no normalisation)
    const Vec3 un_acc_1 = sub(qrotv(tmp_Q, sub(m.linear_acceleration, tmp_Ba)), g);
    for (int k = 0; k < 3; k++) {
        const double un_acc = 0.5 * (un_acc_0[k] + un_acc_1[k]);
        tmp_P[k] = tmp_P[k] + dt * tmp_V[k] + 0.5 * dt * dt * un_acc;
        tmp_V[k] = tmp_V[k] + dt * un_acc;
    }
    acc_0 = m.linear_acceleration;
    gyr_0 = m.angular_velocity;
}

void EstimatorNode::update() {  // :79-98
    const int W = estimator.WINDOW_SIZE;
    latest_time = current_time;
    tmp_P = estimator.Ps[W];
    const auto q = rot_to_quat(estimator.Rs[W]);  // x y z w
    tmp_Q = {q[3], q[0], q[1], q[2]};
    tmp_V = estimator.Vs[W];
    tmp_Ba = estimator.Bas[W];
    tmp_Bg = estimator.Bgs[W];
    acc_0 = estimator.acc_0;
    gyr_0 = estimator.gyr_0;
    for (const ImuMsg &m : imu_buf) predict(m);
}

void EstimatorNode::imu_callback(const ImuMsg &imu_msg) {
    if (imu_msg.stamp <= last_imu_t) return;  // "imu message in disorder!"
    last_imu_t = imu_msg.stamp;
    imu_buf.push_back(imu_msg);
    predict(imu_msg);
}

void EstimatorNode::feature_callback(const PointCloudMsg &msg) {
    if (!init_feature) {  // skip the first detected feature, which doesn't contain optical flow speed
        init_feature = true;
        return;
    }
    feature_buf.push_back(msg);
}

void EstimatorNode::restart_callback(bool restart) {
    if (!restart) return;
    feature_buf.clear();
    imu_buf.clear();
    estimator.clearState();
    estimator.setParameter();
    current_time = -1;
    last_imu_t = 0;
}

int EstimatorNode::process() {
    int processed = 0;
    while (true) {  // getMeasurements (:100-134)
        if (imu_buf.empty() || feature_buf.empty()) break;
        if (!(imu_buf.back().stamp > feature_buf.front().stamp + estimator.td)) {
            sum_of_wait++;
            break;
        }
        if (!(imu_buf.front().stamp < feature_buf.front().stamp + estimator.td)) {
            feature_buf.pop_front();  // "throw img, only should happen at the beginning"
            continue;
        }
        const PointCloudMsg img_msg = std::move(feature_buf.front());
        feature_buf.pop_front();
        std::vector<ImuMsg> IMUs;
        while (imu_buf.front().stamp < img_msg.stamp + estimator.td) {
            IMUs.push_back(imu_buf.front());
            imu_buf.pop_front();
        }
        IMUs.push_back(imu_buf.front());
        // process() (:219-262)
        double dx = 0, dy = 0, dz = 0, rx = 0, ry = 0, rz = 0;
        for (const ImuMsg &m : IMUs) {
            const double t = m.stamp, img_t = img_msg.stamp + estimator.td;
            if (t <= img_t) {
                if (current_time < 0) current_time = t;
                const double dt = t - current_time;
                current_time = t;
                dx = m.linear_acceleration[0], dy = m.linear_acceleration[1], dz = m.linear_acceleration[2];
                rx = m.angular_velocity[0], ry = m.angular_velocity[1], rz = m.angular_velocity[2];
                estimator.processIMU(dt, Vec3{dx, dy, dz}, Vec3{rx, ry, rz});
            } else {  // interpolate the sample at the image time
                const double dt_1 = img_t - current_time, dt_2 = t - img_t;
                current_time = img_t;
                const double w1 = dt_2 / (dt_1 + dt_2), w2 = dt_1 / (dt_1 + dt_2);
                dx = w1 * dx + w2 * m.linear_acceleration[0];
                dy = w1 * dy + w2 * m.linear_acceleration[1];
                dz = w1 * dz + w2 * m.linear_acceleration[2];
                rx = w1 * rx + w2 * m.angular_velocity[0];
                ry = w1 * ry + w2 * m.angular_velocity[1];
                rz = w1 * rz + w2 * m.angular_velocity[2];
                estimator.processIMU(dt_1, Vec3{dx, dy, dz}, Vec3{rx, ry, rz});
            }
        }
        ImageFeatures image;  // :295-318
        for (size_t i = 0; i < img_msg.points.size(); i++) {
            const int v = (int)(img_msg.channels[0][i] + 0.5);
            const int feature_id = v / 1, camera_id = v % 1;  // NUM_OF_CAM = 1
            image[feature_id].emplace_back(camera_id, std::array<double, 8>{img_msg.points[i].x, img_msg.points[i].y, img_msg.points[i].z,
                                                                            img_msg.channels[1][i], img_msg.channels[2][i], img_msg.channels[3][i],
                                                                            img_msg.channels[4][i], img_msg.channels[5][i]});
        }
        estimator.processImage(image, img_msg.stamp);
        pubOdometry(img_msg.stamp);
        processed++;
        if (estimator.solver_flag == Estimator::NON_LINEAR) update();  // :335-341
    }
    return processed;
}

std::string tum_line(double stamp, const Vec3 &p, const std::array<double, 4> &q) {
    std::ostringstream o;
    o.setf(std::ios::fixed, std::ios::floatfield);
    o.precision(12);
    o << stamp << " " << p[0] << " " << p[1] << " " << p[2] << " " << q[0] << " " << q[1] << " " << q[2] << " " << q[3] << std::endl;
    return o.str();
}

void EstimatorNode::pubOdometry(double stamp) {
    if (estimator.solver_flag != Estimator::NON_LINEAR) return;  // visualization.cpp:114
    const int W = estimator.WINDOW_SIZE;
    OdometryMsg m;
    m.stamp = stamp;
    m.position = estimator.Ps[W];
    m.velocity = estimator.Vs[W];
    m.orientation = rot_to_quat(estimator.Rs[W]);
    odometry.push_back(m);
    if (!VINS_RESULT_PATH.empty()) {
        std::ofstream foutC(VINS_RESULT_PATH, std::ios::app);
        foutC << tum_line(stamp, m.position, m.orientation);
    }
}

}  // namespace lfvio_host

// ---- C entry points for the Python tests ---------------------------------------------------------------------------
// device < 0: build the node without a device window (its creation fails and is tolerated) — for CPU-only tests of the
// host-side message synchronisation; such a node must not be fed more than WINDOW_SIZE images (the first device call
// happens when the window is full).
extern "C" void *lfvio_host_estimator_node_create(int window_size, int max_features, int device, const char *result_path) {
    auto *n = new lfvio_host::EstimatorNode(window_size, max_features, device);
    if (n->estimator.last_status != LFVIO_OK && device >= 0) {
        delete n;
        return nullptr;
    }
    if (result_path) n->VINS_RESULT_PATH = result_path;
    return n;
}
extern "C" void lfvio_host_estimator_node_destroy(void *h) { delete static_cast<lfvio_host::EstimatorNode *>(h); }
extern "C" void *lfvio_host_estimator_node_estimator(void *h) { return &static_cast<lfvio_host::EstimatorNode *>(h)->estimator; }
extern "C" void lfvio_host_estimator_node_imu(void *h, double stamp, const double *acc, const double *gyr) {
    lfvio_host::ImuMsg m;
    m.stamp = stamp;
    for (int k = 0; k < 3; k++) m.linear_acceleration[k] = acc[k], m.angular_velocity[k] = gyr[k];
    static_cast<lfvio_host::EstimatorNode *>(h)->imu_callback(m);
}
// feat: n x 9 float32 = [x y z | id | u v | vx vy vz], i.e. the point and the six channels of one PointCloud entry
extern "C" void lfvio_host_estimator_node_feature(void *h, double stamp, const float *feat, int n) {
    lfvio_host::PointCloudMsg m;
    m.stamp = stamp;
    for (int i = 0; i < n; i++) {
        const float *f = feat + 9 * i;
        lfvio_host::Point3f p;
        p.x = f[0], p.y = f[1], p.z = f[2];
        m.points.push_back(p);
        for (int c = 0; c < 6; c++) m.channels[c].push_back(f[3 + c]);
    }
    static_cast<lfvio_host::EstimatorNode *>(h)->feature_callback(m);
}
extern "C" int lfvio_host_estimator_node_process(void *h) { return static_cast<lfvio_host::EstimatorNode *>(h)->process(); }
// latest IMU-rate prediction: P (3), Q (4, w x y z), V (3); returns the number of odometry messages published so far
extern "C" int lfvio_host_estimator_node_latest(void *h, double *P, double *Q, double *V) {
    auto *n = static_cast<lfvio_host::EstimatorNode *>(h);
    for (int k = 0; k < 3; k++) P[k] = n->tmp_P[k], V[k] = n->tmp_V[k];
    for (int k = 0; k < 4; k++) Q[k] = n->tmp_Q[k];
    return (int)n->odometry.size();
}
extern "C" int lfvio_host_tum_line(double stamp, const double *p, const double *q_xyzw, char *out, int cap) {
    const std::string s = lfvio_host::tum_line(stamp, {p[0], p[1], p[2]}, {q_xyzw[0], q_xyzw[1], q_xyzw[2], q_xyzw[3]});
    std::snprintf(out, cap, "%s", s.c_str());
    return (int)s.size();
}

// feature-tracker node: one image in, at most one feature message out.  feat_out: cap x 9 float32 (layout as above).
// Returns the number of points published, -1 when nothing was published, -2 on a stream restart.
extern "C" void *lfvio_host_tracker_node_create() { return new lfvio_host::FeatureTrackerNode(); }
extern "C" void lfvio_host_tracker_node_destroy(void *h) { delete static_cast<lfvio_host::FeatureTrackerNode *>(h); }
extern "C" void *lfvio_host_tracker_node_tracker(void *h) { return &static_cast<lfvio_host::FeatureTrackerNode *>(h)->trackerData[0]; }
extern "C" int lfvio_host_tracker_node_image(void *h, const uint8_t *img, int rows, int cols, int step, double stamp, float *feat_out, int cap) {
    lfvio_host::ImageView v;
    v.data = img, v.rows = rows, v.cols = cols, v.step = step;
    lfvio_host::PointCloudMsg m;
    bool restart = false;
    const bool pub = static_cast<lfvio_host::FeatureTrackerNode *>(h)->img_callback(v, stamp, &m, &restart);
    if (restart) return -2;
    if (!pub) return -1;
    const int n = (int)m.points.size() < cap ? (int)m.points.size() : cap;
    for (int i = 0; i < n; i++) {
        float *f = feat_out + 9 * i;
        f[0] = m.points[i].x, f[1] = m.points[i].y, f[2] = m.points[i].z;
        for (int c = 0; c < 6; c++) f[3 + c] = m.channels[c][i];
    }
    return n;
}
// readParameters() + readIntrinsicParameter() of the node's main() (feature_tracker_node.cpp:222-240) from values
extern "C" int lfvio_host_tracker_node_configure(void *h, const lfvio_ocam *cam, int rows, int cols, int max_cnt, int min_dist, int freq,
                                                 int reject_with_f, unsigned long long ransac_seed) {
    using namespace lfvio_host;
    ROW = rows, COL = cols, MAX_CNT = max_cnt, MIN_DIST = min_dist, FREQ = freq, EQUALIZE = 1;
    REJECT_WITH_F = reject_with_f;
    RANSAC_SEED = ransac_seed;
    FeatureTracker::n_id = 0;
    auto *n = static_cast<FeatureTrackerNode *>(h);
    n->trackerData[0].setCamera(std::make_shared<camodocal::OCAMCamera>(*cam), *cam);
    return LFVIO_OK;
}

// the IMU samples buffered for window slot j (dt_buf / linear_acceleration_buf / angular_velocity_buf, estimator.h:96-98);
// returns their number.  info (may be NULL): [frame_count, sum_of_wait, imu_buf.size(), feature_buf.size()]
extern "C" int lfvio_host_estimator_node_imu_buf(void *h, int j, double *dt, double *acc, double *gyr, int cap, int *info) {
    auto *n = static_cast<lfvio_host::EstimatorNode *>(h);
    const auto &e = n->estimator;
    if (info) {
        info[0] = e.frame_count, info[1] = n->sum_of_wait, info[2] = (int)n->imu_buf.size(), info[3] = (int)n->feature_buf.size();
    }
    if (j < 0 || j > e.WINDOW_SIZE) return -1;
    const int m = (int)e.dt_buf[j].size();
    for (int i = 0; i < m && i < cap; i++) {
        dt[i] = e.dt_buf[j][i];
        for (int k = 0; k < 3; k++) acc[3 * i + k] = e.linear_acceleration_buf[j][i][k], gyr[3 * i + k] = e.angular_velocity_buf[j][i][k];
    }
    return m;
}
extern "C" void lfvio_host_estimator_node_set_td(void *h, double td) { static_cast<lfvio_host::EstimatorNode *>(h)->estimator.td = td; }
extern "C" void lfvio_host_estimator_node_restart(void *h) { static_cast<lfvio_host::EstimatorNode *>(h)->restart_callback(true); }
// FREQ alone (parameters.cpp), for CPU-only tests of the node's frequency control: without a camera the tracker has no
// device handle, readImage returns at once and the node publishes empty messages on the frames it would publish on
extern "C" void lfvio_host_tracker_node_set_freq(int freq) { lfvio_host::FREQ = freq; }
