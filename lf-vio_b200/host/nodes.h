// ROS-free shims of the reference's two nodes: the callbacks and the measurement
// synchronisation of feature_tracker/src/feature_tracker_node.cpp and vins_estimator/src/estimator_node.cpp on plain
// structs with the layout of the ROS messages on the wire (SURVEY.md §8b "Wire format"), plus the TUM trajectory writer
// of utility/visualization.cpp:173-179.  A ROS build wraps these 1:1: the subscriber callbacks forward
// the message fields, the publishers send what the shims return.  Single-threaded: the reference's mutexes / condition
// variable guard exactly the calls made here in sequence, so the caller drains process() after feeding callbacks.
#pragma once
#include <deque>
#include <string>
#include <vector>

#include "estimator.h"
#include "feature_tracker.h"

namespace lfvio_host {

// sensor_msgs/PointCloud as published on /feature_tracker/feature (feature_tracker_node.cpp:113-172):
// points[i] = unit bearing (float32 x y z); channels[0] = id * NUM_OF_CAM + cam as float32, [1] [2] = pixel u v,
// [3..5] = bearing velocity.
struct PointCloudMsg {
    double stamp = 0;
    std::vector<Point3f> points;
    std::vector<float> channels[6];
};
struct ImuMsg {  // sensor_msgs/Imu fields the estimator reads
    double stamp = 0;
    Vec3 linear_acceleration{}, angular_velocity{};
};
struct OdometryMsg {  // nav_msgs/Odometry fields pubOdometry fills (visualization.cpp:118-135)
    double stamp = 0;
    Vec3 position{}, velocity{};
    std::array<double, 4> orientation{};  // x y z w
};

class FeatureTrackerNode {
  public:
    FeatureTracker trackerData[1];  // NUM_OF_CAM = 1 (config/mindvision/mindvision.yaml)
    // img_callback (feature_tracker_node.cpp:27-215).  Returns true when a feature message was published into *out;
    // *restart (may be NULL) is set when the stream is discontinuous (:36-47, the /feature_tracker/restart topic).
    bool img_callback(const ImageView &img, double stamp, PointCloudMsg *out, bool *restart);

    double first_image_time = 0, last_image_time = 0;
    int pub_count = 1;
    bool first_image_flag = true, init_pub = false;
};

class EstimatorNode {
  public:
    explicit EstimatorNode(int window_size = 10, int max_features = 1000, int device = 0) : estimator(window_size, max_features, device) {}
    Estimator estimator;
    void imu_callback(const ImuMsg &imu_msg);            // estimator_node.cpp:136-162
    void feature_callback(const PointCloudMsg &msg);     // :164-176
    void restart_callback(bool restart);                 // :178-198
    // process() (:210-343) for every complete measurement currently buffered; returns the number of images processed.
    int process();
    // pubOdometry (visualization.cpp:112-181): the odometry message of the newest frame and, when VINS_RESULT_PATH is set,
    // one TUM line "stamp px py pz qx qy qz qw" (fixed, 12 decimals) appended to the file.
    std::string VINS_RESULT_PATH;
    std::vector<OdometryMsg> odometry;    // what /vins_estimator/odometry carried
    // latest IMU-rate prediction (predict() / update(), :40-98), published as imu_propagate in the reference
    Vec3 tmp_P{}, tmp_V{}, tmp_Ba{}, tmp_Bg{}, acc_0{}, gyr_0{};
    std::array<double, 4> tmp_Q{1, 0, 0, 0};  // w x y z
    double latest_time = 0, current_time = -1, last_imu_t = 0;
    bool init_feature = false, init_imu = true;
    int sum_of_wait = 0;
    std::deque<ImuMsg> imu_buf;
    std::deque<PointCloudMsg> feature_buf;

  private:
    void predict(const ImuMsg &imu_msg);
    void update();
    void pubOdometry(double stamp);
};

// one line of the TUM trajectory file exactly as visualization.cpp:173-179 writes it
std::string tum_line(double stamp, const Vec3 &p, const std::array<double, 4> &q_xyzw);

}  // namespace lfvio_host
