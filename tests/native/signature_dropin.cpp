// Compiles only where the host mirrors carry the REFERENCE's own signatures (tests/test_signature_dropin.py):
//   FeatureTracker::readImage(const cv::Mat &, double)                         feature_tracker/src/feature_tracker.h:39
//   FeatureTracker::fisheye_mask / cur_img                                       feature_tracker.h:56-57
//   Camera::liftSphere / liftProjective / spaceToPlane on Eigen vectors          camodocal/camera_models/Camera.h:72-80
//   Estimator::processImage(const map<int, vector<pair<int, Matrix<double,8,1>>>> &, const std_msgs::Header &)   estimator.h:35
#include "estimator.h"
#include "feature_tracker.h"

#if !defined(LFVIO_HOST_HAVE_OPENCV) || !defined(LFVIO_HOST_HAVE_EIGEN) || !defined(LFVIO_HOST_HAVE_ROS_HEADER)
#error "the guarded overloads were not enabled by the stand-in headers"
#endif

using namespace lfvio_host;

extern "C" int lfvio_signature_dropin(FeatureTracker *trackerData, Estimator *estimator, unsigned char *pixels, int rows, int cols, double stamp,
                                      const double *xyz, double *uv) {
    cv::Mat img(rows, cols, CV_8UC1, pixels);
    trackerData->readImage(img, stamp);  // as feature_tracker_node.cpp:103 calls it
    trackerData->fisheye_mask = cv::Mat(rows, cols, CV_8UC1, pixels);
    const bool have_cur = !trackerData->cur_img.empty();
    Eigen::Vector3d P(xyz[0], xyz[1], xyz[2]);
    Eigen::Vector2d p;
    trackerData->m_camera->spaceToPlane(P, p);  // Camera.h:80
    uv[0] = p[0];
    uv[1] = p[1];
    Eigen::Vector3d back;
    trackerData->m_camera->liftProjective(p, back);  // Camera.h:76
    std::map<int, std::vector<std::pair<int, Eigen::Matrix<double, 8, 1>>>> image;
    Eigen::Matrix<double, 8, 1> xyz_uv_velocity;
    xyz_uv_velocity << back[0], back[1], back[2], p[0], p[1], 0.0, 0.0, 0.0;
    image[7].emplace_back(0, xyz_uv_velocity);
    std_msgs::Header header;
    header.stamp.t = stamp;
    estimator->processImage(image, header);  // as estimator_node.cpp:311 calls it
    return have_cur ? 1 : 0;
}
