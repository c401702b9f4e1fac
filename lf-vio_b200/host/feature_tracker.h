// Host-side mirror of the reference's front-end classes over the lfvio C-ABI (no OpenCV / Eigen / ROS needed).
//
// Mirrors (same names, argument meaning and member names, so feature_tracker_node.cpp reads the same members):
//   class FeatureTracker            feature_tracker/src/feature_tracker.h:34-80
//   camodocal::Camera / OCAMCamera  camera_model/include/camodocal/camera_models/{Camera.h:12-135,ScaramuzzaCamera.h}
//   the extern globals              feature_tracker/src/parameters.h (ROW, COL, MAX_CNT, MIN_DIST, EQUALIZE, ...)
// cv::Mat / cv::Point2f / Eigen::Vector3d are replaced by the minimal PODs below; INTEGRATION.md shows the one-line
// adapters for a build that has OpenCV and Eigen.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "lfvio.h"

// Where the reference's third-party types exist, the mirrors also carry the reference's own signatures (so that
// feature_tracker_node.cpp / estimator_node.cpp call them unchanged): readImage(const cv::Mat &, double), fisheye_mask / cur_img
// as cv::Mat, liftProjective / spaceToPlane on Eigen vectors.  -DLFVIO_HOST_NO_OPENCV / -DLFVIO_HOST_NO_EIGEN switch them off.
#if !defined(LFVIO_HOST_NO_OPENCV) && defined(__has_include)
#if __has_include(<opencv2/core.hpp>)
#include <opencv2/core.hpp>
#define LFVIO_HOST_HAVE_OPENCV 1
#endif
#endif
#if !defined(LFVIO_HOST_NO_EIGEN) && defined(__has_include)
#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
#define LFVIO_HOST_HAVE_EIGEN 1
#elif __has_include(<eigen3/Eigen/Dense>)
#include <eigen3/Eigen/Dense>
#define LFVIO_HOST_HAVE_EIGEN 1
#endif
#endif

namespace lfvio_host {

struct Point2f {
    float x = 0, y = 0;
};
struct Point3f {
    float x = 0, y = 0, z = 0;
};
using Vector2d = std::array<double, 2>;
using Vector3d = std::array<double, 3>;

// Non-owning view of a CV_8UC1 image (what `const cv::Mat &` carries: data, rows, cols, step).
struct ImageView {
    const uint8_t *data = nullptr;
    int rows = 0, cols = 0, step = 0;
};

// ---- feature_tracker/src/parameters.h globals (read implicitly by the reference) ---------------------
extern int ROW, COL, MAX_CNT, MIN_DIST, EQUALIZE, FISHEYE, FREQ;
extern bool PUB_THIS_FRAME;
extern double CENTER_X, CENTER_Y;
extern int MAX_R, MIN_R;
// rejectWithF (feature_tracker.cpp:149, :329-386) runs on every published frame in the reference.  Its sample sets
// come from std::random_device there; here the random source is the counter-based generator seeded by RANSAC_SEED.
extern int REJECT_WITH_F;
extern unsigned long long RANSAC_SEED;
// readParameters(): the YAML keys of config/mindvision/mindvision.yaml (image_height, image_width, max_cnt,
// min_dist, freq, equalize, center_x, center_y, max_r, min_r).  Returns false if the file cannot be read.
bool readParameters(const std::string &config_file);

namespace camodocal {

class Camera {
  public:
    virtual ~Camera() {}
    virtual void liftSphere(const Vector2d &p, Vector3d &P) const = 0;      // Camera.h:72
    virtual void liftProjective(const Vector2d &p, Vector3d &P) const = 0;  // Camera.h:76
    virtual void spaceToPlane(const Vector3d &P, Vector2d &p) const = 0;    // Camera.h:80
#ifdef LFVIO_HOST_HAVE_EIGEN
    // the reference's own signatures (Camera.h:72-80)
    void liftSphere(const Eigen::Vector2d &p, Eigen::Vector3d &P) const {
        Vector3d o;
        liftSphere(Vector2d{p[0], p[1]}, o);
        P = Eigen::Vector3d(o[0], o[1], o[2]);
    }
    void liftProjective(const Eigen::Vector2d &p, Eigen::Vector3d &P) const {
        Vector3d o;
        liftProjective(Vector2d{p[0], p[1]}, o);
        P = Eigen::Vector3d(o[0], o[1], o[2]);
    }
    void spaceToPlane(const Eigen::Vector3d &P, Eigen::Vector2d &p) const {
        Vector2d o;
        spaceToPlane(Vector3d{P[0], P[1], P[2]}, o);
        p = Eigen::Vector2d(o[0], o[1]);
    }
#endif
    virtual int imageWidth() const = 0;
    virtual int imageHeight() const = 0;
};
using CameraPtr = std::shared_ptr<Camera>;

// Scaramuzza / OCamCalib model; liftProjective runs the batched sm_100a kernel through lfvio_lift_projective.
class OCAMCamera : public Camera {
  public:
    explicit OCAMCamera(const lfvio_ocam &params, int device = 0) : mParameters(params), m_device(device) {}
    void liftSphere(const Vector2d &p, Vector3d &P) const override;
    void liftProjective(const Vector2d &p, Vector3d &P) const override;
    void spaceToPlane(const Vector3d &P, Vector2d &p) const override;  // ScaramuzzaCamera.cc:653-674 (host arithmetic, as in the reference)
    using Camera::liftProjective;
    using Camera::liftSphere;
    using Camera::spaceToPlane;
    // batched forms (what the tracker uses): uv n x 2 -> xyz n x 3
    bool liftProjective(const std::vector<Vector2d> &p, std::vector<Vector3d> &P, bool normalize = false) const;
    int imageWidth() const override { return mParameters.image_width; }
    int imageHeight() const override { return mParameters.image_height; }
    const lfvio_ocam &getParameters() const { return mParameters; }

  private:
    lfvio_ocam mParameters;
    int m_device;
};

// CameraFactory::generateCameraFromYamlFile (CameraFactory.cc:89-170), Scaramuzza branch only; null on failure.
CameraPtr generateCameraFromYamlFile(const std::string &filename, int device = 0);

}  // namespace camodocal

class FeatureTracker {
  public:
    FeatureTracker();
    ~FeatureTracker();
    FeatureTracker(const FeatureTracker &) = delete;
    FeatureTracker &operator=(const FeatureTracker &) = delete;

    void readImage(const ImageView &_img, double _cur_time);        // feature_tracker.cpp:95-184
#ifdef LFVIO_HOST_HAVE_OPENCV
    // the reference's own signature (feature_tracker.h:39): a CV_8UC1 image, continuous or not
    void readImage(const cv::Mat &_img, double _cur_time) { readImage(ImageView{_img.data, _img.rows, _img.cols, (int)_img.step}, _cur_time); }
    // members the node reads (feature_tracker.h:56-57).  The annulus mask is rasterised on the device from CENTER_X/Y, MIN_R,
    // MAX_R (what feature_tracker_node.cpp builds into fisheye_mask); both members are refreshed from the device on demand.
    cv::Mat fisheye_mask, cur_img;
    bool fetchMaskMat() {
        if (!fetchMask()) return false;
        fisheye_mask = cv::Mat(ROW, COL, CV_8UC1, mask.data());
        return true;
    }
#endif
    bool updateID(unsigned int i);                                   // :388-398
    void readIntrinsicParameter(const std::string &calib_file);     // :400-404
    void setCamera(const camodocal::CameraPtr &cam, const lfvio_ocam &params);

    // public members the node reads (feature_tracker_node.cpp:131-160); filled after every readImage
    std::vector<Point2f> cur_pts;
    std::vector<Point3f> cur_un_pts_3d, pts_velocity_3d;
    std::vector<int> ids, track_cnt;
    std::vector<uint8_t> mask;  // ROW x COL, refreshed on demand by fetchMask()
    camodocal::CameraPtr m_camera;
    double cur_time = 0, prev_time = 0;
    static int n_id;

    bool fetchMask();
    const char *lastError() const;
    int device = 0;

  private:
    bool ensureHandle();
    void pull();
    void flushIds();
    lfvio_tracker *m_handle = nullptr;
    lfvio_ocam m_params{};
    bool m_have_params = false, m_ids_dirty = false;
    int m_ids_base = 0;  // value of n_id when the current frame's updateID loop started
};

}  // namespace lfvio_host
