// Host-side mirror of the back-end part of class Estimator over the lfvio C-ABI (no Eigen / Ceres / ROS needed).
//
// Mirrors vins_estimator/src/estimator.h:25-137 for the hot path only:
//   optimization()            estimator.cpp:676-1009   (problem build -> lfvio_window_optimize -> double2vector ->
//                                                       lfvio_window_marginalize -> last_marginalization_*)
//   vector2double()/double2vector()  estimator.cpp:488-530 / :532-626 (incl. the yaw / position gauge re-anchoring)
//   FeatureManager::getFeatureCount/getDepthVector/setDepth   feature_manager.cpp:36-43,145-197
// processIMU / processImage / slideWindow / failureDetection and the initialisation (initialStructure, host/initial.cpp)
// are the SURVEY.md §8(f) "next" rows, mirrored as well.  WINDOW_SIZE and NUM_OF_F are runtime values here
// (compile-time constants in the reference, vins_estimator/src/parameters.h:12-13).
#pragma once
#include <array>
#include <list>
#include <map>
#include <utility>
#include <vector>

#include "lfvio.h"

#if !defined(LFVIO_HOST_NO_EIGEN) && defined(__has_include)
#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
#define LFVIO_HOST_HAVE_EIGEN 1
#elif __has_include(<eigen3/Eigen/Dense>)
#include <eigen3/Eigen/Dense>
#define LFVIO_HOST_HAVE_EIGEN 1
#endif
#if __has_include(<std_msgs/Header.h>)
#include <std_msgs/Header.h>
#define LFVIO_HOST_HAVE_ROS_HEADER 1
#endif
#endif

namespace lfvio_host {

using Vec3 = std::array<double, 3>;
using Mat3 = std::array<double, 9>;  // row-major

// ---- vins_estimator/src/parameters.h globals ---------------------------------------------------------
extern int ESTIMATE_TD, ESTIMATE_EXTRINSIC, NUM_ITERATIONS;
extern double TR, TD, FOCAL_LENGTH, MIN_PARALLAX;  // MIN_PARALLAX = keyframe_parallax / FOCAL_LENGTH (parameters.cpp:56-57)
extern int EST_ROW;  // ROW of the estimator's parameters.cpp
extern Vec3 G;

struct FeaturePerFrame {  // feature_manager.h:17-41
    Vec3 point{}, velocity{};
    double cur_td = 0;
    std::array<double, 2> uv{};
};
struct FeaturePerId {  // feature_manager.h:43-71
    int feature_id = 0, start_frame = 0, used_num = 0, solve_flag = 0;
    double estimated_depth = -1.0;
    std::vector<FeaturePerFrame> feature_per_frame;
};
// what estimator_node hands to processImage: feature id -> [(camera id, [x y z u v vx vy vz])]  (estimator_node.cpp:290-310)
using ImageFeatures = std::map<int, std::vector<std::pair<int, std::array<double, 8>>>>;

class FeatureManager {
  public:
    std::list<FeaturePerId> feature;
    int window_size = 10;
    int last_track_num = 0;
    void clearState() { feature.clear(); }
    // feature_manager.cpp:46-105 (bearing parallax: acos(p_i . p_j) * 10, :353-369)
    bool addFeatureCheckParallax(int frame_count, const ImageFeatures &image, double td);
    void removeBackShiftDepth(const Mat3 &marg_R, const Vec3 &marg_P, const Mat3 &new_R, const Vec3 &new_P);  // :271-310
    void removeBack();                  // :312-325
    void removeFront(int frame_count);  // :327-345
    void removeFailures();              // :158-167 (a no-op in LF-VIO: setDepth never sets solve_flag = 2)
    int getFeatureCount();
    std::vector<double> getDepthVector();
    void setDepth(const std::vector<double> &x);
    // feature_manager.cpp:199-253 — every untriangulated eligible track in ONE lfvio_triangulate call.
    // Returns the C-ABI status (the reference returns void).
    int triangulate(const std::vector<Vec3> &Ps, const std::vector<Mat3> &Rs, const Vec3 tic[], const Mat3 ric[], int device = 0);
    // feature_manager.cpp:118-143: bearings of the features seen in both frames
    std::vector<std::pair<Vec3, Vec3>> getCorresponding(int frame_count_l, int frame_count_r);
    void clearDepth();  // clearDepth(dep = -1) as visualInitialAlign uses it (estimator.cpp:400-403, feature_manager.cpp:169-180)
    double init_depth = 5.0;  // INIT_DEPTH (parameters.cpp:116)
};

// IntegrationBase members the IMU factor reads (factor/integration_base.h:188-206) == lfvio_imu_factor
using IntegrationBase = lfvio_imu_factor;

struct MarginalizationInfo {  // what the reference keeps between frames (marginalization_factor.h:43-73)
    int n = 0, n_blocks = 0;
    std::vector<int> block_id, block_idx;
    std::vector<double> x0, J0, r0;
    bool valid() const { return n > 0; }
};

// initial/initial_alignment.h:13-30.  pre_integration is kept as its IMU samples + linearisation point (so that
// IntegrationBase::repropagate is one batched lfvio_preintegrate call over all frames) and the resulting record.
struct ImageFrame {
    ImageFeatures points;
    double t = 0;
    Mat3 R{};
    Vec3 T{};
    bool is_key_frame = false;
    bool has_pre = false;  // pre_integration != nullptr
    Vec3 acc_start{}, gyr_start{}, ba{}, bg{};
    std::vector<double> dt;
    std::vector<Vec3> acc, gyr;
    IntegrationBase pre{};
};

// initial/initial_ex_rotation.h:16-45 — online calibration of the camera-IMU rotation (ESTIMATE_EXTRINSIC == 2)
class InitialEXRotation {
  public:
    InitialEXRotation() { ric = Mat3{1, 0, 0, 0, 1, 0, 0, 0, 1}; }
    // initial_ex_rotation.cpp:13-68.  delta_q_imu (w, x, y, z).  The two numeric steps are lfvio_solve_relative_r and
    // lfvio_calibrate_ex_rotation; status: last C-ABI status.
    bool CalibrationExRotation(const std::vector<std::pair<Vec3, Vec3>> &corres, const std::array<double, 4> &delta_q_imu,
                               Mat3 &calib_ric_result);
    int window_size = 10, device = 0, status = 0;
    unsigned long long ransac_seed = 1;
    double sv[4] = {0, 0, 0, 0};  // singular values of the last solve (ric_cov = sv[1..3])

  private:
    int frame_count = 0;
    std::vector<double> Rc, Rimu, Rc_g;  // 9 doubles per pair (the identity pushed by the reference's constructor is never read)
    Mat3 ric;
};

class Estimator {
  public:
    enum MarginalizationFlag { MARGIN_OLD = 0, MARGIN_SECOND_NEW = 1 };

    explicit Estimator(int window_size = 10, int max_features = 1000, int device = 0);
    ~Estimator();
    Estimator(const Estimator &) = delete;
    Estimator &operator=(const Estimator &) = delete;

    // estimator.cpp:86-120.  Samples are buffered per frame exactly like dt_buf / linear_acceleration_buf /
    // angular_velocity_buf; the pre-integration record and the predicted Ps/Vs/Rs[frame_count] are brought up to date by
    // flushIMU() (one lfvio_preintegrate call over the frame's whole interval), which optimization() calls itself.
    void processIMU(double dt, const Vec3 &linear_acceleration, const Vec3 &angular_velocity);
    int flushIMU();
    // IntegrationBase::repropagate (integration_base.h:39-53) for frame j with new linearisation biases
    int repropagate(int j, const Vec3 &ba, const Vec3 &bg);
    // estimator.cpp:122-219 without the initialisation (initialStructure): while solver_flag ==
    // INITIAL the window only fills and slides; the caller switches to NON_LINEAR once the states are initialised.
    enum SolverFlag { INITIAL = 0, NON_LINEAR = 1 };
    void processImage(const ImageFeatures &image, double header_stamp);
#if defined(LFVIO_HOST_HAVE_EIGEN) && defined(LFVIO_HOST_HAVE_ROS_HEADER)
    // the reference's own signature (estimator.h:35)
    void processImage(const std::map<int, std::vector<std::pair<int, Eigen::Matrix<double, 8, 1>>>> &image, const std_msgs::Header &header) {
        ImageFeatures f;
        for (const auto &kv : image)
            for (const auto &obs : kv.second) {
                std::array<double, 8> a;
                for (int k = 0; k < 8; k++) a[k] = obs.second(k);
                f[kv.first].emplace_back(obs.first, a);
            }
        processImage(f, header.stamp.toSec());
    }
#endif
    // estimator.cpp:10-21 / :23-84.  TIC / RIC / TD are globals of parameters.cpp in the reference; the mirror keeps the
    // configured extrinsics per instance (TIC0 / RIC0) and TD as the global it is there.
    void setParameter();
    void clearState();
    Vec3 TIC0{0, 0, 0};
    Mat3 RIC0{1, 0, 0, 0, 1, 0, 0, 0, 1};
    int reboot_count = 0;  // how often failureDetection() rebooted the estimator (estimator.cpp:196-204)
    void pokeCovariance(int j);  // test hook: make pre_integrations[j]'s covariance singular
    int solve_failures = 0;      // solves / marginalisations that failed numerically or on capacity and were survived
    int last_solve_failure = 0;  // status of the last such failure
    // estimator.cpp:221-376 / :378-455 / :457-484: the numeric steps run on the device through lfvio_solve_relative_rt,
    // lfvio_global_sfm, lfvio_epnp (all non-keyframes in one call), lfvio_solve_gyro_bias, lfvio_preintegrate (the
    // repropagations, batched) and lfvio_linear_alignment.
    bool initialStructure();
    bool visualInitialAlign();
    bool relativePose(Mat3 &relative_R, Vec3 &relative_T, int &l);
    InitialEXRotation initial_ex_rotation;
    // FeatureManager on the device (lfvio_fm_*).  When set, the track list moves to the device store the
    // first time processImage() runs with solver_flag == NON_LINEAR (from the host-side list the initialisation worked on) and
    // stays there until clearState(): addFeatureCheckParallax, triangulate, the factor list, setDepth and the slides then run
    // where the tracks live, and optimization() never builds or uploads a projection-factor list.  f_manager is stale meanwhile.
    bool use_device_store = false;
    bool storeLive() const { return m_store_live; }
    int storeSize();  // features in the device store (synchronises it)
    bool run_initial_structure = false;  // false: the pre-initialisation behaviour of this mirror (the caller sets the states)
    bool stop_after_initial_structure = false;  // test hook: return from processImage right after a successful initialStructure
    std::map<double, ImageFrame> all_image_frame;
    double initial_timestamp = 0;
    Vec3 g{};
    unsigned long long ransac_seed = 1;  // injected random source of the 8-point RANSAC (see lfvio_reject_with_f)
    unsigned init_attempts = 0;
    lfvio_sfm_summary sfm_summary{};
    int init_fail_stage = 0;  // 0 ok, 1 relativePose, 2 GlobalSFM, 3 PnP of a non-keyframe, 4 alignment (debug)
    void solveOdometry();      // :475-486
    bool failureDetection();   // :628-675
    void slideWindow();        // :1011-1105
    void slideWindowNew();     // :1107-1112
    void slideWindowOld();     // :1114-1131
    void optimization();    // estimator.cpp:676-1009
    void vector2double();   // :488-530
    void double2vector();   // :532-626

    const int WINDOW_SIZE;
    // public state, names as in the reference (estimator.h:70-113)
    std::vector<Vec3> Ps, Vs, Bas, Bgs;
    std::vector<Mat3> Rs;
    Vec3 tic[1];
    Mat3 ric[1];
    double td = 0;
    std::vector<IntegrationBase *> pre_integrations;  // [0 .. WINDOW_SIZE], entry j connects frames j-1 and j
    int frame_count = 0;
    SolverFlag solver_flag = INITIAL;
    std::vector<double> Headers;       // header.stamp.toSec() per window slot
    std::vector<Vec3> key_poses;
    Mat3 back_R0{}, last_R{};
    Vec3 back_P0{}, last_P{};
    int sum_of_back = 0, sum_of_front = 0;
    bool first_imu = false;
    Vec3 acc_0{}, gyr_0{};
    std::vector<std::vector<double>> dt_buf;                                   // estimator.h:96-98
    std::vector<std::vector<Vec3>> linear_acceleration_buf, angular_velocity_buf;
    lfvio_imu_noise imu_noise{0.08, 0.004, 0.00004, 2.0e-6};                  // ACC_N GYR_N ACC_W GYR_W
    FeatureManager f_manager;
    MarginalizationFlag marginalization_flag = MARGIN_OLD;
    MarginalizationInfo last_marginalization_info;  // replaces MarginalizationInfo* + last_marginalization_parameter_blocks
    bool failure_occur = false;
    Mat3 last_R0{};
    Vec3 last_P0{};
    // flat parameter arrays (estimator.h:107-113)
    std::vector<double> para_Pose, para_SpeedBias, para_Feature;
    double para_Ex_Pose[1][7]{};
    double para_Td[1][1]{};
    lfvio_solve_summary summary{};
    int last_status = 0;
    const char *lastError() const;

  private:
    lfvio_window *m_window = nullptr;
    lfvio_fm *m_store = nullptr;
    bool m_store_live = false, m_depth_on_device = false;
    int m_store_L = 0, m_max_features = 0;
    int storeLoad();
    int storeTriangulate();
    int storeDownloadFactors();
    std::vector<int> m_ids;      // staging of one image for lfvio_fm_add_frame
    std::vector<double> m_feat8;
    int m_device;
    std::vector<lfvio_proj_factor> m_proj;
    std::vector<lfvio_imu_factor> m_imu;
    void buildFactors();
    // per-frame interval bookkeeping of processIMU / flushIMU
    struct Interval {
        bool open = false, dirty = false, predict = false;  // predict: Ps/Vs/Rs[j] still to be propagated by flushIMU
        Vec3 acc_start{}, gyr_start{}, ba{}, bg{}, P{}, V{};
        Mat3 R{};
    };
    std::vector<Interval> m_iv;
    std::vector<lfvio_imu_factor> m_pre;  // storage behind pre_integrations[] filled by flushIMU
    int integrate(int j, bool propagate);
    ImageFrame m_tmp;          // tmp_pre_integration (estimator.cpp:101-104, 138-140)
    bool m_tmp_valid = false;
    int integrateImageFrames(bool all, const Vec3 *new_bg);
};

namespace detail {  // 3 x 3 helpers shared by estimator.cpp and initial.cpp
Mat3 matmul(const Mat3 &a, const Mat3 &b);
Vec3 matvec(const Mat3 &a, const Vec3 &v);
Mat3 transpose(const Mat3 &a);
std::array<double, 4> rot_to_quat(const Mat3 &m);  // (x, y, z, w)
Mat3 quat_to_rot(double w, double x, double y, double z, bool normalize);
}  // namespace detail

// Utility::R2ypr / ypr2R (utility/utility.h:66-104), degrees
Vec3 R2ypr(const Mat3 &R);
Mat3 ypr2R(const Vec3 &ypr);

}  // namespace lfvio_host
