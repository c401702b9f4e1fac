/*
 * lfvio.h — C-ABI of the B200-native LF-VIO hot paths.
 *
 * This is the drop-in boundary.  The reference (flysoaryun/LF-VIO,
 * CPU-only C++ on OpenCV + Ceres + Eigen) has no FFI: its hot paths sit behind the C++
 * classes FeatureTracker / camodocal::Camera / Estimator and the ceres::CostFunction
 * virtual interface.  The host-side mirrors of those classes (lf-vio_b200/host/) are thin
 * C++ over the entry points below; every entry point cites the reference interface it
 * replaces (paths relative to the reference tree).
 *
 * Conventions: plain C types; caller-owned HOST buffers unless a name ends in _dev;
 * int status (0 = OK, <0 = error, text via lfvio_last_error()); no exceptions cross the
 * boundary; one CUDA stream per handle; a handle is not thread-safe (the reference uses one
 * thread per node); configuration travels in the cfg structs (replacing the reference's
 * extern globals ROW, COL, MAX_CNT, ... — feature_tracker/src/parameters.h).  There is NO
 * CPU fallback inside the library: without a CUDA device every compute entry point fails.
 */
#ifndef LFVIO_H_
#define LFVIO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LFVIO_OK 0
#define LFVIO_ERR_CUDA (-1)     /* CUDA runtime / driver error (incl. no device)          */
#define LFVIO_ERR_ARG (-2)      /* invalid argument                                      */
#define LFVIO_ERR_CAPACITY (-3) /* a fixed-capacity device buffer would overflow         */
#define LFVIO_ERR_NUMERIC (-4)  /* factorisation failed / non-finite state               */

/* Thread-local message of the last failing call ("" if none). */
const char *lfvio_last_error(void);
/* Number of CUDA devices visible (0 when there is no GPU/driver; never fails). */
int lfvio_device_count(void);
/* Library version string. */
const char *lfvio_version(void);
/* Total number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t lfvio_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Scaramuzza / OCamCalib camera (camodocal::OCAMCamera::Parameters,
 * camera_model/include/camodocal/camera_models/ScaramuzzaCamera.h:13-65).
 * ------------------------------------------------------------------------------------------ */
typedef struct lfvio_ocam {
    double poly[5];      /* poly_parameters p0..p4   (ScaramuzzaCamera.cc:58-101 YAML keys) */
    double inv_poly[20]; /* inv_poly_parameters p0..p19 (spaceToPlane only)                 */
    double C, D, E;      /* affine_parameters ac, ad, ae                                     */
    double cx, cy;       /* affine_parameters cx, cy                                         */
    int image_width, image_height;
} lfvio_ocam;

/*
 * Batched OCAMCamera::liftProjective (camera_model/src/camera_models/ScaramuzzaCamera.cc:622-645;
 * virtual Camera::liftProjective, Camera.h:76).  uv: n x 2 f64 pixel coordinates (AoS),
 * xyz: n x 3 f64.  normalize = 0 gives the un-normalised ray (x_a, y_a, -poly(phi)) exactly as
 * liftProjective; normalize = 1 divides by the norm (liftSphere, ScaramuzzaCamera.cc:609-614,
 * and the callers feature_tracker/src/feature_tracker.cpp:342,351,458).  Host buffers.
 */
int lfvio_lift_projective(const lfvio_ocam *cam, const double *uv, double *xyz, int n,
                          int normalize, int device);
/* Same, on device-resident buffers, asynchronous on `stream` (a cudaStream_t, may be 0). */
int lfvio_lift_projective_dev(const lfvio_ocam *cam, const double *uv_dev, double *xyz_dev, int n,
                              int normalize, void *stream);

/* ------------------------------------------------------------------------------------------
 * Stage-level front-end operators: one entry point per third-party call that
 * FeatureTracker::readImage makes (feature_tracker/src/feature_tracker.cpp:95-184).  Host
 * buffers, synchronous, device scratch allocated per call — these exist so a maintainer can
 * swap a cv:: call 1:1 and so each kernel can be parity-tested in isolation; the fused
 * per-frame pipeline is lfvio_tracker_read_image below.  Images are CV_8UC1, `stride` in bytes.
 * ------------------------------------------------------------------------------------------ */
/* cv::createCLAHE(3.0, Size(8,8))->apply(img, out)              feature_tracker.cpp:101-107 */
int lfvio_clahe(const uint8_t *img, int rows, int cols, int stride, double clip_limit,
                uint8_t *out, int out_stride, int device);
/* cv::pyrDown(img, out): out is ((rows+1)/2) x ((cols+1)/2)     (inside calcOpticalFlowPyrLK) */
int lfvio_pyr_down(const uint8_t *img, int rows, int cols, int stride, uint8_t *out, int out_stride,
                   int device);
/* cv::calcOpticalFlowPyrLK(prev, next, prev_pts, next_pts, status, err, Size(41,41), max_level)
 * feature_tracker.cpp:127.  prev_pts/next_pts: n x 2 f32, status: n u8.                         */
int lfvio_calc_optical_flow_pyr_lk(const uint8_t *prev, const uint8_t *next, int rows, int cols,
                                   int stride, const float *prev_pts, int n, float *next_pts,
                                   uint8_t *status, int max_level, int device);
/* cv::cornerMinEigenVal(img, out, 3, 3): out rows x cols f32 (dense)                            */
int lfvio_corner_min_eigen_val(const uint8_t *img, int rows, int cols, int stride, float *out,
                               int device);
/* FeatureTracker::rejectWithF (feature_tracker.cpp:329-386) on n >= 0 point pairs: lifts cur_pts / forw_pts
 * (n x 2 f32 pixels) to the unit sphere, runs myfindFundamentalMat (:266-327: 100 spherical 8-point hypotheses,
 * float score, refit on the inliers of the best) and writes status (n u8, 1 = keep; all ones when n < 8).
 * Sample sets: word k of hypothesis h is the top half of splitmix64(seed + frame*C1 + h*C2 + k*C3), consumed by
 * create_random_array's procedure (random_array.cc:22-49).  Optional outputs: E_out (9 f64, row-major, the refit
 * E_21), scores_out (100 f32, score of every hypothesis), best_out (index of the winning hypothesis).          */
int lfvio_reject_with_f(const lfvio_ocam *cam, const float *cur_pts, const float *forw_pts, int n,
                        unsigned long long seed, unsigned frame, unsigned char *status, double *E_out,
                        float *scores_out, int *best_out);
/* cv::goodFeaturesToTrack(img, corners, max_corners, quality, min_distance, mask)
 * feature_tracker.cpp:166.  mask may be NULL.  corners: max_corners x 2 f32; *n_out <= max_corners. */
int lfvio_good_features_to_track(const uint8_t *img, int rows, int cols, int stride,
                                 const uint8_t *mask, int mask_stride, int max_corners,
                                 double quality, double min_distance, float *corners, int *n_out,
                                 int device);

/* ------------------------------------------------------------------------------------------
 * FeatureTracker (feature_tracker/src/feature_tracker.h:34-80): whole-frame pipeline resident
 * on one GPU.  The cfg struct replaces the reference's extern globals
 * (feature_tracker/src/parameters.h: ROW COL MAX_CNT MIN_DIST EQUALIZE CENTER_X CENTER_Y
 * MAX_R MIN_R) and FeatureTracker::readIntrinsicParameter().
 * ------------------------------------------------------------------------------------------ */
typedef struct lfvio_tracker_cfg {
    int rows, cols;            /* ROW, COL                                                   */
    int max_cnt;               /* MAX_CNT  (YAML max_cnt)                                    */
    int min_dist;              /* MIN_DIST (YAML min_dist)                                   */
    int equalize;              /* EQUALIZE (YAML equalize): CLAHE on/off                     */
    double center_x, center_y; /* CENTER_X/Y (YAML center_x/center_y), truncated like cv::Point */
    int max_r, min_r;          /* MAX_R, MIN_R: annular mask radii                           */
    int capacity;              /* device track capacity; 0 => max_cnt                        */
    int ransac;                /* 1: run rejectWithF() on published frames (feature_tracker.cpp:149) */
    lfvio_ocam cam;            /* m_camera                                                   */
    unsigned long long ransac_seed; /* injected random source of rejectWithF's sample sets: the reference seeds
                                * a new mt19937 from std::random_device per call (random_array.cc:12-20)   */
} lfvio_tracker_cfg;

/* Caller-allocated SoA view of the public members the ROS node reads after readImage
 * (feature_tracker_node.cpp:131-160): cur_pts, ids, track_cnt, cur_un_pts_3d, pts_velocity_3d. */
typedef struct lfvio_tracks {
    int capacity;            /* in: entries each array can hold                              */
    int n;                   /* out                                                          */
    float *cur_pts;          /* n x 2                                                        */
    int *ids;                /* n                                                            */
    int *track_cnt;          /* n                                                            */
    float *cur_un_pts_3d;    /* n x 3 unit bearings                                          */
    float *pts_velocity_3d;  /* n x 3                                                        */
} lfvio_tracks;

typedef struct lfvio_tracker lfvio_tracker;

int lfvio_tracker_create(const lfvio_tracker_cfg *cfg, int device, lfvio_tracker **out);
void lfvio_tracker_destroy(lfvio_tracker *t);
/* FeatureTracker::readImage(_img, _cur_time) with the global PUB_THIS_FRAME passed explicitly
 * (feature_tracker.cpp:95-184).  Host image; H2D copy, all kernels, D2H of the track arrays and
 * one stream synchronise happen inside the call.                                              */
int lfvio_tracker_read_image(lfvio_tracker *t, const uint8_t *img, int stride, double cur_time,
                             int pub_this_frame);
/* Same pipeline on an image already resident in device memory; asynchronous (no host sync, no
 * download), runs the updateID pass on device too.  Pair with lfvio_tracker_sync().           */
int lfvio_tracker_read_image_dev(lfvio_tracker *t, const uint8_t *img_dev, int stride,
                                 double cur_time, int pub_this_frame);
/* Wait for the tracker's stream and refresh the host mirror of the track arrays.              */
int lfvio_tracker_sync(lfvio_tracker *t);
int lfvio_tracker_count(const lfvio_tracker *t);
int lfvio_tracker_get(const lfvio_tracker *t, lfvio_tracks *out);
/* The updateID loop of the node (feature_tracker.cpp:388-398; feature_tracker_node.cpp:103-111):
 * every id == -1 gets (*n_id)++ in index order.  n_id mirrors `static int FeatureTracker::n_id`. */
int lfvio_tracker_update_ids(lfvio_tracker *t, int *n_id);
/* FeatureTracker::mask after the last published frame (rows x cols u8).                       */
int lfvio_tracker_get_mask(lfvio_tracker *t, uint8_t *out, int out_stride);
/* cur_img (after CLAHE) pyramid level `level` (0 = full size).                                */
int lfvio_tracker_get_image(lfvio_tracker *t, int level, uint8_t *out, int out_stride);
/* Overwrite the tracked state (cur_pts, ids, track_cnt) — test hook for order-dependent paths. */
int lfvio_tracker_set_points(lfvio_tracker *t, const float *cur_pts, const int *ids,
                             const int *track_cnt, int n);
/* The cudaStream_t all of this handle's work is enqueued on (for event timing).               */
void *lfvio_tracker_stream(lfvio_tracker *t);

/* ==========================================================================================
 * Sliding-window back end: Estimator::optimization() (vins_estimator/src/estimator.cpp:676-1009).
 *
 * The Ceres problem the reference builds (:678-772) is passed as flat arrays.  Parameter blocks keep the
 * reference's layout (estimator.h:107-113): para_Pose[F][7] = x y z qx qy qz qw, para_SpeedBias[F][9] =
 * v ba bg, para_Ex_Pose[7], para_Feature[L] (inverse depth), para_Td.  Parameter-block ids used by the
 * prior: pose f -> 2f, speed-bias f -> 2f+1, extrinsic -> 2F, td -> 2F+1 (F = n_frames).
 * ========================================================================================== */
typedef struct lfvio_proj_factor { /* ProjectionTdFactor / ProjectionFactor ctor arguments        */
    int frame_i, frame_j;          /* imu_i (host frame), imu_j            estimator.cpp:735-744  */
    int landmark;                  /* feature_index into para_Feature       estimator.cpp:733      */
    int reserved;
    double pts_i[3], pts_j[3];     /* unit bearings (feature_per_frame.point)                      */
    double vel_i[3], vel_j[3];     /* bearing velocities (.velocity); unused when estimate_td = 0  */
    double td_i, td_j;             /* .cur_td                                                      */
    double row_i, row_j;           /* .uv.y()  (rolling-shutter term)                              */
} lfvio_proj_factor;

typedef struct lfvio_imu_factor { /* IMUFactor(pre_integrations[j]) == IntegrationBase members      */
    int frame_i;                  /* connects frame_i and frame_i + 1       estimator.cpp:717-724 */
    int reserved;
    double sum_dt;
    double delta_p[3];
    double delta_q[4];            /* w x y z                                                       */
    double delta_v[3];
    double linearized_ba[3], linearized_bg[3];
    double jacobian[225];         /* 15 x 15 row-major, order O_P O_R O_V O_BA O_BG (parameters.h:50-57) */
    double covariance[225];       /* 15 x 15 row-major                                             */
} lfvio_imu_factor;

typedef struct lfvio_prior { /* MarginalizationInfo after marginalize() (marginalization_factor.h:43-73) */
    int n;              /* residuals == total LOCAL size of the kept blocks (pose blocks count 6)  */
    int n_blocks;       /* keep_block_size.size()                                                  */
    int *block_id;      /* n_blocks parameter-block ids (see above)                                */
    int *block_idx;     /* n_blocks: keep_block_idx - m  (column offset in J0)                     */
    double *x0;         /* keep_block_data concatenated in block order, GLOBAL sizes (7/9/7/1)     */
    double *J0;         /* linearized_jacobians, n x n row-major                                   */
    double *r0;         /* linearized_residuals, n                                                 */
    int cap_n, cap_blocks; /* capacities of the arrays when used as an output                     */
} lfvio_prior;

typedef struct lfvio_window_state {
    int n_frames;            /* F = WINDOW_SIZE + 1                                                */
    int n_landmarks;         /* L                                                                  */
    double *para_Pose;       /* F x 7                                                              */
    double *para_SpeedBias;  /* F x 9                                                              */
    double *para_Ex_Pose;    /* 7                                                                  */
    double *para_Feature;    /* L                                                                  */
    double *para_Td;         /* 1                                                                  */
} lfvio_window_state;

typedef struct lfvio_factor_list {
    int n_proj;
    const lfvio_proj_factor *proj; /* in the reference's insertion order (estimator.cpp:725-772)    */
    int n_imu;
    const lfvio_imu_factor *imu;
    const lfvio_prior *prior;      /* last_marginalization_info, may be NULL                        */
} lfvio_factor_list;

typedef struct lfvio_solver_opts {
    int max_num_iterations;   /* NUM_ITERATIONS (YAML max_num_iterations, 8)                        */
    int fixed_iterations;     /* 1: ignore the function/parameter/gradient tolerances (timing runs) */
    int estimate_td;          /* ESTIMATE_TD: ProjectionTdFactor vs ProjectionFactor                */
    int estimate_extrinsic;   /* 0 => para_Ex_Pose is SetParameterBlockConstant                     */
    int row;                  /* ROW (image height) for the rolling-shutter term                    */
    int reserved;
    double focal_length;      /* FOCAL_LENGTH; sqrt_info = focal_length / 1.5 * I2 (estimator.cpp:18) */
    double tr;                /* TR, rolling shutter read-out time                                  */
    double g[3];              /* gravity vector G                                                   */
    double cauchy_a;          /* ceres::CauchyLoss(1.0)                                             */
    /* ceres::Solver::Options defaults of release 1.12.0 that the reference relies on               */
    double initial_trust_region_radius, max_trust_region_radius, min_trust_region_radius;
    double min_relative_decrease, function_tolerance, gradient_tolerance, parameter_tolerance;
} lfvio_solver_opts;

#define LFVIO_MAX_ITER_LOG 64
typedef struct lfvio_solve_summary {
    int iterations;            /* trust-region iterations performed (successful + unsuccessful)     */
    int num_successful_steps;
    int termination;           /* 0 max iterations, 1 converged, -1 numerical failure               */
    int reserved;
    double initial_cost, final_cost;
    double cost[LFVIO_MAX_ITER_LOG];    /* cost after each iteration (index 0 = initial)            */
    double radius[LFVIO_MAX_ITER_LOG];  /* trust-region radius after each iteration                 */
    int accepted[LFVIO_MAX_ITER_LOG];
} lfvio_solve_summary;

/* Fills *o with the Ceres 1.12 defaults + the reference's settings (estimator.cpp:810-822, parameters.h). */
void lfvio_solver_default_opts(lfvio_solver_opts *o);

#define LFVIO_MARGIN_OLD 0        /* estimator.h MarginalizationFlag */
#define LFVIO_MARGIN_SECOND_NEW 1

typedef struct lfvio_window lfvio_window;
/* window_size = WINDOW_SIZE (10 in the reference, compile-time there; runtime here), max_landmarks
 * replaces NUM_OF_F, max_proj bounds the projection-factor count.                                   */
int lfvio_window_create(int window_size, int max_landmarks, int max_proj, int device, lfvio_window **out);
void lfvio_window_destroy(lfvio_window *w);
/* ceres::Solve() of estimator.cpp:825 on the problem of :678-772 (DENSE_SCHUR + traditional DOGLEG +
 * CauchyLoss on projection factors), entirely on device; state is updated in place.                 */
int lfvio_window_optimize(lfvio_window *w, lfvio_window_state *state, const lfvio_factor_list *factors,
                          const lfvio_solver_opts *opts, lfvio_solve_summary *summary);
/* Construction of the next prior (estimator.cpp:833-1005 + MarginalizationInfo::preMarginalize /
 * marginalize / getParameterBlocks, marginalization_factor.cpp:110-319) at the given state.  The factors
 * to drop are selected from `factors` as the reference does; block ids in `out` are already shifted
 * (addr_shift).  out->n == 0 if nothing is to be marginalised (SECOND_NEW without a matching prior).  */
int lfvio_window_marginalize(lfvio_window *w, const lfvio_window_state *state, const lfvio_factor_list *factors,
                             const lfvio_solver_opts *opts, int margin_flag, lfvio_prior *out);
/* The whole Estimator::optimization() (estimator.cpp:676-1009) in one call, one upload and one synchronisation:
 * solve, then double2vector()/vector2double() (estimator.cpp:488-626: the window is rotated about gravity and shifted so
 * that frame 0 keeps the yaw and position it had on entry), then the next prior at the re-anchored state.  `state` returns
 * the re-anchored parameters; prior_out as in lfvio_window_marginalize.                                              */
int lfvio_window_optimize_full(lfvio_window *w, lfvio_window_state *state, const lfvio_factor_list *factors,
                               const lfvio_solver_opts *opts, int margin_flag, lfvio_solve_summary *summary,
                               lfvio_prior *prior_out);
/* ceres::CostFunction::Evaluate of every factor at `state` (debug / parity hook): residuals are
 * written factor by factor (2 per projection factor incl. the Cauchy correction when robust != 0,
 * 15 per IMU factor, prior.n for the prior); jacobians likewise in LOCAL parameterisation, row-major:
 * projection 2 x 20 [pose_i 6 | pose_j 6 | ex 6 | lambda 1 | td 1], IMU 15 x 30 [pose_i 6 | sb_i 9 |
 * pose_j 6 | sb_j 9], prior n x n.  Any output pointer may be NULL.                                   */
int lfvio_window_evaluate(lfvio_window *w, const lfvio_window_state *state, const lfvio_factor_list *factors,
                          const lfvio_solver_opts *opts, int robust, double *res_proj, double *jac_proj,
                          double *res_imu, double *jac_imu, double *res_prior, double *cost);
void *lfvio_window_stream(lfvio_window *w);
/* Device-resident variant for throughput runs: upload once, then solve repeatedly from the uploaded
 * initial state without host traffic (each call restores the initial state on device first).        */
int lfvio_window_upload(lfvio_window *w, const lfvio_window_state *state, const lfvio_factor_list *factors,
                        const lfvio_solver_opts *opts);
int lfvio_window_optimize_dev(lfvio_window *w, int with_marginalize, int margin_flag);
int lfvio_window_download(lfvio_window *w, lfvio_window_state *state, lfvio_solve_summary *summary);

/* ==========================================================================================
 * FeatureManager on the device: the track list of the window as a structure of arrays in HBM
 * (FeaturePerId / FeaturePerFrame, feature_manager.h:17-64), its bookkeeping as kernels, and the projection-factor list
 * of Estimator::optimization() (estimator.cpp:725-772) built where it is consumed.  List order == the reference's
 * std::list order (insertion order; ids ascend).  Every call is ordered on the handle's own stream; calls that return
 * counts synchronise it.  A capacity error (LFVIO_ERR_CAPACITY) is sticky until lfvio_fm_clear().
 * ========================================================================================== */
typedef struct lfvio_fm lfvio_fm;
/* FeatureManager(Rs) + setRic: window_size = WINDOW_SIZE, capacity = live features the store can hold,
 * max_proj_factors = size of the handle's own factor buffer, min_parallax = MIN_PARALLAX, init_depth = INIT_DEPTH.     */
int lfvio_fm_create(int window_size, int capacity, int max_proj_factors, double min_parallax, double init_depth,
                    int device, lfvio_fm **out);
void lfvio_fm_destroy(lfvio_fm *f);
/* FeatureManager::clearState (feature_manager.cpp:31-34)                                                              */
int lfvio_fm_clear(lfvio_fm *f);
/* Replace the content of the store with a list held on the host (the state of the reference's std::list<FeaturePerId> at the
 * moment the estimator leaves its initialisation): n features in list order, ids ascending; feature i has the
 * observations obs9[obs_first[i] .. obs_first[i+1]) of 9 doubles each, [point 3 | velocity 3 | uv 2 | cur_td]
 * (FeaturePerFrame, feature_manager.h:17-38), in consecutive frames from start_frame[i].  solve_flag may be NULL (0).   */
int lfvio_fm_load(lfvio_fm *f, int n, const int *ids, const int *start_frame, const int *obs_first, const double *depth,
                  const int *solve_flag, const double *obs9);
/* FeatureManager::addFeatureCheckParallax (feature_manager.cpp:45-95): ids ascending (the reference iterates a std::map),
 * feat8 = n x [x y z u v vx vy vz] as in the feature message (estimator.cpp:118-124, camera 0).  Returns the key-frame
 * decision (1 => MARGIN_OLD) and last_track_num.                                                                       */
int lfvio_fm_add_frame(lfvio_fm *f, int frame_count, const int *ids, const double *feat8, int n, double td,
                       int *is_keyframe, int *last_track_num);
/* removeBack (feature_manager.cpp:312-327), removeBackShiftDepth (:271-310; matrices 3 x 3 row-major),
 * removeFront (:329-351), removeFailures (:134-143)                                                                    */
int lfvio_fm_remove_back(lfvio_fm *f);
int lfvio_fm_remove_back_shift_depth(lfvio_fm *f, const double *marg_R, const double *marg_P, const double *new_R,
                                     const double *new_P);
int lfvio_fm_remove_front(lfvio_fm *f, int frame_count);
int lfvio_fm_remove_failures(lfvio_fm *f);
/* mode 0 setDepth (:145-163), 1 clearDepth (:165-174): x in (one value per selected feature, list order);
 * mode 2 getDepthVector (:176-190): x out.  *n_out = getFeatureCount() (:36-43).                                       */
int lfvio_fm_depth(lfvio_fm *f, int mode, double *x, int n_cap, int *n_out);
/* FeatureManager::triangulate (feature_manager.cpp:199-253) over the store: Ps (F x 3) / Rs (F x 9 row-major) IMU poses,
 * tic (3) / ric (9) extrinsics from the host; estimated_depth updated in place on the device.                         */
int lfvio_fm_triangulate(lfvio_fm *f, const double *Ps, const double *Rs, const double *tic, const double *ric);
/* The residual-block loop of estimator.cpp:725-772 + para_Feature (vector2double, :1044-1046) into the handle's own
 * device buffers; lfvio_fm_download_factors copies the last build out (lm_first has n_landmarks + 1 entries).           */
int lfvio_fm_build_factors(lfvio_fm *f, int *n_proj, int *n_landmarks);
int lfvio_fm_download_factors(lfvio_fm *f, lfvio_proj_factor *proj, int proj_cap, int *lm_first, double *para_feature,
                              int lm_cap);
int lfvio_fm_factors_dev(lfvio_fm *f, const lfvio_proj_factor **proj, const int **lm_first, const double **para_feature,
                         void **stream);
/* The list as (feature_id, start_frame, n_observations) + estimated_depth in list order; returns its length.           */
int lfvio_fm_snapshot(lfvio_fm *f, int *meta3, double *depth, int cap);
/* lfvio_window_optimize_full with the projection factors, landmark ranges and para_Feature taken from the device
 * manager (factors->proj is ignored; state->n_landmarks and state->para_Feature are outputs, the latter with room for the
 * window's max_landmarks) and FeatureManager::setDepth of the result (double2vector, estimator.cpp:560-563) applied on
 * the device.  The factor list never exists on the host.                                                               */
int lfvio_window_optimize_full_fm(lfvio_window *w, lfvio_fm *fm, lfvio_window_state *state,
                                  const lfvio_factor_list *factors, const lfvio_solver_opts *opts, int margin_flag,
                                  lfvio_solve_summary *summary, lfvio_prior *prior_out);

/* ==========================================================================================
 * The two steps right before optimization(): they produce the constants the
 * factors above consume.
 * ========================================================================================== */
typedef struct lfvio_imu_noise { /* ACC_N GYR_N ACC_W GYR_W (vins_estimator/src/parameters.cpp:58-61) */
    double acc_n, gyr_n, acc_w, gyr_w;
} lfvio_imu_noise;

typedef struct lfvio_imu_interval { /* one IntegrationBase: the IMU samples between two camera frames */
    int first, count;              /* samples [first, first + count) of the dt / acc / gyr arrays          */
    int frame_i;                   /* copied to lfvio_imu_factor.frame_i                                   */
    int propagate;                 /* != 0: also run the nominal-state prediction of Estimator::processIMU */
    double acc_0[3], gyr_0[3];     /* IntegrationBase(acc_0, gyr_0, ...)   integration_base.h:13-16        */
    double ba[3], bg[3];           /* ... linearized_ba, linearized_bg (Bas[j], Bgs[j], estimator.cpp:97)  */
    double P[3], V[3], R[9];       /* Ps[j], Vs[j], Rs[j] (row-major) before the first sample              */
} lfvio_imu_interval;

/* IntegrationBase::push_back for every sample of every interval (factor/integration_base.h:31-157: mid-point
 * integration of delta_p/q/v, jacobian <- F jacobian, covariance <- F covariance F^T + V noise V^T), i.e. what
 * Estimator::processIMU (estimator.cpp:86-120) does at 200 Hz and IntegrationBase::repropagate (:39-53) does after a
 * bias change.  out[i] is the IMUFactor input for interval i.  state_out (may be NULL): n_intervals x 15 =
 * [P 3 | V 3 | R 9 row-major], the predicted Ps[j], Vs[j], Rs[j] (estimator.cpp:107-116) when propagate != 0. */
int lfvio_preintegrate(const lfvio_imu_interval *intervals, int n_intervals, const double *dt, const double *acc,
                       const double *gyr, int n_samples, const lfvio_imu_noise *noise, const double *g /* 3 */,
                       lfvio_imu_factor *out, double *state_out, int device);

/* FeatureManager::triangulate (vins_estimator/src/feature_manager.cpp:199-253).  Feature l was first seen in frame
 * start_frame[l] and has the unit bearings obs_pts[obs_first[l] .. obs_first[l+1]) (x y z each) in consecutive frames.
 * Ps (n_frames x 3) / Rs (n_frames x 9 row-major) are the IMU poses, tic / ric the camera extrinsics.  depth[l] is
 * in/out like FeaturePerId::estimated_depth: entries > 0 are left untouched (:207), features with fewer than two
 * observations or start_frame >= WINDOW_SIZE - 2 are skipped (:204), a negative result becomes init_depth (:246-249). */
int lfvio_triangulate(const double *Ps, const double *Rs, int n_frames, const double *tic, const double *ric,
                      const int *start_frame, const int *obs_first, const double *obs_pts, int n_features,
                      double init_depth, double *depth, int device);

/* ==========================================================================================
 * Estimator initialisation: the numeric steps of Estimator::initialStructure
 * (vins_estimator/src/estimator.cpp:221-376) and visualInitialAlign (:378-455).  Host buffers, synchronous.
 * Rotation matrices are row-major, quaternions are (w, x, y, z) as Eigen prints them.
 * ========================================================================================== */
#define LFVIO_INIT_MAX_FRAMES 128   /* frames in all_image_frame accepted by lfvio_linear_alignment          */
#define LFVIO_SFM_MAX_FRAMES 24     /* WINDOW_SIZE + 1 accepted by the SfM entries (shared-memory bound)     */
#define LFVIO_SFM_MAX_BA_ITER 50    /* Ceres default max_num_iterations                                      */

/* solveGyroscopeBias (initial/initial_aligment.cpp:3-29) without the repropagation (:31-37; the caller does it
 * with lfvio_preintegrate).  R: n_frames x 9 = ImageFrame::R of every frame of all_image_frame in time order;
 * delta_q (n_frames-1) x 4 and J_rg (n_frames-1) x 9 = pre_integration->delta_q and
 * jacobian.block<3,3>(O_R, O_BG) of frames 1.. .  delta_bg: 3.                                              */
int lfvio_solve_gyro_bias(const double *R, const double *delta_q, const double *J_rg, int n_frames, double *delta_bg,
                          int device);
/* LinearAlignment + RefineGravity (initial/initial_aligment.cpp:56-205).  T: n_frames x 3 = ImageFrame::T;
 * delta_p / delta_v: (n_frames-1) x 3, sum_dt: n_frames-1, of frames 1.. ; tic = TIC[0]; g_norm = G.norm().
 * Outputs: *ok = the function's return value; g (3) = refined gravity in the frame of R/T; x: caller provides
 * 3 n_frames + 4 doubles; on success the first 3 n_frames + 3 hold [v_0 .. v_{n-1} (body frame) | dg (2) | s]
 * with s already divided by 100 (:199-200); when the first solve is rejected (:186-194) x holds that solve's
 * 3 n_frames + 4 unknowns and g its gravity.                                                                 */
int lfvio_linear_alignment(const double *R, const double *T, const double *delta_p, const double *delta_v,
                           const double *sum_dt, int n_frames, const double *tic, double g_norm, double *g, double *x,
                           int *ok, int device);
/* MotionEstimator::solveRelativeRT (initial/solve_5pts.cpp:538-575): spherical 8-point RANSAC
 * (myfindFundamentalMat :266-335, injected random source seed/frame as in lfvio_reject_with_f) and recoverPose
 * (:392-536).  corres: n x 6 = [bearing in the older frame | bearing in the newest frame].  rotation (9) /
 * translation (3) = R^T, -R^T t of the pose recoverPose selects — the reference passes its outputs by value
 * (:556-569) and returns uninitialised memory; this is the evident intent (stated deviation).
 * The function's boolean is (*inlier_cnt > 12); below 15 correspondences nothing is computed and
 * *inlier_cnt = 0.  E_out (9, may be NULL): the refit essential matrix.                                      */
int lfvio_solve_relative_rt(const double *corres, int n, unsigned long long seed, unsigned frame, double *rotation,
                            double *translation, int *inlier_cnt, double *E_out, int device);
/* InitialEXRotation::solveRelativeR (initial/initial_ex_rotation.cpp:220-289; ESTIMATE_EXTRINSIC == 2 only): the same
 * RANSAC (:89-150), the two rotations of the essential matrix made proper (the reference decomposes -E when the first
 * candidate is a reflection, :268-272), the four testTriangulation ratios (:291-318) and `ratio1 > ratio2 ? R1 : R2`,
 * transposed (:283-287).  rotation (9): identity below 9 correspondences (:222, :288).                        */
int lfvio_solve_relative_r(const double *corres, int n, unsigned long long seed, unsigned frame, double *rotation,
                           int device);
/* The linear system of InitialEXRotation::CalibrationExRotation (initial/initial_ex_rotation.cpp:20-60) for the n
 * rotation pairs collected so far: Rc (camera, from lfvio_solve_relative_r), Rimu (delta_q of the pre-integration),
 * Rc_g (ric^-1 Rimu ric with the ric of the time the pair was added), each n x 9.  ric (9): the new estimate;
 * sv (4): singular values, descending — the function returns true when frame_count >= WINDOW_SIZE and
 * sv[2] > 0.25 (:62-68).                                                                                    */
int lfvio_calibrate_ex_rotation(const double *Rc, const double *Rimu, const double *Rc_g, int n, double *ric, double *sv,
                                int device);
/* PnpSolver::compute_pose with set_internal_parameters(0, 0, 1, 1) (pnp_solver.cpp:275-338; call sites
 * initial_sfm.cpp:55-69, estimator.cpp:343-357), batched: problem p uses correspondences
 * [first[p], first[p+1]) of pws (world points) / us (bearings), one thread block each.  R: n_problems x 9,
 * T: n_problems x 3 with p_cam = R p_w + T, err (may be NULL): the function's return value.                  */
int lfvio_epnp(const double *pws, const double *us, const int *first, int n_problems, double *R, double *T, double *err,
               int device);

typedef struct lfvio_sfm_summary {
    int ok;             /* GlobalSFM::construct's return value                                               */
    int fail_stage;     /* 0 none, 1 solveFrameByPnP (< 10 points, initial_sfm.cpp:50-55), 2 BA not converged */
    int fail_frame;     /* frame of a PnP failure                                                            */
    int ba_iterations, ba_successful, ba_converged;
    double ba_final_cost;
    double ba_costs[LFVIO_SFM_MAX_BA_ITER + 1]; /* cost at the start and after every successful step         */
} lfvio_sfm_summary;

/* GlobalSFM::construct (initial/initial_sfm.cpp:127-316).  Feature j has the observations
 * obs_first[j] .. obs_first[j+1] (SFMFeature::observation: frame index obs_frame[o], ascending, and bearing
 * obs_pts[3 o ..]).  l / relative_R / relative_T: from relativePose.  Outputs: q (frame_num x 4) and
 * T (frame_num x 3) = w_R_cam / camera position in the frame of camera l; points (n_features x 3, may be NULL)
 * and state (n_features, may be NULL) = SFMFeature::position / state (sfm_tracked_points).  run_ba = 0 skips the
 * full BA (:230-297, debugging / parity hook).  The BA is Ceres' Levenberg-Marquardt with DENSE_SCHUR restated;
 * its 0.2 s wall-clock limit (:279) is replaced by the default 50-iteration limit.                            */
int lfvio_global_sfm(int frame_num, int l, const double *relative_R, const double *relative_T, const int *obs_first,
                     const int *obs_frame, const double *obs_pts, int n_features, int run_ba, double *q, double *T,
                     double *points, unsigned char *state, lfvio_sfm_summary *summary, int device);
/* The full BA alone (parity hook): c_rot (frame_num x 4, cam_R_w) / c_trans (frame_num x 3) / points are
 * updated in place; rotation l and translations l, frame_num-1 stay constant (:243-251).                     */
int lfvio_sfm_bundle_adjust(int frame_num, int l, const int *obs_first, const int *obs_frame, const double *obs_pts,
                            int n_features, const unsigned char *state, double *c_rot, double *c_trans, double *points,
                            lfvio_sfm_summary *summary, int device);

/* ------------------------------------------------------------------------------------------
 * Live per-stage device timing (CUDA events recorded on the handle's own stream around each kernel
 * group).  Replaces the reference's TicToc prints (feature_tracker.cpp:106,141,153,170,175;
 * estimator.cpp:775,828,915,919) and gives bench.py the kernel durations its roofline line needs.
 * ------------------------------------------------------------------------------------------ */
int lfvio_tracker_stage_count(void);
const char *lfvio_tracker_stage_name(int i);
int lfvio_tracker_profile_begin(lfvio_tracker *t, int max_frames);
/* Synchronises the stream; stage_ms[i] = total device ms, stage_records[i] = timed launches groups. */
int lfvio_tracker_profile_end(lfvio_tracker *t, float *stage_ms, int *stage_records);
int lfvio_window_stage_count(void);
const char *lfvio_window_stage_name(int i);
int lfvio_window_profile_begin(lfvio_window *w, int max_solves);
int lfvio_window_profile_end(lfvio_window *w, float *stage_ms, int *stage_records);

#ifdef __cplusplus
}
#endif
#endif /* LFVIO_H_ */
