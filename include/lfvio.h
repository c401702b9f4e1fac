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

#ifdef __cplusplus
}
#endif
#endif /* LFVIO_H_ */
