// Front-end kernel launchers (device pointers, asynchronous on the given stream).
// Each mirrors one third-party call the reference's FeatureTracker::readImage makes
// (feature_tracker/src/feature_tracker.cpp:95-184); arithmetic follows SURVEY.md Appendix B.
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched from the driver at run time)

#include "common.cuh"
#include "lift.cuh"

namespace lfvio {

constexpr int kMaxPyrLevels = 8;
constexpr int kClaheTiles = 8;       // cv::createCLAHE(3.0, Size(8,8))  feature_tracker.cpp:103
constexpr int kCandHistBins = 1024;  // value histogram used to bound the corner-selection prefix

// One pyramid level: tightly described u8 image in device memory.
struct ImgLevel {
    uint8_t *data;
    int rows, cols, pitch;
};
struct Pyramid {
    ImgLevel lvl[kMaxPyrLevels];
    int levels;  // number of valid levels (maxLevel + 1 after the "<= winSize" early stop)
};

// TMA descriptors of the first pyramid levels (u8, box 64 x 44 covers the 44 x 44 LK tile at any x), passed to lk_kernel as __grid_constant__.
constexpr int kTmaLevels = 4;
struct alignas(64) PyrTmaps {
    CUtensorMap m[kTmaLevels];
    int valid;  // 0: descriptors could not be encoded (the kernel then stages every tile with plain loads)
};
// Encodes one descriptor per level with cuTensorMapEncodeTiled (entry point resolved through cudaGetDriverEntryPoint).
void make_pyramid_tmaps(const Pyramid &p, PyrTmaps *out);

// ---- K1 CLAHE -------------------------------------------------------------------------------
// hist: 64*256 u32 scratch (zeroed inside), lut: 64*256 u8 scratch.
int launch_clahe(const uint8_t *src, int src_pitch, uint8_t *dst, int dst_pitch, int rows, int cols,
                 float clip_limit, uint32_t *hist, uint8_t *lut, int n_sm, cudaStream_t s);

// ---- K2 pyramid -----------------------------------------------------------------------------
int launch_pyr_down(const ImgLevel &src, const ImgLevel &dst, cudaStream_t s);
// Fills p.lvl[1..] from p.lvl[0]; p.levels must be set by make_pyramid_layout.
int launch_build_pyramid(const Pyramid &p, cudaStream_t s);
// Computes level geometry (OpenCV buildOpticalFlowPyramid early stop) — no allocation.
void make_pyramid_layout(Pyramid *p, int rows, int cols, int max_level, int win);

// ---- K3 pyramidal LK ------------------------------------------------------------------------
struct LkParams {
    int win;            // 41
    int max_iter;       // 30
    double eps2;        // (0.01)^2, compared in f64 like Point2f::ddot
    double min_eig_thr; // 1e-4 (double in OpenCV; the f32 minEig is promoted for the compare)
    int border_check;   // fuse inBorder() (feature_tracker.cpp:6-12,129-131) into status
    int rows, cols;     // level-0 size for inBorder
};
// n_dev: optional device count (if non-null the kernel uses min(*n_dev, n_max)); grid is n_max CTAs.
int launch_lk(const Pyramid &prev, const PyrTmaps &prev_tm, const Pyramid &next, const float2 *prev_pts, float2 *next_pts,
              uint8_t *status, int n_max, const int *n_dev, const LkParams &prm, cudaStream_t s);

// ---- K4 Shi-Tomasi --------------------------------------------------------------------------
// eig: rows x cols f32 (pitch in elements = cols).  max_bits: one u32, ordered-float encoding.
int launch_min_eig(const uint8_t *img, int pitch, const uint8_t *mask, int mask_pitch, float *eig, int rows,
                   int cols, uint32_t *max_bits, cudaStream_t s, const CUtensorMap *tile_map = nullptr);

struct CornerScratch {
    unsigned long long *cand;       // capacity cand_cap: (value bits << 32) | linear index
    unsigned long long *cand_sorted;  // cell-sorted copy
    int cand_cap;
    int *cell_count;   // gw*gh + 1
    int *cell_start;   // gw*gh + 1
    uint32_t *hist;    // kCandHistBins
    uint8_t *state;    // cand_cap
    int *counters;     // 8 ints: [0]=n_cand [1]=overflow flag [2]=n_selected [3]=enc(threshold) [4]=hist shift
    int gw, gh, cell;  // grid of cell x cell pixels
};
// Zeroes the scratch counters, finds NMS survivors above quality*max under the mask.
int launch_corner_candidates(const float *eig, const uint8_t *mask, int mask_pitch, int rows, int cols,
                             const uint32_t *max_bits, double quality, const CornerScratch &cs, cudaStream_t s);
// Greedy min-distance selection in descending (value, index) order.  Appends up to
// (max_total - *n_inout) corners at out_pts[*n_inout ...] and bumps *n_inout (device counters);
// if n_inout is null, max_total is the plain maxCorners and out count is written to cs.counters[2].
int launch_corner_select(const CornerScratch &cs, int cols, float min_distance, int max_total, int *n_inout,
                         float2 *out_pts, int *out_ids, int *out_cnt, uint8_t *out_prev_valid, int out_cap,
                         cudaStream_t s);

// ---- K6 mask + track bookkeeping ----------------------------------------------------------------
struct TrackArrays {  // SoA per-track state, capacity cap
    float2 *cur;      // cur_pts (previous frame position)
    float2 *forw;     // forw_pts (this frame position)
    int *ids;
    int *cnt;
    float *bearing;       // 3 f32 per track: cur_un_pts_3d of the previous undistortedPoints()
    uint8_t *prev_valid;  // id was != -1 when `bearing` was stored (prev_un_pts_map_3d membership)
};
// Stable compaction by status, track_cnt++ (feature_tracker.cpp:132-145).  n_dev in/out.
int launch_compact(const TrackArrays &in, const TrackArrays &out, const uint8_t *status, int *n_dev, int cap, int cnt_add,
                   cudaStream_t s);

// ---- K7 rejectWithF (feature_tracker.cpp:329-386): spherical 8-point RANSAC, ransac.cu --------------------------
struct RansacScratch {
    double3 *b1, *b2;  // cap unit bearings of cur_pts / forw_pts
    double *E;         // 100 x 9 hypotheses
    float *score;      // 100
    uint8_t *flags;    // 100 x cap inlier flags per hypothesis
    double *E_best;    // 9: refit essential matrix (row-major)
    int *best;         // index of the winning hypothesis
};
// status[i] = 1 for the tracks rejectWithF keeps (all ones below 8 tracks).  seed/frame: injected random source.
int launch_reject_with_f(const OcamDev &cam, const float2 *cur, const float2 *forw, const int *n_dev, int n_upper, int cap,
                         unsigned long long seed, unsigned frame, const RansacScratch &rs, uint8_t *status, cudaStream_t s,
                         const unsigned *frame_dev = nullptr /* device word overriding `frame` (graph replay) */);
// the RANSAC alone on bearings already in rs.b1 / rs.b2 (estimator-side copy of the same code, solve_5pts.cpp:266-335)
int launch_ransac_bearings(const int *n_dev, int n_upper, int cap, unsigned long long seed, unsigned frame, const RansacScratch &rs,
                           uint8_t *status, cudaStream_t s);
// setMask (feature_tracker.cpp:46-83): sort (libstdc++ introsort order), greedy keep, reorder.
// half_width: circle_spans(min_dist) table (min_dist+1 ints), base_mask: the annulus.
int launch_set_mask_select(const TrackArrays &in, const TrackArrays &out, int *n_dev, int cap,
                           const uint8_t *base_mask, int mask_pitch, const int *half_width, int min_dist,
                           int *sort_scratch, cudaStream_t s);
// mask = base; stamp a zero disc at each kept track.
int launch_mask_raster(uint8_t *mask, const uint8_t *base_mask, int rows, int cols, int pitch,
                       const float2 *pts, const int *n_dev, int cap, const int *half_width, int min_dist,
                       cudaStream_t s);
// Filled circle raster used to build the annulus (cv::circle(.., -1)).
int launch_fill_circle(uint8_t *img, int rows, int cols, int pitch, int cx, int cy, int radius,
                       const int *half_width_dev, uint8_t value, cudaStream_t s);
// undistortedPoints (feature_tracker.cpp:441-504): lift + normalise in f64, cast to f32, velocity.
int launch_undistort(const OcamDev &cam, const float2 *pts, const int *ids, float *bearing_prev_inout,
                     uint8_t *prev_valid_inout, float *bearing_out, float *vel_out, const int *n_dev, int cap,
                     double dt, const int *have_prev_in, int *have_prev_out, cudaStream_t s,
                     const double *dt_dev = nullptr /* device double overriding `dt` (graph replay) */);
// updateID loop (feature_tracker.cpp:388-398, feature_tracker_node.cpp:103-111).
int launch_update_ids(int *ids, const int *n_dev, int cap, int *n_id_dev, cudaStream_t s);

// host helper: midpoint-circle half widths of cv::circle (filled)
void circle_half_widths(int radius, int *hw /* radius+1 */);

}  // namespace lfvio
